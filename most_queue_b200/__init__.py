"""most_queue_b200 -- B200-native batched discrete-event simulation engine for most-queue's hot
path (QsSim / PriorityQueueSimulator / NetworkSimulator run() loops).

Python here is the host-side mirror of the reference's object API; all simulation happens in
hand-written sm_100a CUDA kernels behind the C ABI of libmqb200.so (include/mqb200.h).
"""

__version__ = "0.1.0"
