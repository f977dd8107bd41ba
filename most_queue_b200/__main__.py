"""`python -m most_queue_b200` builds the CUDA library."""
from .build import build

if __name__ == "__main__":
    print(build(verbose=True))
