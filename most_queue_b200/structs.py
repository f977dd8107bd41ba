"""Result records returned by the simulators.

Field names and order follow the reference's most_queue/structs.py (QueueResults :18-30,
PriorityResults :48-56, NetworkResults :71-84); the extra fields (`ci`, `replications`, ...) are
additive and default to None so reference-style code keeps working.
"""

from dataclasses import dataclass, field


@dataclass
class QueueResults:
    v: list | None = None  # sojourn time raw moments
    w: list | None = None  # waiting time raw moments
    p: list | None = None  # probabilities of states
    pi: list | None = None
    utilization: float | None = None
    duration: float = 0.0
    # additive: replication statistics
    replications: int = 1
    ci: dict | None = None  # 99 % confidence half-widths across replications: {"w": [...], "v": [...], "p": [...]}
    timing: dict | None = None  # device time of the kernels of the run (ms)


@dataclass
class VacationResults(QueueResults):
    """Result of a queue with vacations (structs.py:60-68 of the reference)."""

    warmup_prob: float = 0
    cold_prob: float = 0
    cold_delay_prob: float = 0
    servers_busy_probs: list | None = None


@dataclass
class NegativeArrivalsResults(QueueResults):
    """Results of a queue with negative arrivals (structs.py:147-158 of the reference)."""

    v_served: list | None = None
    v_broken: list | None = None
    q: float | None = None


@dataclass
class MulticlassResults:
    v: list | None = None
    w: list | None = None
    p: list | None = None
    utilization: float | None = None
    duration: float = 0.0


@dataclass
class PriorityResults(MulticlassResults):
    h: list | None = None
    busy: list | None = None
    w_with_pr: list | None = None
    replications: int = 1
    ci: dict | None = None


@dataclass
class NetworkResults:
    v: list | None = None
    intensities: list | None = None
    loads: list | None = None
    duration: float = 0.0
    arrived: int = 0
    served: int = 0
    w: list | None = None  # additive: network waiting-time moments (reference keeps them on the object)
    replications: int = 1
    ci: dict | None = field(default=None)


@dataclass
class NetworkResultsPriority:
    """Per-class network results (structs.py:132-145 of the reference); w, replications, ci are additive."""

    v: list | None = None
    intensities: list | None = None
    loads: list | None = None
    duration: float = 0.0
    arrived: list | int = 0
    served: list | int = 0
    w: list | None = None
    replications: int = 1
    ci: dict | None = field(default=None)


@dataclass
class NetworkMeansResults(NetworkResults):
    """Mean-value results for open networks (structs.py:102-113 of the reference)."""

    mean_jobs: list | None = None  # L_i: mean number of jobs at each node
    v_node: list | None = None
    negative_intensities: list | None = None
