"""Constants shared with the reference (most_queue/constants.py:6)."""

DEFAULT_NUM_STATES = 100000
