"""Distribution parameters and host-side helpers (mirrors most_queue.random)."""
