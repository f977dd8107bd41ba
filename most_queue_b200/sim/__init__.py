"""Simulators with the reference's class names (mirrors most_queue.sim)."""
