"""Open-network simulators computed on the GPU: NetworkSimulator, PriorityNetwork, TandemBlockingSim."""
