"""CPU oracle for the differentiable-A* hot path.  TEST INFRASTRUCTURE ONLY (see astar_oracle.c)."""
