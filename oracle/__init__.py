"""CPU oracle for the prune + 1-D k-means hot path.  TEST INFRASTRUCTURE ONLY (see nnc_oracle.c)."""
