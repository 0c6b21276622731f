"""String constants of the on-disk graph layout (names mirror ``deeprank2.domain``; only what the GNN training path reads)."""
