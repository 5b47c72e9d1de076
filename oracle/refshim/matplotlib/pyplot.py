"""Stand-in: the hot path is always run with plot=False."""
