"""Empty stand-in so the reference imports in a container without matplotlib (oracle/make_golden.py only)."""
