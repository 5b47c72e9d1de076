"""Empty stand-in package: not touched by the hot path (oracle/make_golden.py only)."""
