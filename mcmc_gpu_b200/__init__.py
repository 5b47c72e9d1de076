"""mcmc_gpu_b200 — B200-native (sm_100a) many-chain MCMC step for gstatsMCMC's large-scale chain.

`from mcmc_gpu_b200 import MCMC, Topography` mirrors `from gstatsMCMC import MCMC, Topography` for the hot path.
Importing the package does not need a GPU; any computation does (no CPU fallback).
"""
from . import synthetic  # noqa: F401

__all__ = ["MCMC", "Topography", "Utilities", "synthetic", "drivers"]


def __getattr__(name):
    if name in ("MCMC", "Topography", "Utilities", "drivers", "_lib", "sgs_tables"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
