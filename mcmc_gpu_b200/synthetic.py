"""Deterministic synthetic ice-sheet grids for tests and bench (SURVEY.md §8d recipe).

Pure data generation in numpy; no compute path lives here.  The constants mirror the tutorial/driver
settings of the reference (largeScaleChain_multiprocessing.py:554-598, T3_LargeScaleChain.ipynb cells 40-54)
on an analytic geometry so every box produces byte-identical inputs.
"""
from __future__ import annotations

import numpy as np

RESOLUTION = 500.0


def make_grids(H: int, W: int, resolution: float = RESOLUTION, radar_frac: float = 0.02, seed: int = 0) -> dict:
    """Return the dict of [H,W] float64 fields and masks a chain is constructed from."""
    res = float(resolution)
    xx, yy = np.meshgrid(np.arange(W) * res, np.arange(H) * res)
    surf = 2000.0 + 100.0 * np.sin(xx / 5e4) + 50.0 * np.cos(yy / 7e4)
    bed0 = surf - 1500.0 + 200.0 * np.sin(xx / 2e4) * np.cos(yy / 3e4)
    velx = 100.0 * np.cos(yy / 9e4) + 10.0
    vely = 50.0 * np.sin(xx / 8e4)
    g = np.random.default_rng(seed)
    dhdt = 0.1 * g.standard_normal((H, W))
    smb = 0.3 + 0.05 * g.standard_normal((H, W))
    data_mask = (g.random((H, W)) < radar_frac).astype(np.int64)
    cond_bed = np.where(data_mask == 1, bed0, np.nan)
    grounded_ice_mask = np.ones((H, W), dtype=np.int64)
    highvel_mask = (np.hypot(velx, vely) > 60.0).astype(np.int64)
    return dict(xx=xx, yy=yy, surf=surf, bed0=bed0, velx=velx, vely=vely, dhdt=dhdt, smb=smb,
                data_mask=data_mask, cond_bed=cond_bed, grounded_ice_mask=grounded_ice_mask,
                highvel_mask=highvel_mask, resolution=res)


def chain_initial_beds(bed0: np.ndarray, n_chains: int, amplitude: float = 5.0) -> np.ndarray:
    """[C,H,W] initial beds: bed0 plus a smooth chain-specific bump (chain 0 is bed0 itself)."""
    H, W = bed0.shape
    jj, ii = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    out = np.empty((n_chains, H, W), dtype=np.float64)
    for c in range(n_chains):
        g = np.random.default_rng(10_000 + c)
        ph = g.uniform(0.0, 2.0 * np.pi, size=2)
        amp = 0.0 if c == 0 else amplitude
        out[c] = bed0 + amp * np.sin(ii / 37.0 + ph[0]) * np.cos(jj / 41.0 + ph[1])
    return out


# RandField / chain constants of the large-scale tutorial configuration.
RF_KW = dict(range_min_x=10e3, range_max_x=50e3, range_min_y=10e3, range_max_y=50e3,
             scale_min=50.0, scale_max=150.0, nugget_max=0.0, model_name="Matern", isotropic=True,
             smoothness=0.9)
BLOCKS = (50, 80, 50, 80)            # set_block_sizes(min_x, max_x, min_y, max_y)
LOGISTIC = (2.0, 0.0, 6.0, 1.0)      # set_weight_param(L, x0, k, offset, ...)
MAX_DIST = 30e3
SIGMA_MC = 5.0
