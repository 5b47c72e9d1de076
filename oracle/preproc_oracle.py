"""CPU restatement of the preprocessing step that produces the hot path's region mask — TEST
INFRASTRUCTURE ONLY (tests/, smoke, bench cpu legs).

Follows `Topography.get_highvel_boundary` (/root/reference/gstatsMCMC/Topography.py:546-571).  The mode filter restates
PIL's `ImageFilter.ModeFilter` (Pillow `src/libImaging/ModeFilter.c`: window of half-width size//2 clipped to the image,
most frequent value, lower value on ties, pixel kept when no value occurs more than twice); PIL is a third-party
dependency of the reference, present in this container, and `tests/test_oracle_preproc.py` checks the restatement
against it.  Pinned against the reference itself through `tests/golden/highvel_boundary.npz` (oracle/make_golden.py).
"""
from __future__ import annotations

import numpy as np


def mode_filter_binary(img, size):
    """PIL ModeFilter(size) on a uint8 image with values {0, 255}."""
    img = np.asarray(img, dtype=np.uint8)
    H, W = img.shape
    r = size // 2
    ones = (img != 0).astype(np.int64)
    # windowed counts through a summed-area table (exact integers)
    sat = np.zeros((H + 1, W + 1), dtype=np.int64)
    sat[1:, 1:] = ones.cumsum(0).cumsum(1)
    y0, y1 = np.maximum(np.arange(H) - r, 0), np.minimum(np.arange(H) + r, H - 1) + 1
    x0, x1 = np.maximum(np.arange(W) - r, 0), np.minimum(np.arange(W) + r, W - 1) + 1
    c1 = sat[y1[:, None], x1[None, :]] - sat[y0[:, None], x1[None, :]] - sat[y1[:, None], x0[None, :]] + sat[y0[:, None], x0[None, :]]
    n = (y1 - y0)[:, None] * (x1 - x0)[None, :]
    c0 = n - c1
    out = np.where(c1 > c0, 255, 0).astype(np.uint8)
    return np.where(np.maximum(c0, c1) > 2, out, img)


def highvel_boundary(velx, vely, velmag_threshold, grounded_ice_mask, ocean_mask, distance_max, xx, yy, smooth_mode=10):
    """Topography.py:546-571, with the double loop over cells (:563-565) as one broadcast per row."""
    mask = (grounded_ice_mask) & (np.sqrt(velx ** 2 + vely ** 2) >= velmag_threshold)          # :548
    mask = mask | ocean_mask                                                                  # :549
    mask_mat = np.array(mode_filter_binary((mask * 255).astype(np.uint8), smooth_mode) / 255, dtype=int)   # :551-553
    hard = (mask_mat == 1) & (grounded_ice_mask == 1)                                          # :559-560
    px, py = xx[hard], yy[hard]
    mask_dist = np.empty(xx.shape)
    for i in range(xx.shape[0]):                                                               # :563-565
        d = np.sqrt(np.square(yy[i, :, None] - py[None, :]) + np.square(xx[i, :, None] - px[None, :]))
        mask_dist[i] = d.min(1) if px.size else np.nan
    return (mask_dist < distance_max) & grounded_ice_mask                                      # :568
