"""Mirror of the reference's gstatsMCMC/Utilities.py helper that sits next to the hot path (setup of the block tapers and
the conditioning weight)."""
from __future__ import annotations

import numpy as np

from . import _lib


def min_dist_from_mask(xx, yy, mask):
    """Distance of every grid point to the nearest point where `mask` is True (reference Utilities.py:21-24), on the GPU:
    exact brute-force scan (kernel `min_dist_kernel`), bit-identical to the reference's KD-tree query."""
    import torch
    dev = _lib.require_cuda()
    lib = _lib.load()
    xx = np.asarray(xx, dtype=np.float64)
    yy = np.asarray(yy, dtype=np.float64)
    mask = np.asarray(mask, dtype=bool)
    if not mask.any():
        raise ValueError("mask selects no point")
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).to(dev)      # noqa: E731
    px, py, qx, qy = cu(xx[mask]), cu(yy[mask]), cu(xx.ravel()), cu(yy.ravel())
    out = torch.empty(qx.numel(), dtype=torch.float64, device=dev)
    _lib.check(lib.gmc_min_dist(dev.index, px.data_ptr(), py.data_ptr(), px.numel(), qx.data_ptr(), qy.data_ptr(), qx.numel(),
                                out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out.cpu().numpy().reshape(xx.shape)
