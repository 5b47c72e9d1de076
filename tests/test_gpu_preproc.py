"""get_highvel_boundary on the GPU (mode-filter kernel + exact nearest-region distance) vs the reference's output."""
import os

import numpy as np
import pytest

from cases import highvel_case_inputs
from oracle import preproc_oracle as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_highvel_boundary_equals_reference():
    from mcmc_gpu_b200 import Topography
    hb = highvel_case_inputs()
    got = Topography.get_highvel_boundary(hb["velx"], hb["vely"], hb["threshold"], hb["grounded"], hb["ocean"], hb["distance_max"],
                                          hb["xx"], hb["yy"], smooth_mode=hb["smooth_mode"])
    gold = np.load(os.path.join(GOLD, "highvel_boundary.npz"))["mask_final"]
    assert got.dtype == gold.dtype and np.array_equal(got, gold)


@pytest.mark.parametrize("H,W,size", [(1, 1, 10), (2, 3, 3), (37, 53, 10), (64, 200, 5), (301, 299, 4)])
def test_mode_filter_kernel_matches_oracle(H, W, size):
    import torch
    from mcmc_gpu_b200 import _lib
    g = np.random.default_rng(H * W + size)
    img = ((g.random((H, W)) < 0.45) * 255).astype(np.uint8)
    dev = _lib.require_cuda()
    a = torch.as_tensor(img).to(dev)
    b = torch.empty_like(a)
    _lib.check(_lib.load().gmc_mode_filter_binary(dev.index, a.data_ptr(), b.data_ptr(), H, W, size,
                                                   torch.cuda.current_stream().cuda_stream))
    assert np.array_equal(b.cpu().numpy(), P.mode_filter_binary(img, size))


def test_larger_grid_against_oracle_and_empty_region():
    from mcmc_gpu_b200 import Topography, synthetic as syn
    g = syn.make_grids(150, 170)
    vel = np.hypot(g["velx"], g["vely"])
    thr = float(np.quantile(vel, 0.7))
    grounded = np.ones(vel.shape, dtype=np.int64)
    grounded[:, :12] = 0
    ocean = 1 - grounded
    args = (g["velx"], g["vely"], thr, grounded, ocean, 3000.0, g["xx"], g["yy"])
    assert np.array_equal(Topography.get_highvel_boundary(*args), P.highvel_boundary(*args))
    none = Topography.get_highvel_boundary(g["velx"], g["vely"], 1e9, grounded, np.zeros_like(grounded), 3000.0, g["xx"], g["yy"])
    assert none.shape == vel.shape and not none.any()
