"""Shared builders for the -m gpu parity tests."""
from __future__ import annotations

import contextlib
import io

import numpy as np

from cases import build_case_grids
from oracle import crf_oracle as O


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def oracle_setup(case):
    g = build_case_grids(case)
    cs, fp = O.setup_from_grids(g, sigma_mc=case["sigma_mc"], logistic=case["logistic"], max_dist=case["max_dist"],
                                blocks=case["blocks"], rf_kw=case["rf_kw"], update_in_region=case["update_in_region"],
                                block_type=case["block_type"])
    return g, cs, fp


def product_chain(case, g=None):
    """Build mcmc_gpu_b200's chain_crf + RandField exactly like the reference tutorial does."""
    from mcmc_gpu_b200 import MCMC
    g = build_case_grids(case) if g is None else g
    kw = case["rf_kw"]
    rf = MCMC.RandField(kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
                        kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"],
                        smoothness=kw.get("smoothness"), rng_seed=case["rf_seed"])
    rf.set_block_sizes(*case["blocks"], steps=case.get("steps", 5))
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    rf.set_generation_method(True)
    ch = quiet(MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
               g["cond_bed"], g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    if case["update_in_region"]:
        quiet(ch.set_update_region, True, g["highvel_mask"])
    else:
        quiet(ch.set_update_region, False)
    ch.set_loss_type(sigma_mc=case["sigma_mc"], massConvInRegion=True)
    quiet(ch.set_update_type, case["block_type"])
    ch.set_crf_data_weight(rf)
    ch.set_random_generator(case["chain_seed"])
    return ch, rf, g


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def same_values(a, b):
    """Equal as numbers (NaN == NaN, +0 == -0)."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
