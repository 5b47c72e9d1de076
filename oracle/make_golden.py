#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE (read-only at /root/reference).

Only runs in the build container (the GPU box has no /root/reference); the produced fixtures are committed.
    python oracle/make_golden.py
The reference imports packages that are absent here and unused on the hot path (matplotlib, gstools, ...);
`oracle/refshim/` holds empty stand-ins for them.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]

with contextlib.redirect_stdout(io.StringIO()):
    from gstatsMCMC import MCMC, Topography          # noqa: E402  (the reference)

from mcmc_gpu_b200 import synthetic as syn          # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


# ---- the named trajectory cases: shared with tests/cases.py so oracle and reference see the same inputs ----
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cases import (TRAJECTORY_CASES, FIELD_CASES, SGS_CASES, residual_case_inputs, build_case_grids,   # noqa: E402
                   build_sgs_inputs, highvel_case_inputs, SGS_GRID_CASES, sgs_grid_inputs)


def reference_chain(case):
    g = build_case_grids(case)
    rf = quiet(MCMC.RandField, *[case["rf_kw"][k] for k in
               ("range_min_x", "range_max_x", "range_min_y", "range_max_y", "scale_min", "scale_max",
                "nugget_max", "model_name", "isotropic")], smoothness=case["rf_kw"].get("smoothness"),
               rng_seed=case["rf_seed"])
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
    out = quiet(ch.run, case["n_iter"], rf, only_save_last_bed=True, info_per_iter=10 ** 9, plot=False,
                progress_bar=False)
    bed, loss_mc, loss_data, loss, steps, resampled, blocks = out
    return dict(bed=bed, loss_mc=loss_mc, loss_data=loss_data, loss=loss, steps=steps, resampled_times=resampled,
                blocks=blocks, crf_weight=ch.crf_data_weight,
                edge_mask0=rf.edge_masks[0], edge_mask_last=rf.edge_masks[-1], pairs=rf.pairs)


def reference_sgs_chain(case):
    g = build_sgs_inputs(case)
    ch = quiet(MCMC.chain_sgs, g["xx"], g["yy"], g["bed_init"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
               g["cond_bed"], g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=case["sigma_mc"], massConvInRegion=True)
    ch.set_block_sizes(*case["blocks"])
    ch.set_normal_transformation(g.get("nst"), do_transform=case["transform"])
    ch.set_trend(g["trend"], detrend_map=case["detrend"])
    v = case["vario"]
    quiet(ch.set_variogram, v["vtype"], v["range"], v["sill"], v["nugget"], isotropic=v["isotropic"],
          vario_smoothness=v["smoothness"], vario_azimuth=v["azimuth"])
    quiet(ch.set_sgs_param, case["neighbors"], case["radius"])
    ch.set_random_generator(case["seed"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = quiet(ch.run, case["n_iter"], only_save_last_bed=True, info_per_iter=10 ** 9, plot=False, progress_bar=False)
    bed, loss_mc, loss_data, loss, steps, resampled, blocks = out
    return dict(bed=bed, loss_mc=loss_mc, loss_data=loss_data, loss=loss, steps=steps, resampled_times=resampled, blocks=blocks)


def main():
    os.makedirs(OUT, exist_ok=True)

    # (4) small-scale SGS chain trajectories
    for name, case in SGS_CASES.items():
        out = reference_sgs_chain(case)
        np.savez_compressed(os.path.join(OUT, f"sgs_{name}.npz"), **out)
        print(f"sgs_{name}: final loss {out['loss'][-1]!r}, acceptance {out['steps'].mean():.3f}")

    # (6) whole-grid SGS realisations with bounds: gstatsim_custom/interpolate.sgs
    from gstatsMCMC.gstatsim_custom import interpolate
    import warnings
    for name, case in SGS_GRID_CASES.items():
        gi = sgs_grid_inputs(case)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sim = interpolate.sgs(gi["xx"], gi["yy"], gi["cond"], gi["vario"], radius=case["radius"], num_points=case["num_points"],
                                  bounds=gi["bounds"], seed=np.random.default_rng(case["seed"]), quiet=True)
        np.savez_compressed(os.path.join(OUT, f"sgs_grid_{name}.npz"), sim=sim)
        print(f"sgs_grid_{name}: mean {np.nanmean(sim):.3f}, std {np.nanstd(sim):.3f}")

    # (5) region-mask preprocessing: Topography.get_highvel_boundary (the O(N^2) double loop, so a small grid)
    hb = highvel_case_inputs()
    out = Topography.get_highvel_boundary(hb["velx"], hb["vely"], hb["threshold"], hb["grounded"], hb["ocean"], hb["distance_max"],
                                          hb["xx"], hb["yy"], smooth_mode=hb["smooth_mode"])
    np.savez_compressed(os.path.join(OUT, "highvel_boundary.npz"), mask_final=out)
    print("highvel_boundary:", out.shape, out.dtype, int(out.sum()), "cells")

    # (1) residual + loss known answers
    ri = residual_case_inputs()
    res = Topography.get_mass_conservation_residual(ri["bed"], ri["surf"], ri["velx"], ri["vely"], ri["dhdt"],
                                                    ri["smb"], ri["resolution"])
    ch = quiet(MCMC.chain_crf, ri["xx"], ri["yy"], ri["bed"], ri["surf"], ri["velx"], ri["vely"], ri["dhdt"],
               ri["smb"], ri["bed"], ri["mask"], ri["mask"], ri["resolution"])
    quiet(ch.set_update_region, True, ri["mask"])
    ch.set_loss_type(sigma_mc=ri["sigma_mc"], massConvInRegion=True)
    loss = ch.loss(res, 0)
    np.savez_compressed(os.path.join(OUT, "residual_loss.npz"), residual=res, loss=np.array(loss, dtype=np.float64))
    print("residual_loss: loss =", loss[0])

    # (2) spectral fields
    fields = {}
    for name, fc in FIELD_CASES.items():
        kw = fc["rf_kw"]
        rf = quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"],
                   kw["scale_min"], kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"],
                   smoothness=kw.get("smoothness"), rng_seed=fc["seed"])
        for k, shape in enumerate(fc["shapes"]):
            fields[f"{name}__{k}"] = MCMC.spectral_synthesis_field(rf, tuple(shape), res=fc["res"])
    np.savez_compressed(os.path.join(OUT, "spectral_fields.npz"), **fields)
    print("spectral_fields:", len(fields), "fields")

    # (3) trajectories
    for name, case in TRAJECTORY_CASES.items():
        out = reference_chain(case)
        np.savez_compressed(os.path.join(OUT, f"traj_{name}.npz"), **out)
        print(f"traj_{name}: final loss {out['loss'][-1]!r}, acceptance {out['steps'].mean():.3f}")


if __name__ == "__main__":
    main()
