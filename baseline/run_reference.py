#!/usr/bin/env python
"""Run the UNMODIFIED reference (baseline/_ref, see install_ref.py) on the synthetic workload and print ONE JSON line.

    python baseline/run_reference.py cpu_mp  --grid 500 --iters 1000 [--chains N] [--steps K] [--warmup W]
    python baseline/run_reference.py gpu     --grid 500 --iters 300  [--procs P]
    python baseline/run_reference.py sgs     --grid 300 --iters 8

cpu_mp : `largeScaleChain_mp` (largeScaleChain_multiprocessing.py:19-98; one mp.Pool worker per chain, checkpoint I/O
         included, seed folders pre-created as its __main__ does at :616-620) -> chain-steps/s on this host's cores.
gpu    : the reference's own torch path `chain_crf_gpu.run` (gstatsMCMC/MCMC_gpu.py:233) on cuda:0 -> it/s of one chain
         and, with --procs P, of P concurrent single-chain processes (the only way the reference runs many chains).
sgs    : single-process `chain_sgs.run` (MCMC.py:1599) -> it/s.

This script is measurement infrastructure: it imports the reference from baseline/_ref with the empty stand-in
packages of oracle/refshim (matplotlib, gstools, ... - none on the hot path) and nothing from the product except the
synthetic-grid recipe (mcmc_gpu_b200/synthetic.py, pure numpy data generation).  It runs in its own process so the
reference's fork-based pool never sees the bench process's CUDA context.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def _import_reference():
    if not os.path.isdir(os.path.join(REF, "gstatsMCMC")):
        raise SystemExit(json.dumps({"unavailable": "baseline/_ref is absent (run baseline/install_ref.py where /root/reference exists)"}))
    sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF, ROOT]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def build_lsc(MCMC, syn, grid, chain_cls=None, rf_cls=None):
    """The tutorial / driver configuration on the synthetic grid (largeScaleChain_multiprocessing.py:554-598)."""
    g = syn.make_grids(grid, grid)
    kw = syn.RF_KW
    rf = quiet(rf_cls or MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
               kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*syn.BLOCKS)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, g["resolution"])
    rf.set_generation_method(True)
    ch = quiet(chain_cls or MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
               g["cond_bed"], g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    quiet(ch.set_update_type, "CRF_weight")
    ch.set_crf_data_weight(rf)
    ch.set_random_generator(1000)
    return g, ch, rf


def cpu_mp(a):
    _import_reference()
    import numpy as np
    with contextlib.redirect_stdout(io.StringIO()):
        from gstatsMCMC import MCMC
        import largeScaleChain_multiprocessing as drv
    from mcmc_gpu_b200 import synthetic as syn
    cores = a.chains or os.cpu_count() or 1
    g, ch, rf = build_lsc(MCMC, syn, a.grid)
    beds = list(syn.chain_initial_beds(g["bed0"], cores))

    def one(step, n_iter):
        out = tempfile.mkdtemp(prefix="gmc_ref_")
        seeds = [100000 + 1000 * step + c for c in range(cores)]          # str(seed)[:6] must be unique per chain
        assert len({str(s)[:6] for s in seeds}) == cores, "chains would share a checkpoint folder and resume from each other"
        for s in seeds:
            os.makedirs(os.path.join(out, "LargeScaleChain", str(s)[:6]), exist_ok=True)
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = drv.largeScaleChain_mp(cores, cores, ch, rf, beds, seeds, [n_iter] * cores, out)
        dt = time.perf_counter() - t0
        shutil.rmtree(out, ignore_errors=True)
        acc = float(np.mean([r[4].mean() for r in res]))
        return dt, acc
    for w in range(a.warmup):
        one(800 + w, max(a.iters // 10, 3))                               # 900000 + c: still six digits, so folders stay distinct
    times, acc = [], 0.0
    for s in range(max(a.steps, 1)):
        dt, acc = one(s, a.iters)
        times.append(dt)
    total = sum(times)
    print(json.dumps({"mode": "cpu_mp", "value": cores * (a.iters - 1) * len(times) / total, "unit": "chain-steps/s", "cores": cores,
                      "seconds_per_step": total / len(times), "steps": len(times), "iters": a.iters, "grid": a.grid,
                      "acceptance_rate": acc, "kind": "reference",
                      "what": "largeScaleChain_mp (largeScaleChain_multiprocessing.py:19) of the unmodified reference, one worker per chain, "
                              "pool start-up and checkpoint I/O included"}))


def _gpu_one(grid, iters, seed):
    import numpy as np
    import torch
    with contextlib.redirect_stdout(io.StringIO()):
        from gstatsMCMC import MCMC, MCMC_gpu
    from mcmc_gpu_b200 import synthetic as syn
    rf_cls = getattr(MCMC_gpu, "RandField", MCMC.RandField)
    g, ch, rf = build_lsc(MCMC, syn, grid, chain_cls=MCMC_gpu.chain_crf_gpu, rf_cls=rf_cls)
    ch.set_random_generator(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        ch.run(max(iters // 10, 5), rf, only_save_last_bed=True, info_per_iter=10 ** 9, plot=False, progress_bar=False)   # warm-up
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = ch.run(iters, rf, only_save_last_bed=True, info_per_iter=10 ** 9, plot=False, progress_bar=False)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    steps = out[4]
    steps = steps.detach().cpu().numpy() if hasattr(steps, "detach") else np.asarray(steps)
    return (iters - 1) / dt, float(np.mean(steps)), str(getattr(ch, "device", "?"))


def _gpu_worker(args):
    sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), REF, ROOT]
    return _gpu_one(*args)


def gpu(a):
    _import_reference()
    import torch
    one, acc, dev = _gpu_one(a.grid, a.iters, 1000)
    res = {"mode": "gpu", "it_per_s_one_chain": one, "acceptance_rate": acc, "device": dev, "cuda": bool(torch.cuda.is_available()),
           "iters": a.iters, "grid": a.grid, "dtype": "f32",
           "what": "chain_crf_gpu.run (gstatsMCMC/MCMC_gpu.py:233) of the unmodified reference, torch float32"}
    if a.procs > 1:
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        with ctx.Pool(a.procs) as pool:
            t0 = time.perf_counter()
            rs = pool.map(_gpu_worker, [(a.grid, a.iters, 2000 + p) for p in range(a.procs)])
            wall = time.perf_counter() - t0
        res["procs"] = a.procs
        res["chain_steps_per_s_concurrent"] = float(sum(r[0] for r in rs))       # sum of the processes' own timed rates
        res["wall_s_incl_startup"] = wall
    print(json.dumps(res))


def sgs(a):
    _import_reference()
    import numpy as np
    from scipy.ndimage import gaussian_filter
    from sklearn.preprocessing import QuantileTransformer
    with contextlib.redirect_stdout(io.StringIO()):
        from gstatsMCMC import MCMC
    from mcmc_gpu_b200 import synthetic as syn
    H = W = a.grid
    g = syn.make_grids(H, W)
    bed = g["bed0"] + gaussian_filter(np.random.default_rng(99).standard_normal((H, W)), 2.0) * 30.0
    trend = gaussian_filter(bed, 10.0)
    nst = QuantileTransformer(n_quantiles=1000, output_distribution="normal", subsample=None, random_state=0).fit((bed - trend).reshape(-1, 1))
    with contextlib.redirect_stdout(io.StringIO()):
        ch = MCMC.chain_sgs(g["xx"], g["yy"], bed, g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
                            np.where(g["data_mask"] == 1, bed, np.nan), g["data_mask"], g["grounded_ice_mask"], g["resolution"])
        ch.set_update_region(True, g["highvel_mask"])
        ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
        ch.set_block_sizes(5, 20, 5, 20)
        ch.set_normal_transformation(nst, do_transform=True)
        ch.set_trend(trend, detrend_map=True)
        ch.set_variogram("Matern", 9932.5, 1.02, 0, isotropic=True, vario_smoothness=1.2259)
        ch.set_sgs_param(48, 30e3)
        ch.set_random_generator(7)
    import warnings
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = ch.run(a.iters, only_save_last_bed=True, info_per_iter=10 ** 9, plot=False, progress_bar=False)
    dt = time.perf_counter() - t0
    print(json.dumps({"mode": "sgs", "it_per_s_one_chain": a.iters / dt, "iters": a.iters, "grid": a.grid, "cores": 1, "kind": "reference",
                      "acceptance_rate": float(np.mean(out[4])),
                      "what": "chain_sgs.run (MCMC.py:1599) of the unmodified reference, one process"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["cpu_mp", "gpu", "sgs"])
    ap.add_argument("--grid", type=int, default=500)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--procs", type=int, default=1)
    a = ap.parse_args()
    # keep stdout for the JSON line only
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        {"cpu_mp": cpu_mp, "gpu": gpu, "sgs": sgs}[a.mode](a)
    real.write(buf.getvalue().strip().splitlines()[-1] + "\n")
    real.flush()


if __name__ == "__main__":
    main()
