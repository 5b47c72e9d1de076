#!/usr/bin/env python
"""Throughput of the small-scale SGS chain (BASELINE.json config 4: 512 chains, 300x300, blocks 5-19, 48 neighbours,
30 km radius, Matern nu=1.2259) on one B200, next to the oracle port on the host cores.
usage: python profiles/sgs_bench.py [chains] [iters] [cpu_iters]"""
import contextlib, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cpu_it = int(sys.argv[3]) if len(sys.argv) > 3 else 0
CASE = dict(H=300, W=300, n_iter=n_it, seed=1, sigma_mc=5.0, blocks=(5, 20, 5, 20), neighbors=48, radius=30e3,
            vario=dict(vtype="Matern", range=9932.5, sill=1.02, nugget=0.0, isotropic=True, smoothness=1.2259, azimuth=None),
            transform=True, detrend=True, n_quantiles=1000)


def _cpu_chain(args):
    seed, n = args
    from oracle import sgs_oracle as S
    from sgs_helpers import oracle_sgs_setup
    g, su = oracle_sgs_setup(CASE)
    t0 = time.perf_counter()
    out = S.sgs_chain_run(su, g["bed_init"], n, np.random.default_rng(seed))
    return time.perf_counter() - t0, float(np.nansum(out["blocks"][:, 2] * out["blocks"][:, 3]))


def main():
    import torch
    from mcmc_gpu_b200 import MCMC
    from sgs_helpers import product_sgs_chain
    with contextlib.redirect_stdout(io.StringIO()):
        ch, g = product_sgs_chain(CASE)
    beds = np.stack([g["bed_init"]] * C)
    batch = MCMC.SgsBatch(ch, beds, [MCMC.philox_key(s) for s in range(C)])
    batch.advance(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lc, st, bl = batch.advance(n_it)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    nodes = float((bl[..., 2].astype(np.float64) * bl[..., 3]).sum())
    res = {"workload": "small-scale SGS chain, %d chains, 300x300, blocks 5-19, 48 neighbours, radius 30 km" % C,
           "chain_steps_per_s": C * n_it / (ms * 1e-3), "kriged_nodes_per_s_upper": nodes / (ms * 1e-3), "ms": ms,
           "acceptance": float(st.mean()), "iters": n_it}
    if os.environ.get("GMC_PHASES"):
        batch.ctx.phase_timing(True)
        batch.advance(n_it)
        cyc = batch.ctx.phase_timing(True, read=True)
        names = ["reset+path", "octant search", "compact", "assemble Sigma", "eliminate", "weights+draw (per node)", "inverse+residual+MH"]
        res["phase_cycles_pct"] = {n: round(100.0 * c / max(cyc.sum(), 1), 1) for n, c in zip(names, cyc)}
        res["cycles_per_node"] = float(cyc.sum()) / nodes
    if cpu_it:
        import multiprocessing as mp
        cores = os.cpu_count()
        with mp.get_context("spawn").Pool(cores) as pool:
            t0 = time.perf_counter()
            r = pool.map(_cpu_chain, [(100 + c, cpu_it) for c in range(cores)])
            wall = time.perf_counter() - t0
        res["cpu_port"] = {"cores": cores, "chain_steps_per_s": cores * cpu_it / max(x[0] for x in r), "wall_s_incl_startup": wall,
                           "sample": "%d chains x %d iterations, oracle/sgs_oracle.py, one process per core" % (cores, cpu_it)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
