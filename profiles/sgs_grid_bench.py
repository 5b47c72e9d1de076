#!/usr/bin/env python
"""Whole-grid SGS realisations (the initial beds of the large-scale chains, gstatsim_custom/interpolate.sgs): GPU time for
n_real realisations of an N x N grid next to the oracle port's node rate on one host core.
usage: python profiles/sgs_grid_bench.py [N] [n_real] [radius_m] [num_points] [cpu_nodes]"""
import json, os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

N = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n_real = int(sys.argv[2]) if len(sys.argv) > 2 else 8
radius = float(sys.argv[3]) if len(sys.argv) > 3 else 50e3
k = int(sys.argv[4]) if len(sys.argv) > 4 else 48
cpu_nodes = int(sys.argv[5]) if len(sys.argv) > 5 else 0


def main():
    import torch
    from mcmc_gpu_b200 import synthetic as syn
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    g = syn.make_grids(N, N)
    r = np.random.default_rng(0)
    lines = np.zeros((N, N), dtype=bool)                 # radar-like flight lines: every 25th row / column + 1 % scatter
    lines[::25, :] = True
    lines[:, ::40] = True
    lines |= r.random((N, N)) < 0.01
    cond = np.where(lines, g["bed0"], np.nan)
    vario = dict(azimuth=0.0, nugget=0.0, major_range=9932.5, minor_range=9932.5, sill=1.02, s=1.2259, vtype="matern")
    bounds = (np.full((N, N), -9999.0), g["surf"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        interpolate.sgs_many(g["xx"], g["yy"], cond, vario, [1], radius=radius, num_points=k, bounds=bounds)      # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sims = interpolate.sgs_many(g["xx"], g["yy"], cond, vario, list(range(100, 100 + n_real)), radius=radius, num_points=k,
                                    bounds=bounds, as_tensor=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    nodes = int((~lines).sum())
    res = {"workload": f"whole-grid SGS, {N}x{N}, {100 * lines.mean():.1f} % conditioning data, radius {radius / 1e3:.0f} km, "
                       f"{k} neighbours, bounds (-9999, surface), Matern nu=1.2259",
           "realisations": n_real, "nodes_per_realisation": nodes, "wall_s_incl_host_setup": wall,
           "s_per_realisation": wall / n_real, "kriged_nodes_per_s": nodes * n_real / wall,
           "max_above_surface_m": float((sims - torch.as_tensor(g["surf"]).cuda()).max())}
    if cpu_nodes:
        from oracle import sgs_oracle as S
        sub = 120                                        # a corner of the same problem, sized for seconds
        sl = (slice(0, sub), slice(0, sub))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            S.sgs_grid(g["xx"][sl], g["yy"][sl], cond[sl], vario, radius, k, np.random.default_rng(1), bounds=(bounds[0][sl], bounds[1][sl]))
            dt = time.perf_counter() - t0
        n_sub = int((~lines[sl]).sum())
        res["cpu_port"] = {"cores": 1, "kriged_nodes_per_s": n_sub / dt, "sample": f"{sub}x{sub} corner, {n_sub} nodes, oracle/sgs_oracle.py",
                           "s_per_realisation_extrapolated": nodes / (n_sub / dt)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
