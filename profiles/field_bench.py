#!/usr/bin/env python
"""Stand-alone proposal-field kernel (K1, gmc_field_spectral, device Philox) and masked-loss kernel (K3, gmc_loss):
fields/s, cells/s and bytes written resp. read per second.  usage: python profiles/field_bench.py [n_fields] [reps]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_gpu_b200 import MCMC, synthetic as syn
from mcmc_gpu_b200._lib import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
kw = syn.RF_KW
with contextlib.redirect_stdout(io.StringIO()):
    rf = MCMC.RandField(kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"], kw["scale_max"],
                        kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
rf.set_block_sizes(*syn.BLOCKS); rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0)
ctx = Context(100, 100, n)
ctx.set_field_model(rf.model_name, rf.smoothness, rf.isotropic, rf.range_min_x, rf.range_max_x, rf.range_min_y, rf.range_max_y,
                    rf.scale_min, rf.scale_max, rf.nugget_max)
ctx.set_blocks(rf.pairs, rf.edge_masks, rf.resolution)
g = np.random.default_rng(0)
pick = g.integers(0, rf.pairs.shape[1], n)
cells = float((rf.pairs[0][pick] * rf.pairs[1][pick]).sum())
cu = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a)).to("cuda", dtype=dt)      # noqa: E731
out = torch.empty((n, ctx.max_h * ctx.max_w), dtype=torch.float64, device="cuda")
args = (cu(pick, torch.int32), cu(g.uniform(17, 50, n)), cu(np.zeros(n)), cu(g.uniform(10e3, 50e3, n)), cu(g.uniform(10e3, 50e3, n)), out)
seeds = MCMC.keys_tensor(list(range(n)), "cuda")


def t(fn):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = t(lambda: ctx.field_spectral(*args, seeds=seeds, iteration=7, apply_taper=True))
print(f"field_kernel: {n} tapered fields (blocks 50-80, Matern, device Philox) in {ms:.3f} ms = {n / ms / 1e3:.3f} M fields/s, "
      f"{cells / ms / 1e6:.2f} G cells/s, {cells * 8 / ms / 1e6:.1f} GB/s written ({cells * 8 / ms / 1e6 / 6545.6 * 100:.1f} % of HBM peak: "
      "compute bound, ~1 MFLOP of FP64 per field)")
C, N = 256, 500
gr = syn.make_grids(N, N)
c2 = Context(N, N, C)
c2.set_static(gr["surf"], gr["velx"], gr["vely"], gr["dhdt"], gr["smb"], gr["highvel_mask"], gr["highvel_mask"], None, None, 500.0, 5.0)
res = torch.randn((C, N, N), dtype=torch.float64, device="cuda")
loss = torch.empty(C, dtype=torch.float64, device="cuda")
ms = t(lambda: c2.loss(res, loss))
print(f"loss_kernel + finalize: {C} x {N}x{N} residuals in {ms:.4f} ms = {C * N * N * 8 / ms / 1e6:.1f} GB/s read "
      f"({C * N * N * 8 / ms / 1e6 / 6545.6 * 100:.1f} % of the measured 6545.6 GB/s)")
