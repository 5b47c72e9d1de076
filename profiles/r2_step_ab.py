#!/usr/bin/env python
"""Round 2: device-resident chain-steps/s of the fused step kernel for one (chains, grid) shape.
usage: r2_step_ab.py <chains> <grid> [iters] [reps]      (GMC_STEP_WIDE=0/1 forces the 256 / 512-thread CTA)"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcmc_gpu_b200 import MCMC, synthetic as syn
C, N = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 200
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
dev = torch.device("cuda", 0)
g, ch, rf = bench.build_chain(MCMC, syn, N, N, quiet)
beds = bench.device_initial_beds(torch, g["bed0"], 0, C, dev)
batch = MCMC.ChainBatch(ch, rf, beds, [MCMC.philox_key(1000 + c, 1000 + c) for c in range(C)], device=dev, track_resampled=True)
for _ in range(2):
    batch.advance(iters, want_caches=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    batch.advance(iters, want_caches=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
batch.ctx.check()
print(f"chains {C:5d} grid {N:4d} iters {iters:4d} wide={os.environ.get('GMC_STEP_WIDE', 'auto'):4s} "
      f"threads {batch.ctx.step_kernel_info(C)['threads']}: {ms:9.3f} ms/launch  {C * iters / ms / 1e3:8.3f} M chain-steps/s")
