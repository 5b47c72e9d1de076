#!/usr/bin/env python
"""Round 2: end-to-end step time of chain_crf.run_many with pinned host buffers for (steps in flight, chain ranges per step).
usage: r2_e2e_ab.py <steps_in_flight> <ranges> [chains] [iters] [steps]     (CUDA_DEVICE_MAX_CONNECTIONS from the environment)"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcmc_gpu_b200 import MCMC, synthetic as syn
NBUF, groups = int(sys.argv[1]), int(sys.argv[2])
C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
n_it = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 6
ON_DEVICE = os.environ.get("E2E_ON_DEVICE") == "1"      # diagnostic: "host" buffers live on the device (no PCIe traffic)
def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
dev = torch.device("cuda", 0)
H = W = 500
g, ch, rf = bench.build_chain(MCMC, syn, H, W, quiet)
seeds = [1000 + c for c in range(C)]
keys = [MCMC.philox_key(s, s) for s in seeds]
pin = (lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)) if ON_DEVICE else (lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory())
host_beds = pin((C, H, W), torch.float64)
host_beds.copy_(bench.device_initial_beds(torch, g["bed0"], 0, C, dev))
outs = [{"bed": pin((C, H, W), torch.float64), "loss": pin((C, n_it + 1), torch.float64), "steps": pin((C, n_it + 1), torch.uint8),
         "blocks": pin((C, n_it + 1, 4), torch.int32), "resampled": pin((C, H, W), torch.int32)} for _ in range(NBUF)]
batches = [MCMC.ChainBatch(ch, rf, host_beds, keys, device=dev, track_resampled=True) for _ in range(NBUF)]
def loop(n):
    pend = [None] * NBUF
    for k in range(n):
        b = k % NBUF
        if pend[b] is not None:
            pend[b].wait()
        pend[b] = batches[b].run_pipelined(host_beds, keys, n_it, outs[b], groups=groups, wait=(NBUF == 1 and False))
    for p in pend:
        if p is not None:
            p.wait()
loop(NBUF + 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); loop(steps); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
print(("[buffers on device] " if ON_DEVICE else "") + f"in_flight {NBUF} ranges {groups:3d} connections {os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS', 'default(8)'):>10s}: "
      f"{ms:8.2f} ms/step  {C * n_it / ms / 1e3:7.3f} M chain-steps/s end to end")
