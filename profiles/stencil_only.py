#!/usr/bin/env python
"""Time (and give ncu something small to profile) the full-grid stencil kernels: U1 residual, U2 residual+loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_gpu_b200 import synthetic as syn
from mcmc_gpu_b200._lib import Context
C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 500
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
g = syn.make_grids(N, N)
ctx = Context(N, N, C)
ctx.set_static(g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["highvel_mask"], g["highvel_mask"], None, None, 500.0, 5.0)
bed = torch.as_tensor(syn.chain_initial_beds(g["bed0"], min(C, 8))).cuda().repeat((C + 7) // 8, 1, 1)[:C].contiguous()
res = torch.empty_like(bed); loss = torch.empty(C, dtype=torch.float64, device="cuda")
def t(fn):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cells = C * N * N
for name, fn, b in (("U1 residual (8 B read + 8 B write / cell)", lambda: ctx.residual(bed, res), 16),
                    ("U2 residual+loss (8 B read / cell)", lambda: ctx.residual_loss(bed, None, loss, None), 8),
                    ("residual+loss+write", lambda: ctx.residual_loss(bed, res, loss, None), 16),
                    ("loss only (8 B read / cell)", lambda: ctx.loss(res, loss), 8),
                    ("torch copy (8 B read + 8 B write / cell)", lambda: res.copy_(bed), 16)):
    ms = t(fn)
    print(f"{name:45s} {ms:8.4f} ms  {cells * b / ms / 1e6:8.1f} GB/s  ({cells * b / ms / 1e6 / 6545.6 * 100:5.1f}% of measured 6545.6 GB/s)")
# the loss-only variant evaluates the residual as a linear form of the bed: its loss against the exact (bit-identical) residual's
ctx.residual_loss(bed, None, loss, None); l_lin = loss.cpu().numpy().copy()
ctx.residual(bed, res); ctx.loss(res, loss); l_exact = loss.cpu().numpy()
print(f"loss-only (linear form) vs loss of the exact residual: max relative difference {np.abs(l_lin / l_exact - 1).max():.2e}")
