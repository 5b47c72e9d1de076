#!/usr/bin/env python
"""Round 2: per-range timeline of one steady-state end-to-end step (GMC_TRACE_PIPELINE events of ChainBatch.run_pipelined).
usage: r2_e2e_trace.py <steps_in_flight> <ranges>"""
import contextlib, io, os, sys
os.environ["GMC_TRACE_PIPELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from mcmc_gpu_b200 import MCMC, synthetic as syn
NBUF, groups = int(sys.argv[1]), int(sys.argv[2])
C, n_it, H, W = 256, 1000, 500, 500
def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
dev = torch.device("cuda", 0)
g, ch, rf = bench.build_chain(MCMC, syn, H, W, quiet)
seeds = [1000 + c for c in range(C)]
keys = [MCMC.philox_key(s, s) for s in seeds]
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
host_beds = pin((C, H, W), torch.float64)
host_beds.copy_(bench.device_initial_beds(torch, g["bed0"], 0, C, dev).cpu())
outs = [{"bed": pin((C, H, W), torch.float64), "loss": pin((C, n_it + 1), torch.float64), "steps": pin((C, n_it + 1), torch.uint8),
         "blocks": pin((C, n_it + 1, 4), torch.int32), "resampled": pin((C, H, W), torch.int32)} for _ in range(NBUF)]
batches = [MCMC.ChainBatch(ch, rf, host_beds, keys, device=dev, track_resampled=True) for _ in range(NBUF)]
t0 = torch.cuda.Event(enable_timing=True)
pend, traces = [None] * NBUF, []
n = 3 * NBUF + 2
for k in range(n):
    b = k % NBUF
    if pend[b] is not None:
        pend[b].wait()
    if k == NBUF:
        t0.record()
    pend[b] = ch.run_many(n_it + 1, rf, host_beds, seeds, as_arrays=True, batch=batches[b], out=outs[b], wait=False, pipeline_groups=groups)
    traces.append(list(batches[b]._trace))
for p in pend:
    if p is not None:
        p.wait()
torch.cuda.synchronize()
names = ["upload queued", "kernel queued", "kernel done", "download done"]
for k in range(NBUF, n):
    tr = traces[k]
    row = {}
    for gi, what, ev in tr:
        row.setdefault(what, []).append(t0.elapsed_time(ev))
    print(f"step {k}: " + "  ".join(f"{names[w]} {min(v):7.1f}..{max(v):7.1f}" for w, v in sorted(row.items())))
