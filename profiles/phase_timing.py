#!/usr/bin/env python
"""Per-phase SM-cycle breakdown of the fused step kernel (debug counters, thread 0 of each CTA).
usage (on a GPU box): python profiles/phase_timing.py [chains] [iters]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mcmc_gpu_b200 import MCMC, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 200
g = syn.make_grids(500, 500)
BLOCKS = tuple(int(x) for x in os.environ.get("GMC_BLOCKS", ",".join(map(str, syn.BLOCKS))).split(","))
kw = syn.RF_KW
with contextlib.redirect_stdout(io.StringIO()):
    rf = MCMC.RandField(kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"], kw["scale_max"],
                        kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*BLOCKS); rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0); rf.set_generation_method(True)
    ch = MCMC.chain_crf(g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"], g["data_mask"],
                        g["grounded_ice_mask"], 500.0)
    ch.set_update_region(True, g["highvel_mask"]); ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    ch.set_update_type("CRF_weight"); ch.set_crf_data_weight(rf)
batch = MCMC.ChainBatch(ch, rf, syn.chain_initial_beds(g["bed0"], C), [MCMC.philox_key(s, s) for s in range(C)])
batch.advance(n_it, want_caches=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); batch.advance(n_it, want_caches=False); e1.record(); torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1)
batch.ctx.phase_timing(True)
e0.record(); batch.advance(n_it, want_caches=False); e1.record(); torch.cuda.synchronize()
ms_timed = e0.elapsed_time(e1)
cyc = batch.ctx.phase_timing(True, read=True)
names = ["scalars+prefetch", "spectrum fill (RNG, sqrt S, power)", "column DFT", "row recombination", "row DFT", "candidate tile",
         "block residual+loss+decision", "write-back"]
tot = cyc.sum()
print(batch.ctx.step_kernel_info(), "blocks", BLOCKS)
print(f"chains {C} iters {n_it}: {ms_plain:.2f} ms plain, {ms_timed:.2f} ms with counters; {C*n_it/ms_plain/1e3:.3f} M chain-steps/s")
print(f"cycles per chain-step (CTA wall, thread 0): {tot/(C*n_it):.0f}")
for n, c in zip(names, cyc):
    print(f"  {n:36s} {c/(C*n_it):9.0f} cyc  {100*c/tot:5.1f}%")
