#!/usr/bin/env python
"""chain-steps/s of the large-scale chain on the randomization-method proposal (A5, set_generation_method(False)).
usage (on a GPU box): python profiles/randmeth_bench.py [chains] [iters] [n_modes]"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mcmc_gpu_b200 import MCMC, synthetic as syn

C = int(sys.argv[1]) if len(sys.argv) > 1 else 296
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n_modes = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
g = syn.make_grids(500, 500)
kw = syn.RF_KW
with contextlib.redirect_stdout(io.StringIO()):
    rf = MCMC.RandField(kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"], kw["scale_max"],
                        kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*syn.BLOCKS); rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0); rf.set_generation_method(False, n_modes)
    ch = MCMC.chain_crf(g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"], g["data_mask"],
                        g["grounded_ice_mask"], 500.0)
    ch.set_update_region(True, g["highvel_mask"]); ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    ch.set_update_type("CRF_weight"); ch.set_crf_data_weight(rf)
batch = MCMC.ChainBatch(ch, rf, syn.chain_initial_beds(g["bed0"], C), [MCMC.philox_key(s, s) for s in range(C)])
batch.advance(n_it, want_caches=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); lc, st, bl = batch.advance(n_it); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"randomization method, {n_modes} modes, {C} chains x {n_it} iterations: {ms:.1f} ms, {C*n_it/ms:.1f} k chain-steps/s, "
      f"acceptance {st.mean():.3f}")
