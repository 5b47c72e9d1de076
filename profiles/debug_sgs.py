import os, sys
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")]
import numpy as np, torch
from cases import SGS_CASES
from oracle import sgs_oracle as S
from sgs_helpers import oracle_sgs_setup, product_sgs_chain
from mcmc_gpu_b200 import MCMC
name = sys.argv[1] if len(sys.argv) > 1 else "matern_nst"
case = SGS_CASES[name]
g, su = oracle_sgs_setup(case)
ora = S.sgs_chain_run(su, g["bed_init"], case["n_iter"], np.random.default_rng(case["seed"]), record=True)
ch, _ = product_sgs_chain(case, g)
batch = MCMC.SgsBatch(ch, g["bed_init"][None], [1])
# oracle step-by-step states by replay
H, W = g["bed_init"].shape
trend = su.trend if su.trend is not None else 0.0
for it, t in enumerate(ora["tape"]):
    dev = batch.dev
    acc = torch.empty(1, dtype=torch.uint8, device=dev); loss = torch.empty(1, dtype=torch.float64, device=dev); ln = torch.empty(1, dtype=torch.float64, device=dev)
    x0, x1 = max(0, int(t["idx_x"] - t["bsx"] / 2)), min(H, int(t["idx_x"] + t["bsx"] / 2))
    y0, y1 = max(0, int(t["idx_y"] - t["bsy"] / 2)), min(W, int(t["idx_y"] + t["bsy"] / 2))
    p = np.asarray(t["path"]); path = ((p[:, 0] - x0) * (y1 - y0) + (p[:, 1] - y0)).astype(np.int32)[None]
    zn = np.nan_to_num(np.asarray(t["z"]))[None]
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    batch.ctx.sgs_step_injected(batch.bedc, batch.z, batch.mcres, batch.ssq, batch.nviol, cu(np.array([[t["idx_x"], t["idx_y"]]], dtype=np.int32)),
                                cu(np.array([[t["bsx"], t["bsy"]]], dtype=np.int32)), cu(path), cu(zn), cu(np.array([t["u"]])), acc, loss, ln, batch.resampled, batch.err)
    gl = ln.item(); ol = t["loss_next"]
    rel = abs(gl - ol) / abs(ol) if np.isfinite(ol) else (0.0 if gl == ol else np.inf)
    flag = "" if (bool(acc.item()) == bool(ora["steps"][it])) else "  <-- ACCEPT MISMATCH"
    if rel > 1e-10 or flag:
        print(f"it {it}: block ({x0}:{x1},{y0}:{y1}) n={len(path[0])} loss_next gpu {gl!r} oracle {ol!r} rel {rel:.3e} nviol {batch.nviol.item()}{flag}")
        if flag or rel > 1e-6:
            break
print("done", it)
