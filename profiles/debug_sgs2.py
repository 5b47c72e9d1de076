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
H, W = g["bed_init"].shape
ora = S.sgs_chain_run(su, g["bed_init"], 1, np.random.default_rng(case["seed"]), record=True)
t = ora["tape"][0]
ch, _ = product_sgs_chain(case, g)
batch = MCMC.SgsBatch(ch, g["bed_init"][None], [1])
z0 = batch.z.cpu().numpy()[0].copy()
trend = su.trend if su.trend is not None else 0.0
bed_c = g["bed_init"] - trend
z_ref0 = su.nst.forward(bed_c.reshape(-1)).reshape(H, W) if su.nst is not None else bed_c
print("init z max diff", np.abs(z0 - z_ref0).max(), "mcres/ssq loss gpu", batch.loss()[0], "oracle", ora["loss"][0] if ora["steps"][0] == 0 else "n/a")
x0, x1 = max(0, int(t["idx_x"] - t["bsx"] / 2)), min(H, int(t["idx_x"] + t["bsx"] / 2))
y0, y1 = max(0, int(t["idx_y"] - t["bsy"] / 2)), min(W, int(t["idx_y"] + t["bsy"] / 2))
# oracle newsim for this block
z_cond = su.nst.forward((su.cond_bed - trend).reshape(-1)).reshape(H, W) if su.nst is not None else (su.cond_bed - trend)
tosim = z_ref0.copy(); tosim[x0:x1, y0:y1] = z_cond[x0:x1, y0:y1]
mask = np.zeros((H, W), bool); mask[x0:x1, y0:y1] = True
newsim = S.sgs_block(su.xx, su.yy, tosim, su.vario, su.radius, su.num_points, mask, None, None, t)
dev = batch.dev
acc = torch.empty(1, dtype=torch.uint8, device=dev); loss = torch.empty(1, dtype=torch.float64, device=dev); ln = torch.empty(1, dtype=torch.float64, device=dev)
p = np.asarray(t["path"]); path = ((p[:, 0] - x0) * (y1 - y0) + (p[:, 1] - y0)).astype(np.int32)[None]
zn = np.nan_to_num(np.asarray(t["z"]))[None]
cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
batch.ctx.sgs_step_injected(batch.bedc, batch.z, batch.mcres, batch.ssq, batch.nviol, cu(np.array([[t["idx_x"], t["idx_y"]]], dtype=np.int32)),
                            cu(np.array([[t["bsx"], t["bsy"]]], dtype=np.int32)), cu(path), cu(zn), cu(np.array([0.0])), acc, loss, ln, batch.resampled, batch.err)
zg = batch.z.cpu().numpy()[0]
print("accepted", acc.item(), "err", batch.err.item(), "block", (x0, x1, y0, y1))
np.set_printoptions(precision=6, linewidth=200)
print("oracle newsim block\n", newsim[x0:x1, y0:y1])
print("gpu z block\n", zg[x0:x1, y0:y1])
print("path (block-local)", path[0], "\nzn", zn[0])
# per-node check of the first simulated node: neighbours from oracle
i, j = p[0]
cond = ~np.isnan(tosim)
hw = S.stencil_half_width(su.xx[0, :], su.radius)
nb = S.octant_neighbors(i, j, su.xx, su.yy, tosim, cond, su.radius, su.num_points, hw)
print("first node", (i, j), "cond?", cond[i, j], "oracle neighbours (i,j,val):\n", nb[:, [3, 4, 2]])
est, var = S.ok_solve((su.xx[i, j], su.yy[i, j]), nb, su.vario)
print("oracle est,var", est, var)
