"""K4 parity: replaying the oracle's recorded proposals through gmc_step_injected must reproduce the reference
trajectory — accept/reject flags, bed, tracked residual and resampled_times bit-for-bit, losses within 1e-9."""
import os

import numpy as np
import pytest

from cases import TRAJECTORY_CASES
from gpu_helpers import bits_equal, oracle_setup, product_chain, quiet, same_values
from oracle import crf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _oracle_run(case, n_iter=None):
    g, cs, fp = oracle_setup(case)
    out = O.run_chain(cs, fp, g["bed0"], n_iter or case["n_iter"], np.random.default_rng(case["chain_seed"]),
                      np.random.default_rng(case["rf_seed"]), record=True)
    return g, cs, fp, out


@pytest.mark.parametrize("name", sorted(TRAJECTORY_CASES))
def test_replay_matches_reference_trajectory(name):
    case = TRAJECTORY_CASES[name]
    gold = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
    g, cs, fp, ora = _oracle_run(case)
    ch, rf, _ = product_chain(case, g)
    assert bits_equal(ch.crf_data_weight, gold["crf_weight"])
    assert bits_equal(rf.edge_masks[0], gold["edge_mask0"]) and bits_equal(rf.edge_masks[-1], gold["edge_mask_last"])
    out = quiet(ch.run, case["n_iter"], rf, only_save_last_bed=True, plot=False, progress_bar=False, info_per_iter=10 ** 9,
                replay=ora["tape"])
    bed, loss_mc, loss_data, loss, steps, resampled, blocks = out
    assert np.array_equal(steps, gold["steps"]), "accept/reject sequence differs from the reference"
    assert bits_equal(bed, gold["bed"]), "final bed differs from the reference"
    assert same_values(resampled, gold["resampled_times"])
    assert same_values(blocks, gold["blocks"])
    finite = np.isfinite(gold["loss"])
    assert np.array_equal(finite, np.isfinite(loss))
    rel = np.abs(loss[finite] - gold["loss"][finite]) / np.maximum(np.abs(gold["loss"][finite]), 1e-300)
    assert rel.max() <= 1e-9, rel.max()
    assert np.array_equal(loss_mc, loss) and not loss_data.any()


def test_replay_many_chains_at_once_and_tracked_residual():
    """Three different chains advance in one launch each step; state must equal three independent oracle runs."""
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = dict(TRAJECTORY_CASES["ragged_rf"], n_iter=120)
    runs = []
    for k in range(3):
        c = dict(case, chain_seed=100 + k, rf_seed=200 + k)
        runs.append(_oracle_run(c))
    g = runs[0][0]
    ch, rf, _ = product_chain(case, g)
    beds0 = np.stack([g["bed0"]] * 3)
    batch = ChainBatch(ch, rf, beds0, [1, 2, 3], track_resampled=True)
    for i in range(case["n_iter"] - 1):
        tp = [r[3]["tape"][i] for r in runs]
        acc, loss = batch.step_injected([t["f"] for t in tp], [(t["idx_x"], t["idx_y"]) for t in tp], [t["u"] for t in tp])
        for k in range(3):
            assert acc[k] == bool(runs[k][3]["steps"][i + 1])
    beds, res = batch.beds(), batch.residuals()
    for k in range(3):
        assert bits_equal(beds[k], runs[k][3]["bed"])
        assert same_values(res[k], runs[k][3]["mc_res"])          # includes the stale-ring quirk of this taper
        assert same_values(batch.resampled_times()[k], runs[k][3]["resampled_times"])


def test_loss_next_inf_on_thickness_violation():
    """A proposal that lifts the bed above the surface must be rejected with loss_next = inf (MCMC.py:1321-1329)."""
    import torch
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = TRAJECTORY_CASES["thin_ice"]
    g, cs, fp = oracle_setup(case)
    ch, rf, _ = product_chain(case, g)
    batch = ChainBatch(ch, rf, g["bed0"][None], [9])
    f = np.full((16, 16), 1e4) * rf.edge_masks[0][:16, :16].clip(0, 1)
    f[8, 8] = 1e4
    centre = np.argwhere(g["highvel_mask"] == 1)[len(np.argwhere(g["highvel_mask"] == 1)) // 2]
    dev = batch.dev
    acc = torch.empty(1, dtype=torch.uint8, device=dev)
    loss = torch.empty(1, dtype=torch.float64, device=dev)
    lnext = torch.empty(1, dtype=torch.float64, device=dev)
    before = batch.beds().copy()
    batch.ctx.step_injected(batch.bed, batch.mcres, batch.ssq, torch.as_tensor(f.reshape(1, -1)).to(dev),
                            torch.tensor([[16, 16]], dtype=torch.int32, device=dev),
                            torch.tensor([[int(centre[0]), int(centre[1])]], dtype=torch.int32, device=dev),
                            torch.tensor([0.5], dtype=torch.float64, device=dev), 16, 16, acc, loss, lnext)
    assert acc.item() == 0 and np.isinf(lnext.item())
    assert bits_equal(batch.beds(), before)


def test_run_rejects_non_randfield():
    case = TRAJECTORY_CASES["ragged_rf"]
    ch, rf, _ = product_chain(case)
    with pytest.raises(TypeError):
        ch.run(10, object(), plot=False, progress_bar=False)
