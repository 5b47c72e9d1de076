"""K6 parity: the small-scale (SGS) chain on the GPU vs the oracle (stable tie order, the kernel's rule): normal-score
transform, replayed trajectories (accept flags identical, loss/bed <= 1e-9), batch invariance of the free-running kernel."""
import os

import numpy as np
import pytest

from cases import SGS_CASES
from gpu_helpers import bits_equal, quiet
from oracle import sgs_oracle as S
from sgs_helpers import oracle_sgs_setup, product_sgs_chain

pytestmark = pytest.mark.gpu
TOL = 1e-9


def test_normal_score_transform_matches_sklearn_restatement():
    import torch
    case = SGS_CASES["matern_nst"]
    g, su = oracle_sgs_setup(case)
    ch, _ = product_sgs_chain(case, g)
    ctx = ch._sgs_context(1)
    x = (g["bed_init"] - g["trend"]).reshape(-1)
    x = np.concatenate([x, [x.min() - 5.0, x.max() + 5.0, np.nan, g["quantiles"][0], g["quantiles"][-1], g["quantiles"][7]]])
    xd = torch.as_tensor(x).cuda()
    zd = torch.empty_like(xd)
    ctx.sgs_transform(xd, zd, inverse=False)
    z_ref = su.nst.forward(x)
    z = zd.cpu().numpy()
    assert np.array_equal(np.isnan(z), np.isnan(z_ref))
    assert np.nanmax(np.abs(z - z_ref)) <= 1e-11
    zz = np.concatenate([z_ref, [-9.0, 9.0, 0.0, 5.3, -5.3]])
    bd = torch.empty(zz.size, dtype=torch.float64, device="cuda")
    ctx.sgs_transform(torch.as_tensor(zz).cuda(), bd, inverse=True)
    b_ref = su.nst.inverse(zz)
    b = bd.cpu().numpy()
    assert np.array_equal(np.isnan(b), np.isnan(b_ref))
    assert np.nanmax(np.abs(b - b_ref) / np.maximum(np.abs(b_ref), 1.0)) <= 1e-11


@pytest.mark.parametrize("solver", ["warp", "cta"])
@pytest.mark.parametrize("name", sorted(SGS_CASES))
def test_replay_matches_oracle_trajectory(name, solver, monkeypatch):
    """Both kriging solvers: warp-per-node (num_points <= 48, the default) and the CTA-wide one it falls back to."""
    if solver == "cta":
        monkeypatch.setenv("GMC_SGS_SOLVER", "cta")          # read by gmc_sgs_setup
    case = SGS_CASES[name]
    g, su = oracle_sgs_setup(case)
    ora = S.sgs_chain_run(su, g["bed_init"], case["n_iter"], np.random.default_rng(case["seed"]), record=True)
    ch, _ = product_sgs_chain(case, g)
    out = quiet(ch.run, case["n_iter"], only_save_last_bed=True, plot=False, progress_bar=False, replay=ora["tape"])
    bed, loss_mc, loss_data, loss, steps, resampled, blocks = out
    assert np.array_equal(steps, ora["steps"]), "accept/reject sequence differs from the oracle"
    assert np.array_equal(blocks, ora["blocks"]) and np.array_equal(resampled, ora["resampled_times"])
    fin = np.isfinite(ora["loss"])
    assert np.array_equal(fin, np.isfinite(loss))
    assert (np.abs(loss[fin] - ora["loss"][fin]) <= TOL * np.abs(ora["loss"][fin])).all()
    assert np.abs(bed - ora["bed"]).max() <= TOL * np.abs(ora["bed"]).max()
    assert 0.1 < steps.mean() < 0.95
    n_inf = sum(1 for t in ora["tape"] if np.isinf(t["loss_next"]))
    assert (n_inf > 0) == bool(case.get("thin")), "the thin-ice case must exercise the thickness guard"


def test_free_run_is_batch_invariant_and_sane():
    from mcmc_gpu_b200 import MCMC
    case = SGS_CASES["matern_nst"]
    ch, g = product_sgs_chain(case)
    beds0 = np.stack([g["bed_init"] + 0.1 * k for k in range(4)])
    keys = [MCMC.philox_key(s) for s in (5, 6, 7, 8)]
    a = MCMC.SgsBatch(ch, beds0, keys)
    la, sa, ba = a.advance(25)
    b = MCMC.SgsBatch(ch, beds0[[2, 0]], [keys[2], keys[0]])
    l1, s1, b1 = b.advance(10)
    l2, s2, b2 = b.advance(15)
    assert bits_equal(b.beds()[0], a.beds()[2]) and bits_equal(b.beds()[1], a.beds()[0])
    assert np.array_equal(np.concatenate([s1, s2], 1), sa[[2, 0]]) and np.array_equal(np.concatenate([b1, b2], 1), ba[[2, 0]])
    assert np.isfinite(la).all() and 0.05 < sa.mean() < 0.99
    bx, by = ba[..., 2], ba[..., 3]
    assert bx.min() >= case["blocks"][0] and bx.max() < case["blocks"][1] and by.min() >= case["blocks"][2] and by.max() < case["blocks"][3]
    # the tracked loss equals a full recompute from the final bed (the SGS chain has no stale ring)
    import torch
    full = torch.as_tensor(a.beds(with_trend=True)).cuda()
    loss = torch.empty(4, dtype=torch.float64, device="cuda")
    a.ctx.residual_loss(full, None, loss, None)
    assert np.allclose(loss.cpu().numpy(), la[:, -1], rtol=1e-9, atol=0)
    # pinned host tensors in and out (the end-to-end path of bench.py's sgs.e2e) give the same beds
    pinned_in = torch.as_tensor(beds0).pin_memory()
    pinned_out = torch.empty_like(pinned_in).pin_memory()
    c = MCMC.SgsBatch(ch, pinned_in.numpy(), keys)
    c.advance(25)
    assert bits_equal(c.beds(out=pinned_out), a.beds())


def test_public_run_api_free_rng():
    case = SGS_CASES["expo_raw"]
    ch, g = product_sgs_chain(case)
    out = quiet(ch.run, 12, only_save_last_bed=False, plot=False, progress_bar=False)
    beds, loss_mc, loss_data, loss, steps, resampled, blocks = out
    assert beds.shape == (12,) + g["bed_init"].shape and loss.shape == (12,) and blocks.shape == (12, 4)
    changed = [not np.array_equal(beds[i], beds[i - 1]) for i in range(1, 12)]
    assert np.array_equal(np.array(changed), steps[1:].astype(bool))


def test_small_scale_driver_layout_and_resume(tmp_path):
    """smallScaleChain_mp on the GPU: the reference's folder layout; two runs of 6 == one uninterrupted run of 12."""
    from mcmc_gpu_b200 import MCMC, drivers
    case = SGS_CASES["matern_nst"]
    ch, g = product_sgs_chain(case)
    seeds, lsc = [901, 902], 777
    beds0 = [g["bed_init"], g["bed_init"] + 0.2]
    r1 = quiet(drivers.smallScaleChain_mp, 2, 4, ch, beds0, seeds, lsc, [6, 6], str(tmp_path))
    r2 = quiet(drivers.smallScaleChain_mp, 2, 4, ch, beds0, seeds, lsc, [6, 6], str(tmp_path))
    folder = tmp_path / "LargeScaleChain" / "777" / "SmallScaleChain" / "901"
    assert sorted(p.name for p in folder.iterdir()) == ["RNGState_RandField.txt", "RNGState_chain.txt", "bed_0k.npy",
                                                        "current_iter.txt", "results_0k.npz"]
    assert int(np.loadtxt(folder / "current_iter.txt")) == 12
    batch = MCMC.SgsBatch(ch, np.stack(beds0), [MCMC.philox_key(s) for s in seeds])
    batch.advance(12)
    final = batch.beds(with_trend=True)
    for k in range(2):
        # resume re-derives bedc = bed - trend from the saved full bed: one rounding of (x + t) - t per cell
        assert np.abs(final[k] - r2[k][0]).max() <= 1e-9 * np.abs(final[k]).max()
        assert len(r1[k]) == 7 and r1[k][3].shape == (6,)


def test_more_chains_than_resident_ctas_is_bit_identical_to_the_static_schedule(monkeypatch):
    """sgs_run_kernel hands out (chunk, chain) items dynamically when the launch has more chains than resident CTAs."""
    import torch
    from mcmc_gpu_b200 import MCMC
    case = SGS_CASES["matern_nst"]
    ch, g = product_sgs_chain(case)
    C = 2 * torch.cuda.get_device_properties(0).multi_processor_count + 5
    beds0 = np.stack([g["bed_init"] + 0.01 * (k % 7) for k in range(C)])
    keys = [MCMC.philox_key(500 + k) for k in range(C)]
    a = MCMC.SgsBatch(ch, beds0, keys)
    la, sa, ba = a.advance(9)
    monkeypatch.setenv("GMC_STATIC_SCHED", "1")
    b = MCMC.SgsBatch(ch, beds0, keys)
    lb, sb, bb = b.advance(9)
    assert bits_equal(a.beds(), b.beds()) and np.array_equal(sa, sb) and np.array_equal(ba, bb) and bits_equal(la, lb)
    assert np.isfinite(la).all()


def test_search_radius_is_widened_in_the_chain_like_the_reference():
    """MCMC.py:149-155: a node that finds no conditioned cell within the radius searches again 100 km wider.  With a radius
    of 3 cells and blocks of up to 13 cells the first nodes of most paths see nothing at level 0; the replayed oracle
    trajectory (which widens exactly like the reference) must be reproduced, with both solvers."""
    case = dict(H=40, W=44, n_iter=14, seed=5, sigma_mc=1.5, blocks=(8, 14, 8, 14), neighbors=16, radius=1.5e3,
                vario=dict(vtype="Matern", range=3000.0, sill=1.0, nugget=0.0, isotropic=True, smoothness=1.2259, azimuth=None),
                transform=True, detrend=True, n_quantiles=300)
    g, su = oracle_sgs_setup(case)
    ora = S.sgs_chain_run(su, g["bed_init"], case["n_iter"], np.random.default_rng(case["seed"]), record=True)
    for solver in ("warp", "cta"):
        if solver == "cta":
            os.environ["GMC_SGS_SOLVER"] = "cta"
        try:
            ch, _ = product_sgs_chain(case, g)
            assert ch._sgs_context(1)._h is not None and ch._sgs_levels          # the widened tables were built
            ch._ctx = None
            out = quiet(ch.run, case["n_iter"], only_save_last_bed=True, plot=False, progress_bar=False, replay=ora["tape"])
        finally:
            os.environ.pop("GMC_SGS_SOLVER", None)
        bed, _, _, loss, steps, resampled, blocks = out
        assert np.array_equal(steps, ora["steps"]) and np.array_equal(blocks, ora["blocks"])
        fin = np.isfinite(ora["loss"])
        assert (np.abs(loss[fin] - ora["loss"][fin]) <= TOL * np.abs(ora["loss"][fin])).all()
        assert np.abs(bed - ora["bed"]).max() <= TOL * np.abs(ora["bed"]).max()


def test_out_of_range_cells_are_reported():
    """cells outside the transformer's fitted range are where the block-local round trip and the reference's
    whole-grid round trip part ways; the batch constructor says so instead of diverging silently."""
    from mcmc_gpu_b200 import MCMC
    case = SGS_CASES["matern_nst"]
    ch, g = product_sgs_chain(case)
    bed = g["bed_init"].copy()
    import warnings
    with warnings.catch_warnings(record=True) as rec:                # in range: no warning
        warnings.simplefilter("always")
        MCMC.SgsBatch(ch, bed[None], [1])
    assert not [w for w in rec if issubclass(w.category, RuntimeWarning)]
    bed[3, 4] += 1e4                                                 # far above the largest fitted quantile
    with pytest.warns(RuntimeWarning, match="outside the normal-score transformer's fitted range"):
        MCMC.SgsBatch(ch, bed[None], [1])
