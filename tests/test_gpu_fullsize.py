"""Parity at BASELINE.json's full sizes (500x500, 2000x2000, SGS 300x300 with 48 neighbours) through size-independent
properties: the stencil against the numpy oracle (bit-exact), the tracked loss against a full recompute, invariance of a
chain's trajectory to the batch it runs in and to chunking, and one oracle step replayed at full size."""
import numpy as np
import pytest

from gpu_helpers import bits_equal, quiet, same_values
from oracle import crf_oracle as O

pytestmark = pytest.mark.gpu


def _tutorial_chain(N):
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    g = syn.make_grids(N, N)
    kw = syn.RF_KW
    rf = quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
               kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*syn.BLOCKS)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0)
    rf.set_generation_method(True)
    ch = quiet(MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"],
               g["data_mask"], g["grounded_ice_mask"], 500.0)
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    quiet(ch.set_update_type, "CRF_weight")
    if N > 600:                      # the conditioning weight is a one-time O(N^2 x data) setup: keep the test short
        ch.crf_data_weight = np.ones((N, N))
    else:
        ch.set_crf_data_weight(rf)
    return ch, rf, g


@pytest.mark.parametrize("N,C,iters", [(500, 6, 250), (2000, 2, 120)])
def test_large_scale_chain_at_full_size(N, C, iters):
    import torch
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    ch, rf, g = _tutorial_chain(N)
    beds0 = syn.chain_initial_beds(g["bed0"], C)
    keys = [MCMC.philox_key(1000 + c) for c in range(C)]
    a = MCMC.ChainBatch(ch, rf, beds0, keys)
    # (1) the initial residual / loss of every chain equal the oracle's, bit for bit / to 1e-12
    res0 = a.mcres.cpu().numpy()
    mask = np.asarray(g["highvel_mask"])
    for c in range(C):
        ref = O.mass_conservation_residual(beds0[c], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], 500.0)
        assert same_values(res0[c], ref)
        ref_loss = O.masked_loss(ref, mask, syn.SIGMA_MC)[0]
        assert abs(float(a.ssq[c]) / (2 * syn.SIGMA_MC ** 2) - ref_loss) <= 1e-12 * ref_loss
    la, sa, ba = a.advance(iters, resync_every=0)
    # (with the unit weight of the 2000x2000 case nearly every proposal raises the loss: a handful of accepts is expected)
    assert (0.05 if N <= 600 else 0.0) < sa.mean() < 0.95 and np.isfinite(la).all()
    # (2) tracked loss == full recompute from the final beds (zero-rim taper: no stale ring)
    loss = torch.empty(C, dtype=torch.float64, device="cuda")
    a.ctx.residual_loss(a.bed, None, loss, None)
    assert np.allclose(loss.cpu().numpy(), la[:, -1], rtol=1e-9, atol=0)
    # (3) the tracked residual equals the oracle's residual of the final bed, bit for bit
    final = a.beds()
    ref = O.mass_conservation_residual(final[C - 1], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], 500.0)
    assert same_values(a.mcres[C - 1].cpu().numpy(), ref)
    # (4) the last chain alone, in two chunks: the same trajectory bit for bit
    b = MCMC.ChainBatch(ch, rf, beds0[C - 1:], keys[C - 1:])
    _, s1, b1 = b.advance(iters // 3, resync_every=0)
    _, s2, b2 = b.advance(iters - iters // 3, resync_every=0)
    assert bits_equal(b.beds()[0], final[C - 1])
    assert np.array_equal(np.concatenate([s1, s2], 1)[0], sa[C - 1]) and np.array_equal(np.concatenate([b1, b2], 1)[0], ba[C - 1])
    # (5) blocks stay inside the drawn sizes and the update region
    assert set(np.unique(ba[..., 2])) <= set(rf.pairs[1]) and set(np.unique(ba[..., 3])) <= set(rf.pairs[0])
    assert mask[ba[..., 0], ba[..., 1]].all()


def test_one_replayed_oracle_step_at_500():
    """A full-size oracle step (numpy, 500x500) against gmc_step_injected fed the same field, centre and uniform."""
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    ch, rf, g = _tutorial_chain(500)
    cs, fp = O.setup_from_grids(g, sigma_mc=syn.SIGMA_MC, logistic=syn.LOGISTIC, max_dist=syn.MAX_DIST, blocks=syn.BLOCKS)
    rng = np.random.default_rng(3)
    bed = g["bed0"].copy()
    mc = O.mass_conservation_residual(bed, cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution)
    loss = O.masked_loss(mc, cs.mc_region_mask, cs.sigma_mc)[0]
    batch = MCMC.ChainBatch(ch, rf, bed[None], [1])
    cells = np.argwhere(np.asarray(g["highvel_mask"]) == 1)
    for k in range(6):
        pick = int(rng.integers(fp.pairs.shape[1]))
        bw, bh = int(fp.pairs[0, pick]), int(fp.pairs[1, pick])
        d = O.draw_field_inputs(fp, rng, (bh, bw))
        f = O.field_from_draws(fp, (bh, bw), **d) * fp.edge_masks[pick]
        ix, iy = (int(x) for x in cells[rng.integers(len(cells))]) if k else (3, 497)       # first block clipped by two edges
        u = float(rng.random())
        bed, mc, loss, ok, _, _ = O.crf_step(cs, bed, mc, loss, f, ix, iy, u)
        acc, l_gpu = batch.step_injected([f], [(ix, iy)], [u])
        assert bool(acc[0]) == bool(ok)
        assert bits_equal(batch.beds()[0], bed) and same_values(batch.mcres[0].cpu().numpy(), mc)
        if np.isfinite(loss):
            assert abs(l_gpu[0] - loss) <= 1e-9 * abs(loss)


def test_sgs_chain_at_config4_shape():
    """300x300, blocks 5-19, 48 neighbours, 30 km radius: batch invariance (bit-identical) and loss == recompute."""
    import torch
    from mcmc_gpu_b200 import MCMC
    from sgs_helpers import product_sgs_chain
    case = dict(H=300, W=300, n_iter=12, seed=1, sigma_mc=5.0, blocks=(5, 20, 5, 20), neighbors=48, radius=30e3,
                vario=dict(vtype="Matern", range=9932.5, sill=1.02, nugget=0.0, isotropic=True, smoothness=1.2259, azimuth=None),
                transform=True, detrend=True, n_quantiles=1000)
    ch, g = product_sgs_chain(case)
    beds0 = np.stack([g["bed_init"] + 0.05 * k for k in range(3)])
    keys = [MCMC.philox_key(s) for s in (21, 22, 23)]
    a = MCMC.SgsBatch(ch, beds0, keys)
    la, sa, ba = a.advance(12)
    b = MCMC.SgsBatch(ch, beds0[[2]], [keys[2]])
    _, s1, _ = b.advance(5)
    _, s2, _ = b.advance(7)
    assert bits_equal(b.beds()[0], a.beds()[2]) and np.array_equal(np.concatenate([s1, s2], 1)[0], sa[2])
    full = torch.as_tensor(a.beds(with_trend=True)).cuda()
    loss = torch.empty(3, dtype=torch.float64, device="cuda")
    a.ctx.residual_loss(full, None, loss, None)
    assert np.allclose(loss.cpu().numpy(), la[:, -1], rtol=1e-9, atol=0)
    assert np.isfinite(la).all() and 0.0 < sa.mean() <= 1.0


def test_sgs_config4_shape_replays_oracle_steps():
    """the SGS chain at the config-4 shape (300x300, 48 neighbours, 30 km) against the ORACLE, not only
    against itself: five recorded oracle proposals replayed through gmc_sgs_step_injected."""
    from oracle import sgs_oracle as S
    from sgs_helpers import oracle_sgs_setup, product_sgs_chain
    case = dict(H=300, W=300, n_iter=5, seed=1, sigma_mc=5.0, blocks=(5, 20, 5, 20), neighbors=48, radius=30e3,
                vario=dict(vtype="Matern", range=9932.5, sill=1.02, nugget=0.0, isotropic=True, smoothness=1.2259, azimuth=None),
                transform=True, detrend=True, n_quantiles=1000)
    g, su = oracle_sgs_setup(case)
    ora = S.sgs_chain_run(su, g["bed_init"], case["n_iter"], np.random.default_rng(case["seed"]), record=True)
    ch, _ = product_sgs_chain(case, g)
    bed, _, _, loss, steps, resampled, blocks = quiet(ch.run, case["n_iter"], only_save_last_bed=True, plot=False,
                                                      progress_bar=False, replay=ora["tape"])
    assert np.array_equal(steps, ora["steps"]) and np.array_equal(blocks, ora["blocks"])
    assert np.array_equal(resampled, ora["resampled_times"])
    fin = np.isfinite(ora["loss"])
    assert (np.abs(loss[fin] - ora["loss"][fin]) <= 1e-9 * np.abs(ora["loss"][fin])).all()
    assert np.abs(bed - ora["bed"]).max() <= 1e-9 * np.abs(ora["bed"]).max()


def test_free_run_at_500_matches_the_emulated_oracle():
    """150 free-running steps at 500x500 against the oracle fed the numpy emulation of the device RNG
    (the same check tests/test_gpu_run.py makes at <= 200x200): identical block draws and accept sequence, bed <= 1e-9."""
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    from test_gpu_run import _emulated_oracle_run
    ch, rf, g = _tutorial_chain(500)
    cs, fp = O.setup_from_grids(g, sigma_mc=syn.SIGMA_MC, logistic=syn.LOGISTIC, max_dist=syn.MAX_DIST, blocks=syn.BLOCKS)
    case = dict(rf_kw=dict(syn.RF_KW))
    key = MCMC.philox_key(4242)
    n_steps = 150
    batch = MCMC.ChainBatch(ch, rf, g["bed0"][None], [key], iter0=1, track_resampled=True)
    lc, st, bl = batch.advance(n_steps, resync_every=0)
    bed_ref, acc_ref, loss_ref, blocks_ref = _emulated_oracle_run(case, key, n_steps, g, cs, fp)
    assert np.array_equal(bl[0], blocks_ref) and np.array_equal(st[0].astype(bool), acc_ref)
    fin = np.isfinite(loss_ref)
    assert (np.abs(lc[0][fin] - loss_ref[fin]) <= 1e-9 * np.abs(loss_ref[fin])).all()
    assert np.abs(batch.beds()[0] - bed_ref).max() <= 1e-9 * np.abs(bed_ref).max()
    assert 0.05 < acc_ref.mean() < 0.98
    # the same chain inside a launch wider than the SM count (256-thread CTAs instead of the 512-thread ones): bit-identical
    wide = MCMC.ChainBatch(ch, rf, np.stack([g["bed0"]] * 150), [key] + [MCMC.philox_key(s) for s in range(149)], iter0=1)
    _, st2, bl2 = wide.advance(n_steps, resync_every=0)
    assert np.array_equal(st2[0], st[0]) and np.array_equal(bl2[0], bl[0]) and bits_equal(wide.beds()[0], batch.beds()[0])
    # ... and in split mode (a field-producer CTA and a Metropolis-tail CTA per chain, pipelined across steps): bit-identical
    split = MCMC.ChainBatch(ch, rf, np.stack([g["bed0"]] * 3), [key, MCMC.philox_key(7), MCMC.philox_key(8)], iter0=1,
                            track_resampled=True)
    split.step_cta = "split"
    _, st3, bl3 = split.advance(n_steps, resync_every=0)
    assert split.ctx.step_kernel_info(3).get("ctas_per_chain") == 2
    assert np.array_equal(st3[0], st[0]) and np.array_equal(bl3[0], bl[0]) and bits_equal(split.beds()[0], batch.beds()[0])
    assert np.array_equal(split.resampled_times()[0], batch.resampled_times()[0])


def test_2000_grid_real_weight_and_replayed_oracle_steps():
    """the 2000x2000 configuration with the REAL conditioning weight (GPU nearest-data distance vs the
    oracle's KD-tree, bit for bit) and three oracle steps replayed through gmc_step_injected, the first block clipped by
    the top-right corner of the grid."""
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    N = 2000
    g = syn.make_grids(N, N)
    kw = syn.RF_KW
    rf = quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
               kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw["smoothness"], rng_seed=0)
    rf.set_block_sizes(*syn.BLOCKS)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0)
    rf.set_generation_method(True)
    ch = quiet(MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"],
               g["data_mask"], g["grounded_ice_mask"], 500.0)
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=syn.SIGMA_MC, massConvInRegion=True)
    quiet(ch.set_update_type, "CRF_weight")
    ch.set_crf_data_weight(rf)                                    # min_dist_kernel on 4 M cells x 80 k data points
    cs, fp = O.setup_from_grids(g, sigma_mc=syn.SIGMA_MC, logistic=syn.LOGISTIC, max_dist=syn.MAX_DIST, blocks=syn.BLOCKS)
    assert bits_equal(ch.crf_data_weight, cs.crf_weight)
    rng = np.random.default_rng(8)
    bed = g["bed0"].copy()
    mc = O.mass_conservation_residual(bed, cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution)
    loss = O.masked_loss(mc, cs.mc_region_mask, cs.sigma_mc)[0]
    batch = MCMC.ChainBatch(ch, rf, bed[None], [1], track_resampled=True)
    assert same_values(batch.mcres[0].cpu().numpy(), mc)
    cells = np.argwhere(np.asarray(g["highvel_mask"]) == 1)
    for k in range(3):
        pick = int(rng.integers(fp.pairs.shape[1]))
        bw, bh = int(fp.pairs[0, pick]), int(fp.pairs[1, pick])
        d = O.draw_field_inputs(fp, rng, (bh, bw))
        f = O.field_from_draws(fp, (bh, bw), **d) * fp.edge_masks[pick]
        ix, iy = (int(x) for x in cells[rng.integers(len(cells))]) if k else (2, N - 3)       # first block: corner-clipped
        u = float(rng.random()) * (0.2 if k == 2 else 1.0)
        bed, mc, loss, ok, _, _ = O.crf_step(cs, bed, mc, loss, f, ix, iy, u)
        acc, l_gpu = batch.step_injected([f], [(ix, iy)], [u])
        assert bool(acc[0]) == bool(ok)
        assert bits_equal(batch.beds()[0], bed) and same_values(batch.mcres[0].cpu().numpy(), mc)
        if np.isfinite(loss):
            assert abs(l_gpu[0] - loss) <= 1e-9 * abs(loss)
