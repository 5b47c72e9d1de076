"""A5 on the GPU (randomization-method proposal, RandField.get_random_field, MCMC.py:625-687): the summation kernel
against oracle/randmeth_oracle.py with injected modes, the device sampling against its numpy emulation, and the chain
running on this proposal draw-for-draw against the oracle step.  gstools is absent: parity with it is unpinned."""
import numpy as np
import pytest

from cases import TRAJECTORY_CASES
from gpu_helpers import oracle_setup, product_chain, quiet
from oracle import crf_oracle as O
from oracle import randmeth_oracle as R
from philox_ref import STREAM_NUGGET, box_muller, philox4x32, randmeth_modes, step_draws

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _ctx(model, nu, isotropic, pairs, res=500.0, nugget_max=0.0):
    from mcmc_gpu_b200._lib import Context
    pairs = np.asarray(pairs)
    masks = O.edge_taper_masks(pairs, (2.0, 0.0, 6.0, 1.0), 30e3, res)
    ctx = Context(int(pairs[1].max()) + 4, int(pairs[0].max()) + 4, pairs.shape[1])
    ctx.set_field_model(model, nu, isotropic, 10e3, 50e3, 8e3, 30e3, 50.0, 150.0, nugget_max)
    ctx.set_blocks(pairs, masks, res)
    return ctx, masks


@pytest.mark.parametrize("model,nu", [("Gaussian", None), ("Exponential", None), ("Matern", 0.9)])
def test_injected_modes_match_oracle_summation(model, nu):
    """Blocks smaller than, equal to and larger than one 80 x 80 register-tile pass; nugget and taper on."""
    import torch
    pairs = np.array([[10, 56, 80, 96, 112], [14, 50, 80, 112, 90]])      # row 0 widths, row 1 heights
    ctx, masks = _ctx(model, nu, False, pairs, nugget_max=4.0)
    n, n_modes = pairs.shape[1], 173
    stride = ctx.max_h * ctx.max_w
    g = np.random.default_rng(11)
    modes = np.zeros((n, n_modes, 4))
    znug = np.zeros((n, stride))
    sc, ng, rx, ry, ang, refs = [], [], [], [], [], []
    for i in range(n):
        bw, bh = int(pairs[0, i]), int(pairs[1, i])
        scale, nug, r1, r2, a = g.uniform(20, 50), g.uniform(0.5, 4.0), g.uniform(10e3, 50e3), g.uniform(8e3, 30e3), g.uniform(0, 180)
        kx, ky = R.grid_wave_vectors(model, g.random(n_modes), g.random(n_modes), r1, r2, a, nu or 1.0)
        z1, z2, zn = g.normal(size=n_modes), g.normal(size=n_modes), g.normal(size=(bh, bw))
        modes[i] = np.stack([kx, ky, z1, z2], 1)
        znug[i, :bh * bw] = zn.ravel()
        refs.append(R.randmeth_field(kx, ky, z1, z2, (bh, bw), 500.0, scale, nug, zn) * masks[i])
        sc.append(scale); ng.append(nug); rx.append(r1); ry.append(r2); ang.append(a)
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a)).to("cuda", dtype=dt)      # noqa: E731
    out = torch.full((n, stride), np.nan, dtype=torch.float64, device="cuda")
    ctx.field_randmeth(cu(np.arange(n), torch.int32), cu(sc), cu(ng), cu(rx), cu(ry), cu(ang), out, n_modes=n_modes,
                       modes=cu(modes), z_nug=cu(znug), apply_taper=True)
    got = out.cpu().numpy()
    for i in range(n):
        bw, bh = int(pairs[0, i]), int(pairs[1, i])
        f = got[i, :bh * bw].reshape(bh, bw)
        assert np.abs(f - refs[i]).max() <= TOL * np.abs(refs[i]).max(), (bh, bw)
        assert np.isnan(got[i, bh * bw:]).all()


@pytest.mark.parametrize("model,nu,isotropic", [("Gaussian", None, True), ("Exponential", None, False), ("Matern", 1.7, False)])
def test_device_sampling_matches_numpy_emulation(model, nu, isotropic):
    import torch
    from mcmc_gpu_b200 import MCMC
    pairs = np.array([[64], [72]])
    ctx, _ = _ctx(model, nu, isotropic, pairs, nugget_max=1.0)
    keys = [0x0123456789ABCDEF, 42, 2 ** 64 - 1]
    it, n_modes = 7_000_000_001, 1000
    bw, bh = 64, 72
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a)).to("cuda", dtype=dt)      # noqa: E731
    out = torch.empty((3, ctx.max_h * ctx.max_w), dtype=torch.float64, device="cuda")
    ry, ang = (2.5e4, 0.0) if isotropic else (1.1e4, 63.0)
    ctx.field_randmeth(cu([0] * 3, torch.int32), cu([30.0] * 3), cu([0.81] * 3), cu([2.5e4] * 3), cu([ry] * 3), cu([ang] * 3), out,
                       n_modes=n_modes, seeds=MCMC.keys_tensor(keys, "cuda"), iteration=it, apply_taper=False)
    got = out.cpu().numpy()
    for i, key in enumerate(keys):
        kx, ky, z1, z2 = randmeth_modes(key, it, n_modes, model, 2.5e4, ry, ang, nu or 1.0)
        e = np.arange(bh * bw, dtype=np.uint32)
        zn, _ = box_muller(philox4x32(key, e, it & 0xFFFFFFFF, it >> 32, STREAM_NUGGET))
        ref = R.randmeth_field(kx, ky, z1, z2, (bh, bw), 500.0, 30.0, 0.81, zn.reshape(bh, bw))
        assert np.abs(got[i, :bh * bw].reshape(bh, bw) - ref).max() <= TOL * np.abs(ref).max()


@pytest.mark.parametrize("isotropic", [True, False])
def test_chain_on_randmeth_proposal_matches_oracle_on_emulated_draws(isotropic):
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = dict(TRAJECTORY_CASES["ragged_rf"])
    case["rf_kw"] = dict(case["rf_kw"], isotropic=isotropic)
    n_steps, n_modes = 60, 96
    g, cs, fp = oracle_setup(case)
    ch, rf, _ = product_chain(case, g)
    rf.set_generation_method(False, n_modes=n_modes)
    key = 0xC0FFEE1234
    batch = ChainBatch(ch, rf, g["bed0"][None], [key], iter0=1)
    lc, st, bl = batch.advance(n_steps, resync_every=0)
    H, W = g["bed0"].shape
    centre_cells = np.flatnonzero(cs.region_mask.ravel() == 1) if cs.update_in_region else None
    bed = g["bed0"]
    mc_res = O.mass_conservation_residual(bed, cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution)
    loss = O.masked_loss(mc_res, cs.mc_region_mask, cs.sigma_mc)[0]
    kw = case["rf_kw"]
    acc, losses, blocks = [], [], []
    for k in range(n_steps):
        d = step_draws(key, 1 + k, fp.pairs.shape[1], fp.pairs, kw, H, W, centre_cells, randmeth=True)
        bw, bh = int(fp.pairs[0, d["pair"]]), int(fp.pairs[1, d["pair"]])
        kx, ky, z1, z2 = randmeth_modes(key, 1 + k, n_modes, kw["model_name"], d["range_x"], d["range_y"], d["angle"],
                                        kw.get("smoothness") or 1.0)
        f = R.randmeth_field(kx, ky, z1, z2, (bh, bw), cs.resolution, d["scale"], d["nug"], d["z_nug"]) * fp.edge_masks[d["pair"]]
        bed, mc_res, loss, ok, _, _ = O.crf_step(cs, bed, mc_res, loss, f, d["idx_x"], d["idx_y"], d["u"])
        acc.append(ok); losses.append(loss); blocks.append([d["idx_x"], d["idx_y"], bh, bw])
    assert np.array_equal(bl[0], np.array(blocks))
    assert np.array_equal(st[0].astype(bool), np.array(acc)), "accept/reject sequence differs"
    losses = np.array(losses)
    fin = np.isfinite(losses)
    assert (np.abs(lc[0][fin] - losses[fin]) <= TOL * np.abs(losses[fin])).all()
    assert np.abs(batch.beds()[0] - bed).max() <= TOL * np.abs(bed).max()
    assert 0.02 < np.mean(acc) < 0.99


def test_public_api_get_random_field_and_run():
    """RandField.get_random_field / get_rfblock / chain_crf.run with set_generation_method(False)."""
    from mcmc_gpu_b200 import MCMC
    case = dict(TRAJECTORY_CASES["tutorial200"])
    ch, rf, g = product_chain(case)
    rf.set_generation_method(False)
    X, Y = np.arange(0, 40 * 500.0, 500.0), np.arange(0, 30 * 500.0, 500.0)
    fields = np.stack([rf.get_random_field(X, Y) for _ in range(40)])
    assert fields.shape == (40, 30, 40) and np.isfinite(fields).all()
    # scale ~ U(scale_min, scale_max)/3: pooled standard deviation inside the implied range
    kw = case["rf_kw"]
    assert kw["scale_min"] / 3 * 0.7 < fields.std() < kw["scale_max"] / 3 * 1.3
    f = rf.get_rfblock()
    assert f.shape[0] in set(rf.pairs[1]) and f.shape[1] in set(rf.pairs[0])
    assert not (f[0].any() or f[:, -1].any())            # zero-rim taper
    ch.set_random_generator(5)
    out = quiet(ch.run, 12, rf, only_save_last_bed=True, plot=False, progress_bar=False)
    assert out[0].shape == g["bed0"].shape and out[3].shape == (12,) and np.isfinite(out[3]).all()
    assert 0 < out[4].sum() <= 11
    assert isinstance(rf, MCMC.RandField)
