"""The fused free-running kernel (K1+K4, device Philox): draw-for-draw agreement with the oracle fed the emulated draws,
invariance to batching / chunking, and two-sample agreement with reference-RNG ensembles."""
import numpy as np
import pytest

from cases import TRAJECTORY_CASES
from gpu_helpers import bits_equal, oracle_setup, product_chain, quiet
from oracle import crf_oracle as O
from philox_ref import step_draws

pytestmark = pytest.mark.gpu


def _fm(case):
    return dict(case["rf_kw"])


def _emulated_oracle_run(case, key, n_steps, g, cs, fp, iter0=1):
    """Oracle trajectory driven by the numpy emulation of the device RNG."""
    H, W = g["bed0"].shape
    centre_cells = np.flatnonzero(cs.region_mask.ravel() == 1) if cs.update_in_region else None
    bed = g["bed0"]
    mc_res = O.mass_conservation_residual(bed, cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution)
    loss = O.masked_loss(mc_res, cs.mc_region_mask, cs.sigma_mc)[0]
    acc, losses, blocks = [], [], []
    for k in range(n_steps):
        d = step_draws(key, iter0 + k, fp.pairs.shape[1], fp.pairs, _fm(case), H, W, centre_cells)
        bw, bh = int(fp.pairs[0, d["pair"]]), int(fp.pairs[1, d["pair"]])
        f = O.field_from_draws(fp, (bh, bw), d["scale"], d["nug"], d["range_x"], d["range_y"], d["z_re"], d["z_im"],
                               d["z_nug"]) * fp.edge_masks[d["pair"]]
        bed, mc_res, loss, ok, _, _ = O.crf_step(cs, bed, mc_res, loss, f, d["idx_x"], d["idx_y"], d["u"])
        acc.append(ok); losses.append(loss); blocks.append([d["idx_x"], d["idx_y"], bh, bw])
    return bed, np.array(acc), np.array(losses), np.array(blocks)


@pytest.mark.parametrize("name", ["ragged_rf", "thin_ice", "tutorial200"])
def test_free_run_matches_oracle_on_emulated_draws(name):
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = TRAJECTORY_CASES[name]
    n_steps = 60 if name == "tutorial200" else 150
    g, cs, fp = oracle_setup(case)
    ch, rf, _ = product_chain(case, g)
    key = 0x9E3779B97F4A7C15 ^ (len(name) << 40)
    batch = ChainBatch(ch, rf, g["bed0"][None], [key], iter0=1)
    lc, st, bl = batch.advance(n_steps, resync_every=0)
    bed_ref, acc_ref, loss_ref, blocks_ref = _emulated_oracle_run(case, key, n_steps, g, cs, fp)
    assert np.array_equal(bl[0], blocks_ref), "block size / centre draws differ"
    assert np.array_equal(st[0].astype(bool), acc_ref), "accept/reject sequence differs"
    fin = np.isfinite(loss_ref)
    assert (np.abs(lc[0][fin] - loss_ref[fin]) <= 1e-9 * np.abs(loss_ref[fin])).all()
    assert np.abs(batch.beds()[0] - bed_ref).max() <= 1e-9 * np.abs(bed_ref).max()
    assert 0.05 < acc_ref.mean() < 0.98


def test_batch_and_chunk_invariance_and_resync():
    """Chains are independent and counter-addressed: a chain's trajectory does not depend on which other chains share
    the launch, nor on how the iterations are chunked; the tracked loss equals a full recompute (zero-rim taper)."""
    import torch
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = dict(TRAJECTORY_CASES["tutorial200"])
    ch, rf, g = product_chain(case)
    keys = [11, 22, 33, 44, 55]
    beds0 = np.stack([g["bed0"] + k for k in range(5)])
    a = ChainBatch(ch, rf, beds0, keys, track_resampled=True)
    la, sa, ba = a.advance(90, resync_every=0)
    beds_a = a.beds()
    # (1) chains 3 and 1 alone, in two chunks, with periodic resync
    b = ChainBatch(ch, rf, beds0[[3, 1]], [keys[3], keys[1]], track_resampled=True)
    l1, s1, b1 = b.advance(40, resync_every=16)
    l2, s2, b2 = b.advance(50, resync_every=16)
    beds_b = b.beds()
    assert bits_equal(beds_b[0], beds_a[3]) and bits_equal(beds_b[1], beds_a[1])
    assert np.array_equal(np.concatenate([s1, s2], 1), sa[[3, 1]])
    assert np.array_equal(np.concatenate([b1, b2], 1), ba[[3, 1]])
    assert np.allclose(np.concatenate([l1, l2], 1), la[[3, 1]], rtol=1e-12, atol=0)
    assert np.array_equal(b.resampled_times()[0], a.resampled_times()[3])
    # (2) tracked loss == loss recomputed from the final bed (the default taper is exactly 0 on the rim)
    loss_d = torch.empty(5, dtype=torch.float64, device="cuda")
    a.ctx.residual_loss(a.bed, None, loss_d, None)
    assert np.allclose(loss_d.cpu().numpy(), la[:, -1], rtol=1e-11, atol=0)
    # (3) reference API, batched
    outs = quiet(ch.run_many, 31, rf, beds0[:2], [5, 6])
    assert len(outs) == 2 and outs[0][0].shape == g["bed0"].shape and outs[0][3].shape == (31,)
    assert np.isnan(outs[0][6][0]).all() and outs[0][4][0] == 0


def test_public_run_api_free_rng():
    case = dict(TRAJECTORY_CASES["thin_ice"])
    ch, rf, g = product_chain(case)
    ch.set_sample_points_locations(np.array([[g["xx"][3, 4], g["yy"][3, 4]], [g["xx"][50, 40], g["yy"][50, 40]]]))
    out = quiet(ch.run, 25, rf, only_save_last_bed=False, plot=False, progress_bar=False, info_per_iter=10)
    beds, loss_mc, loss_data, loss, steps, resampled, blocks, samples = out
    assert beds.shape == (25,) + g["bed0"].shape and samples.shape == (2, 25)
    assert bits_equal(beds[0], g["bed0"]) and np.array_equal(samples[:, -1], beds[-1][[3, 50], [4, 40]])
    changed = [not np.array_equal(beds[i], beds[i - 1]) for i in range(1, 25)]
    assert np.array_equal(np.array(changed), steps[1:].astype(bool))
    # continuing the chain continues the Philox counters (T3_LargeScaleChain.ipynb cell 58 pattern)
    ch.initial_bed = beds[-1]
    out2 = quiet(ch.run, 10, rf, only_save_last_bed=True, plot=False, progress_bar=False)
    assert out2[3][0] == pytest.approx(loss[-1], rel=1e-12)


def test_two_sample_agreement_with_reference_rng():
    """Free RNG: GPU (Philox) and oracle (numpy PCG64, i.e. the reference's generator) ensembles of the same chain must be
    statistically indistinguishable.  KS two-sample tests on per-chain acceptance rate and final loss, alpha = 1e-3."""
    from scipy.stats import ks_2samp
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = dict(TRAJECTORY_CASES["ragged_rf"])
    n_chains, n_iter = 96, 81
    g, cs, fp = oracle_setup(case)
    acc_o, loss_o = [], []
    for c in range(n_chains):
        r = O.run_chain(cs, fp, g["bed0"], n_iter, np.random.default_rng(5000 + c), np.random.default_rng(9000 + c))
        acc_o.append(r["steps"][1:].mean()); loss_o.append(r["loss"][-1])
    ch, rf, _ = product_chain(case, g)
    batch = ChainBatch(ch, rf, np.stack([g["bed0"]] * n_chains), [777 + c for c in range(n_chains)])
    lc, st, _ = batch.advance(n_iter - 1)
    acc_g, loss_g = st.mean(axis=1), lc[:, -1]
    p_acc = ks_2samp(acc_o, acc_g).pvalue
    p_loss = ks_2samp(loss_o, loss_g).pvalue
    assert p_acc > 1e-3 and p_loss > 1e-3, (p_acc, p_loss)


def test_driver_checkpoint_resume_and_ensemble(tmp_path):
    """largeScaleChain_mp on the GPU: reference file layout, resume continues the Philox counters (two runs of 11
    iterations == one uninterrupted chain of 20 proposals for the zero-rim taper), ensemble moments == numpy."""
    import torch
    from mcmc_gpu_b200 import MCMC, drivers
    case = dict(TRAJECTORY_CASES["tutorial200"])
    ch, rf, g = product_chain(case)
    seeds = [31, 32, 33]
    beds0 = [g["bed0"] + 0.25 * k for k in range(3)]
    r1 = quiet(drivers.largeScaleChain_mp, 3, 8, ch, rf, beds0, seeds, [11] * 3, str(tmp_path))
    r2 = quiet(drivers.largeScaleChain_mp, 3, 8, ch, rf, beds0, seeds, [11] * 3, str(tmp_path))
    batch = MCMC.ChainBatch(ch, rf, np.stack(beds0), [MCMC.philox_key(s, s) for s in seeds], track_resampled=True)
    batch.advance(10)
    lc, st, bl = batch.advance(10)
    for k in range(3):
        assert bits_equal(batch.beds()[k], r2[k][0])
        assert np.array_equal(st[k], r2[k][4][1:].astype(np.uint8))
        folder = tmp_path / "LargeScaleChain" / str(seeds[k])
        assert int(np.loadtxt(folder / "current_iter.txt")) == 22
        assert bits_equal(np.load(folder / "bed_0k.npy"), r2[k][0])
        with np.load(folder / "results_0k.npz") as z:
            assert z["loss"].shape == (22,) and np.array_equal(z["resampled_times"], r1[k][5] + r2[k][5])
    mean, var = drivers.ensemble_mean_var(batch, g["bed0"])
    stack = batch.beds()
    assert np.allclose(mean.cpu().numpy(), stack.mean(0), rtol=0, atol=1e-9)
    assert np.allclose(var.cpu().numpy(), stack.var(0), rtol=1e-9, atol=1e-12)


def test_pipelined_run_many_equals_plain_run_many():
    """The stream-pipelined end-to-end path (pinned host buffers, 4 chain ranges) returns the same bits as the plain one."""
    import torch
    from mcmc_gpu_b200 import MCMC
    case = dict(TRAJECTORY_CASES["tutorial200"])
    ch, rf, g = product_chain(case)
    C, n_iter = 6, 21
    beds0 = np.stack([g["bed0"] + 0.5 * k for k in range(C)])
    seeds = list(range(40, 40 + C))
    plain = quiet(ch.run_many, n_iter, rf, beds0, seeds, as_arrays=True, track_resampled=False)
    host = torch.as_tensor(beds0).pin_memory()
    out = {"bed": torch.empty((C,) + g["bed0"].shape, dtype=torch.float64).pin_memory(),
           "loss": torch.empty((C, n_iter), dtype=torch.float64).pin_memory(),
           "steps": torch.empty((C, n_iter), dtype=torch.uint8).pin_memory(),
           "blocks": torch.empty((C, n_iter, 4), dtype=torch.int32).pin_memory()}
    batch = MCMC.ChainBatch(ch, rf, host, [MCMC.philox_key(s, s) for s in seeds])
    res = quiet(ch.run_many, n_iter, rf, host, seeds, as_arrays=True, batch=batch, out=out, track_resampled=False)
    assert bits_equal(res["bed"].numpy(), plain["bed"])
    assert np.array_equal(res["steps"].numpy(), plain["steps"]) and np.array_equal(res["blocks"].numpy()[:, 1:], plain["blocks"][:, 1:].astype(np.int32))
    assert np.allclose(res["loss"].numpy(), plain["loss"], rtol=1e-13, atol=0)
    # round 2: full output contract (coverage counts, int32 or int16 on the link), two steps in flight on two batches
    # (wait=False handles), and the reference's list of 7-tuples from the pipelined path without running twice
    ref_b = MCMC.ChainBatch(ch, rf, beds0, [MCMC.philox_key(s, s) for s in seeds], track_resampled=True)
    ref_b.advance(n_iter - 1)
    counts = ref_b.resampled.cpu().numpy()
    batches = [MCMC.ChainBatch(ch, rf, host, [MCMC.philox_key(s, s) for s in seeds], track_resampled=True) for _ in range(2)]
    outs = []
    for dt in (torch.int32, torch.int16):
        o = {k: torch.empty_like(v).pin_memory() for k, v in out.items()}
        o["resampled"] = torch.empty((C,) + g["bed0"].shape, dtype=dt).pin_memory()
        outs.append(o)
    pend = [quiet(ch.run_many, n_iter, rf, host, seeds, as_arrays=True, batch=batches[k], out=outs[k], wait=False) for k in range(2)]
    for k in range(2):
        r = pend[k].wait()
        assert bits_equal(r["bed"].numpy(), plain["bed"]) and np.array_equal(r["steps"].numpy(), plain["steps"])
        assert np.array_equal(r["resampled"].numpy().astype(np.int32), counts)
    tuples = quiet(ch.run_many, n_iter, rf, host, seeds, batch=batches[0], out=outs[0])
    assert len(tuples) == C and bits_equal(tuples[2][0], plain["bed"][2]) and np.isnan(tuples[2][6][0]).all()
    assert np.array_equal(tuples[2][5], ref_b.resampled_times()[2]) and batches[0].bed is not None      # caller's batch stays open
    with pytest.raises(ValueError, match="wait=False"):
        ch.run_many(n_iter, rf, beds0, seeds, wait=False)


def test_more_chains_than_cta_slots_uses_chunked_scheduling_and_stays_bit_identical(monkeypatch):
    """With more chains than resident CTAs the launch hands out (chunk, chain) items dynamically and chains migrate
    between CTAs; every trajectory must equal the one-CTA-per-chain schedule bit for bit."""
    import torch
    from mcmc_gpu_b200.MCMC import ChainBatch
    case = dict(TRAJECTORY_CASES["ragged_rf"])
    ch, rf, g = product_chain(case)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    C = 2 * sm * 2 + 37                                     # more than any plausible number of resident CTAs
    beds0 = np.stack([g["bed0"] + 0.01 * (k % 11) for k in range(C)])
    keys = [1000 + 7 * k for k in range(C)]
    a = ChainBatch(ch, rf, beds0, keys, track_resampled=True)
    la, sa, ba = a.advance(57, resync_every=16)
    monkeypatch.setenv("GMC_STATIC_SCHED", "1")             # read by gmc_run
    b = ChainBatch(ch, rf, beds0, keys, track_resampled=True)
    lb, sb, bb = b.advance(57, resync_every=16)
    assert bits_equal(a.beds(), b.beds())
    assert np.array_equal(sa, sb) and np.array_equal(ba, bb) and bits_equal(la, lb)
    assert np.array_equal(a.resampled_times(), b.resampled_times())
    assert 0.05 < sa.mean() < 0.98


def test_block_larger_than_the_grid_raises_like_the_reference():
    """A block cut by both edges makes the reference's slices differ in length (numpy broadcast ValueError)."""
    from mcmc_gpu_b200 import MCMC
    case = dict(TRAJECTORY_CASES["ragged_rf"])
    case["blocks"] = (60, 100, 12, 26)               # wider than the 90-column grid
    ch, rf, g = product_chain(case)
    with pytest.raises(ValueError, match="broadcast"):
        MCMC.ChainBatch(ch, rf, g["bed0"][None], [1])


def test_concurrent_chunk_scheduled_launches_do_not_share_scheduler_state(monkeypatch):
    """run_pipelined issues one launch per chain range on its own stream; with more chains per range than resident CTAs
    every launch is chunk-scheduled and they run concurrently - each must use its own scheduler area."""
    import torch
    from mcmc_gpu_b200 import MCMC
    case = dict(TRAJECTORY_CASES["ragged_rf"])
    ch, rf, g = product_chain(case)
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    C, n_iter = 3 * (2 * sm + 9), 19                          # three ranges, each above the CTA slots
    beds0 = np.stack([g["bed0"] + 0.01 * (k % 13) for k in range(C)])
    seeds = list(range(700, 700 + C))
    keys = [MCMC.philox_key(s, s) for s in seeds]
    host = torch.as_tensor(beds0).pin_memory()
    out = {"bed": torch.empty((C,) + g["bed0"].shape, dtype=torch.float64).pin_memory(),
           "loss": torch.empty((C, n_iter), dtype=torch.float64).pin_memory(),
           "steps": torch.empty((C, n_iter), dtype=torch.uint8).pin_memory(),
           "blocks": torch.empty((C, n_iter, 4), dtype=torch.int32).pin_memory()}
    batch = MCMC.ChainBatch(ch, rf, host, keys)
    res = batch.run_pipelined(host, keys, n_iter - 1, out, groups=3)
    piped = {k: v.numpy().copy() for k, v in res.items()}
    monkeypatch.setenv("GMC_STATIC_SCHED", "1")
    ref = MCMC.ChainBatch(ch, rf, beds0, keys)
    lr, sr, br = ref.advance(n_iter - 1)
    assert bits_equal(piped["bed"], ref.beds())
    assert np.array_equal(piped["steps"][:, 1:], sr) and np.array_equal(piped["blocks"][:, 1:], br)
    assert bits_equal(piped["loss"][:, 1:], lr)
