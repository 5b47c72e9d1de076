"""K1 parity: spectral synthesis with every random input injected vs the reference's golden fields (<= 1e-9 of max|f|),
plus sanity of the device Philox normals."""
import os

import numpy as np
import pytest

from cases import FIELD_CASES
from gpu_helpers import quiet
from oracle import crf_oracle as O
from philox_ref import STREAM_NOISE, box_muller, hermitian_noise, philox4x32

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-9      # north_star: "within 1e-9 relative in fp64"


def _rf(fc):
    from mcmc_gpu_b200 import MCMC
    kw = fc["rf_kw"]
    return quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
                 kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw.get("smoothness"),
                 rng_seed=fc["seed"])


@pytest.mark.parametrize("name", sorted(FIELD_CASES))
def test_injected_field_matches_reference(name):
    from mcmc_gpu_b200 import MCMC
    fc = FIELD_CASES[name]
    gold = np.load(os.path.join(GOLD, "spectral_fields.npz"))
    fp = O.FieldParams(**fc["rf_kw"], resolution=fc["res"])
    rng = np.random.default_rng(fc["seed"])
    rf = _rf(fc)
    for k, shape in enumerate(fc["shapes"]):
        draws = O.draw_field_inputs(fp, rng, tuple(shape))
        got = MCMC.spectral_synthesis_field(rf, tuple(shape), res=fc["res"], draws=draws)
        ref = gold[f"{name}__{k}"]
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= TOL, (name, shape, err)


def test_tapered_block_matches_oracle_for_all_default_pairs():
    """gmc_field_spectral with apply_taper over all 25 tutorial block sizes (covers radix 2,4,5,7,8,9 stages)."""
    import torch
    from mcmc_gpu_b200 import synthetic as syn
    from mcmc_gpu_b200._lib import Context
    from mcmc_gpu_b200 import MCMC
    rf = quiet(MCMC.RandField, *[syn.RF_KW[k] for k in ("range_min_x", "range_max_x", "range_min_y", "range_max_y", "scale_min",
                                                         "scale_max", "nugget_max", "model_name", "isotropic")],
               smoothness=syn.RF_KW["smoothness"], rng_seed=1)
    rf.set_block_sizes(*syn.BLOCKS)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0)
    fp = O.FieldParams(**syn.RF_KW, resolution=500.0, pairs=rf.pairs)
    fp.edge_masks = O.edge_taper_masks(fp.pairs, syn.LOGISTIC, syn.MAX_DIST, 500.0)
    n = rf.pairs.shape[1]
    ctx = Context(100, 100, n)
    ctx.set_field_model(rf.model_name, rf.smoothness, rf.isotropic, rf.range_min_x, rf.range_max_x, rf.range_min_y,
                        rf.range_max_y, rf.scale_min, rf.scale_max, rf.nugget_max)
    ctx.set_blocks(rf.pairs, rf.edge_masks, rf.resolution)
    stride = ctx.max_h * ctx.max_w
    rng = np.random.default_rng(5)
    zre, zim, znug = (np.zeros((n, stride)) for _ in range(3))
    sc, ng, rx, ry, refs = [], [], [], [], []
    for i in range(n):
        bw, bh = int(rf.pairs[0, i]), int(rf.pairs[1, i])
        d = O.draw_field_inputs(fp, rng, (bh, bw))
        refs.append(O.field_from_draws(fp, (bh, bw), **d) * fp.edge_masks[i])
        zre[i, :bh * bw], zim[i, :bh * bw], znug[i, :bh * bw] = d["z_re"].ravel(), d["z_im"].ravel(), d["z_nug"].ravel()
        sc.append(d["scale"]); ng.append(d["nug"]); rx.append(d["range_x"]); ry.append(d["range_y"])
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a)).to("cuda", dtype=dt)      # noqa: E731
    out = torch.empty((n, stride), dtype=torch.float64, device="cuda")
    ctx.field_spectral(cu(np.arange(n), torch.int32), cu(sc), cu(ng), cu(rx), cu(ry), out, z_re=cu(zre), z_im=cu(zim),
                       z_nug=cu(znug), apply_taper=True)
    got = out.cpu().numpy()
    for i in range(n):
        bw, bh = int(rf.pairs[0, i]), int(rf.pairs[1, i])
        err = np.abs(got[i, :bh * bw].reshape(bh, bw) - refs[i]).max() / np.abs(refs[i]).max()
        assert err <= TOL, (bh, bw, err)
        assert bits_zero_rim(got[i, :bh * bw].reshape(bh, bw))


def bits_zero_rim(f):
    """With logistic (2,0,6,1) the taper is exactly 0 on the rim."""
    return not (f[0].any() or f[-1].any() or f[:, 0].any() or f[:, -1].any())


def test_device_normals_match_numpy_emulation_and_are_standard():
    """Free-RNG field with scale fixed: compare against the oracle fed with the emulated Philox normals."""
    import torch
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    from mcmc_gpu_b200._lib import Context
    rf = quiet(MCMC.RandField, *[syn.RF_KW[k] for k in ("range_min_x", "range_max_x", "range_min_y", "range_max_y", "scale_min",
                                                         "scale_max", "nugget_max", "model_name", "isotropic")],
               smoothness=syn.RF_KW["smoothness"], rng_seed=1)
    rf.set_block_sizes(64, 64, 80, 80, steps=1)
    rf.set_weight_param(*syn.LOGISTIC, syn.MAX_DIST, 500.0)
    fp = O.FieldParams(**syn.RF_KW, resolution=500.0, pairs=rf.pairs)
    ctx = Context(100, 100, 4)
    ctx.set_field_model(rf.model_name, rf.smoothness, rf.isotropic, rf.range_min_x, rf.range_max_x, rf.range_min_y,
                        rf.range_max_y, rf.scale_min, rf.scale_max, rf.nugget_max)
    ctx.set_blocks(rf.pairs, rf.edge_masks, rf.resolution)
    bw, bh = int(rf.pairs[0, 0]), int(rf.pairs[1, 0])
    keys = [0x0123456789ABCDEF, 42, 2 ** 64 - 1, 7]
    it = 5_000_000_123
    out = torch.empty((4, bh * bw), dtype=torch.float64, device="cuda")
    cu = lambda a, dt=torch.float64: torch.as_tensor(np.asarray(a)).to("cuda", dtype=dt)      # noqa: E731
    ctx.field_spectral(cu([0] * 4, torch.int32), cu([30.0] * 4), cu([0.0] * 4), cu([2e4] * 4), cu([2e4] * 4), out,
                       seeds=MCMC.keys_tensor(keys, "cuda"), iteration=it, apply_taper=False)
    got = out.cpu().numpy()
    allz = []
    for i, key in enumerate(keys):
        e = np.arange(bh * bw, dtype=np.uint32)
        allz += list(box_muller(philox4x32(key, e, it & 0xFFFFFFFF, it >> 32, STREAM_NOISE)))
        zr, zi = hermitian_noise(key, it, bh, bw)
        ref = O.field_from_draws(fp, (bh, bw), 30.0, 0.0, 2e4, 2e4, zr, zi, np.zeros((bh, bw)))
        err = np.abs(got[i].reshape(bh, bw) - ref).max() / np.abs(ref).max()
        assert err <= TOL, err
    z = np.concatenate(allz)
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 6 * np.sqrt(2 / z.size)


def test_bad_generation_method_arguments_are_loud():
    """set_generation_method(False) selects the randomization method (tests/test_gpu_randmeth.py); nonsense mode counts
    are rejected by the library instead of being clamped."""
    from mcmc_gpu_b200 import MCMC, synthetic as syn
    rf = quiet(MCMC.RandField, 1e3, 2e3, 1e3, 2e3, 1, 2, 0, "Gaussian", True)
    rf.set_block_sizes(10, 12, 10, 12, steps=2)
    rf.set_weight_param(*syn.LOGISTIC, 1e3, 100.0)
    rf.set_generation_method(False, n_modes=0)
    with pytest.raises(ValueError):
        rf.get_rfblock()
    rf.set_generation_method(False, n_modes=8)
    assert rf.get_rfblock().shape in {(10, 10), (10, 12), (12, 10), (12, 12)}
