"""K2/K3 parity: batched residual and masked loss vs the oracle (bit-exact residual, loss <= 1e-12 rel)."""
import os

import numpy as np
import pytest

from cases import residual_case_inputs
from gpu_helpers import same_values
from oracle import crf_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _rand_inputs(H, W, seed, nan_frac=0.0):
    g = np.random.default_rng(seed)
    surf = 1500.0 + 300.0 * g.standard_normal((H, W))
    d = dict(surf=surf, velx=200.0 * g.standard_normal((H, W)), vely=150.0 * g.standard_normal((H, W)),
             dhdt=g.standard_normal((H, W)), smb=g.standard_normal((H, W)))
    if nan_frac:
        d["velx"][g.random((H, W)) < nan_frac] = np.nan
    return d, g


def test_topography_api_matches_reference_golden():
    from mcmc_gpu_b200 import Topography
    ri = residual_case_inputs()
    gold = np.load(os.path.join(GOLD, "residual_loss.npz"))
    res = Topography.get_mass_conservation_residual(ri["bed"], ri["surf"], ri["velx"], ri["vely"], ri["dhdt"], ri["smb"],
                                                    ri["resolution"])
    assert same_values(res, gold["residual"])


def test_topography_tensor_api():
    import torch
    from mcmc_gpu_b200 import Topography
    ri = residual_case_inputs()
    t = {k: torch.as_tensor(ri[k]).cuda() for k in ("bed", "surf", "velx", "vely", "dhdt", "smb")}
    res = Topography.get_mass_conservation_residual_tensor(t["bed"], t["surf"], t["velx"], t["vely"], t["dhdt"], t["smb"],
                                                           torch.tensor(ri["resolution"]))
    assert res.is_cuda and res.dtype == torch.float64
    gold = np.load(os.path.join(GOLD, "residual_loss.npz"))
    assert same_values(res.cpu().numpy(), gold["residual"])


@pytest.mark.parametrize("H,W,C", [(2, 2, 1), (2, 9, 3), (33, 2, 2), (37, 53, 4), (64, 128, 2), (200, 200, 3), (131, 257, 5),
                                   (70, 64, 2), (45, 100, 3), (40, 500, 2), (19, 66, 9)])
def test_batched_residual_and_loss(H, W, C):
    import torch
    from mcmc_gpu_b200._lib import Context
    st, g = _rand_inputs(H, W, 7 + H * W, nan_frac=0.01 if H * W > 100 else 0.0)
    beds = st["surf"][None] - 800.0 + 100.0 * g.standard_normal((C, H, W))
    if H * W > 100:
        beds[0, H // 2, W // 3] = np.nan
    mask = (g.random((H, W)) < 0.6).astype(np.uint8)
    sigma, res_m = 3.5, 250.0
    ctx = Context(H, W, C)
    ctx.set_static(st["surf"], st["velx"], st["vely"], st["dhdt"], st["smb"], np.ones((H, W)), mask, None, None, res_m, sigma)
    bed_d = torch.as_tensor(beds).cuda()
    res_d = torch.empty_like(bed_d)
    loss_d = torch.empty(C, dtype=torch.float64, device="cuda")
    ssq_d = torch.empty_like(loss_d)
    ctx.residual(bed_d, res_d)
    res = res_d.cpu().numpy()
    ctx.residual_loss(bed_d, None, loss_d, ssq_d)
    loss_fused = loss_d.cpu().numpy()
    ctx.loss(res_d, loss_d)
    loss_sep = loss_d.cpu().numpy()
    for c in range(C):
        ref = O.mass_conservation_residual(beds[c], st["surf"], st["velx"], st["vely"], st["dhdt"], st["smb"], res_m)
        assert same_values(res[c], ref), f"chain {c}"
        ref_loss = O.masked_loss(ref, mask, sigma)[0]
        tol = 1e-12 * max(abs(ref_loss), 1e-300)
        assert abs(loss_fused[c] - ref_loss) <= tol and abs(loss_sep[c] - ref_loss) <= tol
    assert np.allclose(ssq_d.cpu().numpy() / (2 * sigma ** 2), loss_fused, rtol=1e-15)


def test_loss_only_linear_form_agrees_with_the_exact_residual():
    """The loss-only stencil variant evaluates the residual as a 5-point linear form of the bed (csrc/stencil_tma.cuh,
    R2Lin); the variants that write the residual keep the flux form.  Same loss to 1e-12 (contract 1e-9) over a wide
    dynamic range of velocities, nan statics, nan / inf bed cells, a sparse loss mask and every edge rule."""
    import torch
    from mcmc_gpu_b200._lib import Context
    H, W, C = 97, 192, 12
    g = np.random.default_rng(11)
    surf = 1500.0 + 300.0 * g.standard_normal((H, W))
    velx = g.standard_normal((H, W)) * 10.0 ** g.uniform(-3, 3, (H, W))
    vely = g.standard_normal((H, W)) * 10.0 ** g.uniform(-3, 3, (H, W))
    velx[g.random((H, W)) < 0.02] = np.nan
    vely[0, :7] = 0.0
    dhdt, smb = g.standard_normal((H, W)), g.standard_normal((H, W))
    mask = (g.random((H, W)) < 0.35).astype(np.uint8)
    mask[0, :], mask[-1, :], mask[:, 0], mask[:, -1] = 1, 1, 1, 1           # the one-sided rows / columns count
    mask[50, 99], velx[50, 98:103] = 1, 3.0                                   # a counted neighbour of the infinite bed cell below
    beds = surf[None] - 900.0 + 200.0 * g.standard_normal((C, H, W))
    beds[1, 40, 60] = np.nan
    beds[2, 0, 0] = beds[2, H - 1, W - 1] = np.nan
    beds[3, 50, 100] = np.inf                                                 # an infinite residual is an infinite loss
    sigma, res_m = 2.0, 500.0
    ctx = Context(H, W, C)
    ctx.set_static(surf, velx, vely, dhdt, smb, np.ones((H, W)), mask, None, None, res_m, sigma)
    assert "tma" in ctx.stencil_kernel_name()
    bed_d = torch.as_tensor(beds).cuda()
    res_d = torch.empty_like(bed_d)
    lin, exact = (torch.empty(C, dtype=torch.float64, device="cuda") for _ in range(2))
    ctx.residual_loss(bed_d, None, lin, None)                                 # loss-only: linear form
    ctx.residual_loss(bed_d, res_d, exact, None)                              # residual written: flux form, exact division
    lin, exact = lin.cpu().numpy(), exact.cpu().numpy()
    with np.errstate(all="ignore"):
        for c in range(C):
            ref = O.mass_conservation_residual(beds[c], surf, velx, vely, dhdt, smb, res_m)
            ref_loss = O.masked_loss(ref, mask, sigma)[0]
            if np.isinf(ref_loss):
                assert c == 3 and np.isinf(lin[c]) and np.isinf(exact[c])
                continue
            assert abs(exact[c] - ref_loss) <= 1e-12 * ref_loss
            assert abs(lin[c] - ref_loss) <= 1e-12 * ref_loss, (c, lin[c], ref_loss)


def test_signed_zero_and_special_quotients_are_bit_identical():
    """Zero thickness makes every flux +-0, so every quotient is +-0 (sign from the velocities); with dhdt = -0 the
    sign survives into the residual.  Also huge / tiny / infinite fluxes, which leave the fast division path."""
    import torch
    from gpu_helpers import bits_equal
    from mcmc_gpu_b200._lib import Context
    H, W = 40, 200
    g = np.random.default_rng(5)
    surf = 1000.0 + g.standard_normal((H, W))
    velx, vely = g.standard_normal((H, W)), g.standard_normal((H, W))
    dhdt, smb = np.full((H, W), -0.0), np.zeros((H, W))
    beds = np.stack([surf.copy(), surf - 1.0, surf.copy()])
    beds[2, 10:20, 50:90] -= 1e300           # huge thickness -> huge fluxes, inf differences
    beds[2, 25:30, 100:140] -= 1e-310        # below the spacing of surf: thickness still exactly 0
    velx2 = velx.copy()
    velx2[5, 5], velx2[6, 100] = np.inf, 1e-320
    for vx in (velx, velx2):
        ctx = Context(H, W, 3)
        ctx.set_static(surf, vx, vely, dhdt, smb, np.ones((H, W)), np.ones((H, W)), None, None, 500.0, 1.0)
        bed_d = torch.as_tensor(beds).cuda()
        res_d = torch.empty_like(bed_d)
        ctx.residual(bed_d, res_d)
        res = res_d.cpu().numpy()
        with np.errstate(all="ignore"):
            for c in range(3):
                ref = O.mass_conservation_residual(beds[c], surf, vx, vely, dhdt, smb, 500.0)
                nan = np.isnan(ref)
                assert np.array_equal(np.isnan(res[c]), nan)
                assert bits_equal(np.where(nan, 0.0, res[c]), np.where(nan, 0.0, ref)), c
        assert np.signbit(res[0]).any() and (~np.signbit(res[0])).any()


def test_chain_loss_method_matches_golden():
    from gpu_helpers import quiet
    from mcmc_gpu_b200 import MCMC
    ri = residual_case_inputs()
    gold = np.load(os.path.join(GOLD, "residual_loss.npz"))
    ch = quiet(MCMC.chain_crf, ri["xx"], ri["yy"], ri["bed"], ri["surf"], ri["velx"], ri["vely"], ri["dhdt"], ri["smb"],
               ri["bed"], ri["mask"], ri["mask"], ri["resolution"])
    quiet(ch.set_update_region, True, ri["mask"])
    ch.set_loss_type(sigma_mc=ri["sigma_mc"], massConvInRegion=True)
    total, mc, data = ch.loss(gold["residual"], 0)
    assert data == 0 and total == mc
    assert abs(total - gold["loss"][0]) <= 1e-12 * gold["loss"][0]


def test_errors_are_loud():
    from mcmc_gpu_b200._lib import Context, GmcError, GmcShapeError
    import torch
    ctx = Context(8, 8, 2)
    bed = torch.zeros((2, 8, 8), dtype=torch.float64, device="cuda")
    with pytest.raises(GmcError):                       # set_static not called
        ctx.residual(bed, torch.empty_like(bed))
    z = np.zeros((8, 8))
    ctx.set_static(z, z, z, z, z, np.ones((8, 8)), np.ones((8, 8)), None, None, 1.0, 1.0)
    with pytest.raises(GmcShapeError):                  # more chains than the context holds
        big = torch.zeros((3, 8, 8), dtype=torch.float64, device="cuda")
        ctx.residual(big, torch.empty_like(big))
    with pytest.raises(GmcShapeError):
        Context(1, 8, 1)


@pytest.mark.parametrize("divisor", [500.0, 1000.0, 125.0, 250.0, 0.1, 3.0, 7e-3, 1.0 / 3.0, 999.9999999999999, 1.9999999999999998, 1.0000000000000002])
def test_division_by_constant_is_correctly_rounded(divisor):
    """The stencil divides by res and 2*res through a reciprocal + one FMA residual correction; it must return the same
    bits as IEEE division (numpy's `/`) for every dividend: random, near powers of two, huge, tiny, zero, inf, nan."""
    import torch
    from mcmc_gpu_b200._lib import Context
    ctx = Context(4, 4, 1)
    g = np.random.default_rng(int(divisor * 1000) % 2 ** 31)
    n = 1 << 22
    x = np.concatenate([
        g.standard_normal(n) * 10.0 ** g.integers(-12, 12, n),
        np.ldexp(1.0 + g.integers(0, 4, n // 4) * 2.0 ** -52, g.integers(-60, 60, n // 4)) * divisor,     # quotients next to 2^k
        np.ldexp(2.0 - g.integers(1, 4, n // 4) * 2.0 ** -52, g.integers(-60, 60, n // 4)) * divisor,
        (g.integers(1, 2 ** 53, n // 4).astype(np.float64)) * divisor,                                       # exact-ish quotients
        np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-320, -1e-320, 1e308, -1e308, 5e-324, 1e-300, 1e300, 2.0 ** -1022]),
    ])
    xd = torch.as_tensor(x).cuda()
    assert ctx.div_check(xd, divisor) == 0


def test_min_dist_from_mask_is_bit_identical_to_the_kdtree():
    """Utilities.min_dist_from_mask on the GPU vs the reference's KD-tree query (Utilities.py:21-24)."""
    from scipy.spatial import KDTree
    from mcmc_gpu_b200 import Utilities
    g = np.random.default_rng(3)
    for (H, W, res, frac) in [(64, 80, 500.0, 0.02), (37, 53, 0.1, 0.3), (120, 90, 123.456, 0.001)]:
        xx, yy = np.meshgrid(np.arange(W) * res, np.arange(H) * res * 0.7)
        mask = g.random((H, W)) < frac
        mask[H // 2, W // 2] = True
        tree = KDTree(np.array([xx[mask], yy[mask]]).T)
        ref = tree.query(np.array([xx.ravel(), yy.ravel()]).T)[0].reshape(xx.shape)
        got = Utilities.min_dist_from_mask(xx, yy, mask)
        assert np.array_equal(got, ref), (H, W)
