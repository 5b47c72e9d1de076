"""Named test cases shared by oracle/make_golden.py (reference run) and the tests (oracle / CUDA runs)."""
from __future__ import annotations

import numpy as np

from mcmc_gpu_b200 import synthetic as syn

MATERN = dict(syn.RF_KW)

# Each trajectory case is one `chain_crf.run(n_iter, RF, only_save_last_bed=True)` of the reference.
TRAJECTORY_CASES = {
    # BASELINE.json config 1 at reduced length: the tutorial configuration, SURVEY §8c anchor
    # (final loss 348.2912467039959 at n_iter=500).
    "tutorial200": dict(H=200, W=200, n_iter=500, rf_seed=7, chain_seed=11, rf_kw=MATERN, blocks=(50, 80, 50, 80),
                        logistic=(2.0, 0.0, 6.0, 1.0), max_dist=30e3, sigma_mc=5.0, update_in_region=True,
                        block_type="CRF_weight"),
    # Non-square grid, small blocks clipped by every edge, taper that is NOT zero on the rim (stale-ring quirk),
    # nugget noise, anisotropic draw order, plain 'RF' update over the whole map.
    "ragged_rf": dict(H=72, W=90, n_iter=400, rf_seed=3, chain_seed=5,
                      rf_kw=dict(range_min_x=2e3, range_max_x=9e3, range_min_y=3e3, range_max_y=8e3, scale_min=20.0,
                                 scale_max=60.0, nugget_max=4.0, model_name="Exponential", isotropic=False,
                                 smoothness=None),
                      blocks=(10, 30, 12, 26), logistic=(1.0, 0.5, 6.0, 0.0), max_dist=3e3, sigma_mc=10.0,
                      update_in_region=False, block_type="RF"),
    # Thin ice: the thickness guard (loss = inf) fires on part of the proposals; Gaussian model.
    "thin_ice": dict(H=96, W=80, n_iter=400, rf_seed=21, chain_seed=22,
                     rf_kw=dict(range_min_x=3e3, range_max_x=10e3, range_min_y=3e3, range_max_y=10e3, scale_min=60.0,
                                scale_max=200.0, nugget_max=0.0, model_name="Gaussian", isotropic=True,
                                smoothness=None),
                     blocks=(16, 32, 16, 32), logistic=(2.0, 0.0, 6.0, 1.0), max_dist=4e3, sigma_mc=40.0,
                     update_in_region=True, block_type="CRF_weight", thin=True),
}


def build_case_grids(case: dict) -> dict:
    g = syn.make_grids(case["H"], case["W"])
    if case.get("thin"):
        # pull the surface down to ~100 m above the bed so about half the proposals hit non-positive thickness
        g["surf"] = g["bed0"] + 100.0 + 10.0 * np.sin(g["xx"] / 7e3)
        g["highvel_mask"] = ((np.hypot(g["velx"], g["vely"]) > 60.0) | (g["yy"] < 15e3)).astype(np.int64)
    return g


# Ranges are kept within a few cells of the block size: for a Gaussian spectrum with range >> block the reference's own
# output is rounding noise of the FFT around a cancelled DC term (ill-conditioned by ~1e9), so no two FFTs agree to 1e-9.
FIELD_CASES = {
    "matern_iso": dict(rf_kw=MATERN, seed=101, res=500.0, shapes=[(50, 56), (64, 64), (72, 80), (80, 50)]),
    "gauss_aniso_nug": dict(rf_kw=dict(range_min_x=1e3, range_max_x=3e3, range_min_y=1.5e3, range_max_y=4e3,
                                       scale_min=10.0, scale_max=30.0, nugget_max=2.5, model_name="Gaussian",
                                       isotropic=False, smoothness=None), seed=102, res=250.0,
                            shapes=[(20, 24), (30, 18), (58, 22)]),
    "expo_iso": dict(rf_kw=dict(range_min_x=1e3, range_max_x=4e3, range_min_y=1e3, range_max_y=4e3, scale_min=1.0,
                                scale_max=2.0, nugget_max=0.0, model_name="Exponential", isotropic=True,
                                smoothness=None), seed=103, res=100.0, shapes=[(14, 66), (26, 26)]),
}


def residual_case_inputs() -> dict:
    """A ragged 37x53 grid with NaN holes, a mask with values {0,1} and rough random fields."""
    H, W = 37, 53
    g = np.random.default_rng(2024)
    res = 125.0
    xx, yy = np.meshgrid(np.arange(W) * res, np.arange(H) * res)
    surf = 1500.0 + 300.0 * g.standard_normal((H, W))
    bed = surf - 800.0 + 100.0 * g.standard_normal((H, W))
    velx = 200.0 * g.standard_normal((H, W))
    vely = 150.0 * g.standard_normal((H, W))
    dhdt = g.standard_normal((H, W))
    smb = g.standard_normal((H, W))
    bed[5, 7] = np.nan
    velx[20, 0] = np.nan
    smb[36, 52] = np.nan
    mask = (g.random((H, W)) < 0.6).astype(np.int64)
    mask[5, 6] = 1
    return dict(xx=xx, yy=yy, bed=bed, surf=surf, velx=velx, vely=vely, dhdt=dhdt, smb=smb, mask=mask,
                resolution=res, sigma_mc=2.5)


def highvel_case_inputs() -> dict:
    """44x57 grid: a fast ice stream with a noisy edge, a floating/ocean corner, speckle that the mode filter removes."""
    H, W = 44, 57
    g = np.random.default_rng(77)
    res = 500.0
    xx, yy = np.meshgrid(np.arange(W) * res, np.arange(H) * res)
    stream = 120.0 * np.exp(-((yy - 9e3 - 0.15 * xx) / 4e3) ** 2)
    velx = stream * 0.9 + 8.0 * g.standard_normal((H, W))
    vely = stream * 0.3 + 8.0 * g.standard_normal((H, W))
    velx[g.random((H, W)) < 0.03] = 300.0                    # isolated fast pixels
    grounded = np.ones((H, W), dtype=np.int64)
    grounded[30:, 40:] = 0                                   # floating ice / ocean
    grounded[3:6, 3:5] = 0                                   # an ungrounded hole inside the sheet
    ocean = (1 - grounded).astype(np.int64)
    ocean[3:6, 3:5] = 0
    return dict(xx=xx, yy=yy, velx=velx, vely=vely, threshold=50.0, grounded=grounded, ocean=ocean, distance_max=2200.0,
                smooth_mode=10)


# ---- small-scale SGS chain cases (chain_sgs.run, MCMC.py:1599) ------------------------------------------------------
SGS_CASES = {
    # tutorial-like: detrended + normal-score transformed bed, Matern variogram, ordinary kriging with octant search
    "matern_nst": dict(H=60, W=64, n_iter=60, seed=17, sigma_mc=1.5, blocks=(4, 10, 4, 10), neighbors=16, radius=5e3,
                       vario=dict(vtype="Matern", range=3000.0, sill=1.0, nugget=0.0, isotropic=True, smoothness=1.2259,
                                  azimuth=None),
                       transform=True, detrend=True, n_quantiles=500),
    # raw bed (no transform, no trend), anisotropic exponential variogram, one neighbour per octant, thin ice (guard fires)
    "expo_raw": dict(H=48, W=56, n_iter=60, seed=23, sigma_mc=20.0, blocks=(3, 9, 5, 8), neighbors=8, radius=4e3,
                     vario=dict(vtype="Exponential", range=[4000.0, 2500.0], sill=900.0, nugget=0.0, isotropic=False,
                                smoothness=None, azimuth=30.0),
                     transform=False, detrend=False, thin=True),
    # the tutorial's neighbourhood: 48 neighbours (6 per octant), i.e. full 48 x 48 kriging systems
    "matern_k48": dict(H=64, W=70, n_iter=30, seed=31, sigma_mc=1.5, blocks=(5, 12, 5, 12), neighbors=48, radius=9e3,
                       vario=dict(vtype="Matern", range=4000.0, sill=1.0, nugget=0.0, isotropic=True, smoothness=1.2259,
                                  azimuth=None),
                       transform=True, detrend=True, n_quantiles=500),
}


# ---- whole-grid SGS realisations (gstatsim_custom/interpolate.sgs): initial beds of the large-scale chains -----------
SGS_GRID_CASES = {
    "free": dict(H=30, W=34, cond_frac=0.12, seed=11, radius=6e3, num_points=16, bounds=False,
                 vario=dict(azimuth=0.0, nugget=0.0, major_range=4000.0, minor_range=4000.0, sill=1.0, s=1.2, vtype="matern")),
    "bounded_k48": dict(H=40, W=36, cond_frac=0.2, seed=12, radius=5e3, num_points=48, bounds=True,
                        vario=dict(azimuth=25.0, nugget=0.0, major_range=5000.0, minor_range=3000.0, sill=1.0, s=None,
                                   vtype="exponential")),
}


def sgs_grid_inputs(case: dict) -> dict:
    """Sparse conditioning data on a synthetic bed; bounds as in T2_StatisticalAnalysis (lower -9999, upper the surface,
    here lowered so that the truncation binds at some nodes)."""
    g = syn.make_grids(case["H"], case["W"])
    r = np.random.default_rng(1000 + case["seed"])
    cond = np.where(r.random((case["H"], case["W"])) < case["cond_frac"], g["bed0"], np.nan)
    bounds = None
    if case["bounds"]:
        upper = np.where(np.isnan(cond), g["bed0"] + 40.0, np.maximum(cond, g["bed0"] + 40.0))
        upper[5:9, 5:9] = -9999.0                  # lower == upper: the value is pinned (interpolate.py:177-178)
        upper[np.isfinite(cond)] = np.maximum(upper[np.isfinite(cond)], cond[np.isfinite(cond)])
        bounds = (np.full(cond.shape, -9999.0), upper)
    v = {k: x for k, x in case["vario"].items() if x is not None}
    return dict(xx=g["xx"], yy=g["yy"], cond=cond, bounds=bounds, vario=v)


def build_sgs_inputs(case: dict) -> dict:
    """Grids + trend + fitted normal-score tables for an SGS case (tables come from sklearn, as in the tutorials)."""
    from scipy.ndimage import gaussian_filter
    g = syn.make_grids(case["H"], case["W"])
    g = dict(g)
    if case.get("thin"):
        g["surf"] = g["bed0"] + 60.0 + 10.0 * np.sin(g["xx"] / 7e3)
    # rough initial bed so the residual is informative
    rng = np.random.default_rng(99)
    g["bed_init"] = g["bed0"] + gaussian_filter(rng.standard_normal(g["bed0"].shape), 2.0) * 30.0
    g["cond_bed"] = np.where(g["data_mask"] == 1, g["bed_init"], np.nan)
    g["trend"] = gaussian_filter(g["bed_init"], 5.0) if case["detrend"] else None
    g["quantiles"] = g["references"] = None
    if case["transform"]:
        from sklearn.preprocessing import QuantileTransformer
        base = g["bed_init"] - (g["trend"] if case["detrend"] else 0.0)
        nst = QuantileTransformer(n_quantiles=case["n_quantiles"], output_distribution="normal", subsample=None,
                                  random_state=0).fit(base.reshape(-1, 1))
        g["nst"] = nst
        g["quantiles"], g["references"] = nst.quantiles_[:, 0].copy(), nst.references_.copy()
    return g
