"""Whole-grid Sequential Gaussian Simulation on the GPU — mirror of gstatsMCMC/gstatsim_custom/interpolate.py:92-191.

`sgs(xx, yy, grid, variogram, ...)` keeps the reference's signature and semantics for what the workflows use
(T2_StatisticalAnalysis: ordinary kriging, octant search, scalar variogram, per-cell bounds, integer seeds).  The path
shuffle and the per-node random numbers are drawn on the host from the SAME numpy generator calls, in the same order, as
the reference makes (`rng.shuffle(inds)`, then one `rng.normal` or — with bounds — one uniform inside
`truncnorm.rvs` per simulated node), so a given seed reproduces the reference's realisation to rounding; the
neighbour searches and kriging solves — all of the cost, 16 minutes per realisation in the reference — run as two
kernels (csrc/sgs.cu: every node's search+solve in parallel, then the values in path order).  `sgs_many` simulates
several seeds in one batch.
"""
from __future__ import annotations

import numbers

import numpy as np

from .. import _lib
from ..sgs_tables import MAX_HALF_WIDTH, MAX_LEVELS, covariance_lut, grid_steps, octant_stencil, search_levels  # noqa: F401

MAX_POINTS = 48


def _generator(seed):
    """utilities.get_random_generator (utilities.py:48-68)."""
    if seed is None:
        return np.random.default_rng()
    if isinstance(seed, numbers.Integral) and not isinstance(seed, bool):
        return np.random.default_rng(seed=int(seed))
    if isinstance(seed, np.random.Generator):
        return seed
    raise ValueError("Seed should be an integer, a NumPy random Generator, or None")


def _sanity_checks(xx, yy, grid, vario, radius, num_points, ktype, sim_mask):
    """interpolate._sanity_checks (interpolate.py:262-330), same messages."""
    for name, a in (("xx", xx), ("yy", yy), ("grid", grid)):
        if not isinstance(a, np.ndarray) or a.ndim != 2:
            raise ValueError(f"{name} must be a 2D NumPy array")
    if (xx.shape != yy.shape) or (xx.shape != grid.shape):
        raise ValueError("xx, yy, and grid must have same shape")
    missing = [k for k in ("major_range", "minor_range", "azimuth", "sill", "nugget", "vtype") if k not in vario]
    if missing:
        raise ValueError(f"Variogram missing {', '.join(missing)}")
    if vario["vtype"].lower() not in ("exponential", "gaussian", "spherical", "matern"):
        raise ValueError("vtype must be exponential, gaussian, spherical, or matern")
    if vario["vtype"].lower() == "matern" and "s" not in vario:
        raise ValueError("Matern covariance requires the s parameter in the variogram")
    if sim_mask is not None and (not isinstance(sim_mask, np.ndarray) or sim_mask.shape != grid.shape):
        raise ValueError("sim_mask shape must be same as grid if provided")
    if ktype not in ("ok", "sk"):
        raise ValueError("ktype must be 'ok' or 'sk'")


def sgs_many(xx, yy, grid, variogram, seeds, radius=100e3, num_points=20, ktype="ok", sim_mask=None, bounds=None,
             n_quantiles=500, nst_tables=None, as_tensor=False):
    """Realisations for every seed in `seeds` (ints or numpy Generators), stacked [len(seeds), H, W].

    nst_tables=(quantiles, references) reuses a fitted normal-score transform instead of fitting
    QuantileTransformer(n_quantiles) on the conditioning data as the reference does (utilities.py:20)."""
    import torch
    _sanity_checks(xx, yy, grid, variogram, radius, num_points, ktype, sim_mask)
    if ktype != "ok":
        raise NotImplementedError("only ordinary kriging (ktype='ok', the reference's default) runs on the GPU")
    if num_points < 8 or num_points > MAX_POINTS:
        raise NotImplementedError(f"num_points must lie in [8, {MAX_POINTS}] (num_points//8 per octant)")
    for k, v in variogram.items():
        if k != "vtype" and not isinstance(v, numbers.Number):
            raise NotImplementedError("spatially varying variogram parameters are not supported on the GPU path")
    dev = _lib.require_cuda()
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    H, W = grid.shape
    n_real = len(seeds)
    cond = ~np.isnan(grid)
    if not cond.any():
        raise ValueError("grid holds no conditioning data")
    if nst_tables is None:
        from sklearn.preprocessing import QuantileTransformer          # the reference's own dependency (utilities.py:2)
        qt = QuantileTransformer(n_quantiles=n_quantiles, output_distribution="normal").fit(grid[cond].reshape(-1, 1))
        quant, refs = qt.quantiles_[:, 0], qt.references_
    else:
        quant, refs = (np.asarray(a, dtype=np.float64) for a in nst_tables)

    def cu(a, dt=torch.float64):
        return torch.as_tensor(np.ascontiguousarray(a)).to(dev, dtype=dt)

    q_d, r_d = cu(quant), cu(refs)

    def transform(t, inverse):
        out = torch.empty_like(t)
        _lib.check(lib.gmc_nst_transform(dev.index, q_d.data_ptr(), r_d.data_ptr(), int(q_d.numel()), t.data_ptr(), out.data_ptr(),
                                         int(t.numel()), int(inverse), st))
        return out

    z0 = transform(cu(np.where(cond, grid, np.nan).ravel()), False)       # NaN stays NaN
    blo = bhi = None
    if bounds is not None:
        try:
            if len(bounds) != 2:
                raise ValueError
            tb = []
            for b in bounds:
                if isinstance(b, numbers.Number):
                    arr = np.full(grid.shape, float(b))
                elif isinstance(b, np.ndarray) and b.shape == grid.shape:
                    arr = b.astype(np.float64)
                else:
                    raise ValueError
                tb.append(transform(cu(arr.ravel()), False))
        except Exception:
            raise ValueError("bounds must be None or a 2D numpy array")
        blo, bhi = tb
        pinned = (blo == bhi).cpu().numpy().reshape(H, W)

    # the reference's generator calls, in its order (interpolate.py:122-125, 173, 181)
    ii, jj = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    sel = np.full((H, W), True) if sim_mask is None else sim_mask.astype(bool)
    n_path = int(sel.sum())
    paths = np.empty((n_real, n_path), dtype=np.int32)
    ords = np.empty((n_real, H * W), dtype=np.int32)
    noise = np.zeros((n_real, n_path))
    base_ord = np.where(cond.ravel(), -1, np.iinfo(np.int32).max).astype(np.int32)     # never simulated: never available
    for r, seed in enumerate(seeds):
        rng = _generator(seed)
        inds = np.array([ii[sel].flatten(), jj[sel].flatten()]).T
        rng.shuffle(inds)
        cells = (inds[:, 0] * W + inds[:, 1]).astype(np.int32)
        paths[r] = cells
        o = base_ord.copy()
        sim = ~cond.ravel()[cells]                                         # nodes that get a draw, in path order
        o[cells[sim]] = np.nonzero(sim)[0].astype(np.int32)
        ords[r] = o
        if bounds is None:
            noise[r, sim] = rng.standard_normal(int(sim.sum()))
        else:
            draws = sim & ~pinned.ravel()[cells]
            noise[r, draws] = rng.random(int(draws.sum()))
    path_d, ord_d, noise_d = cu(paths, torch.int32), cu(ords, torch.int32), cu(noise)

    dx, dy = grid_steps(xx, yy)
    off, cnt, hw, radii = search_levels(dx, dy, H, W, radius)
    vario = {k: (v.lower() if k == "vtype" else float(v)) for k, v in variogram.items()}
    lut = covariance_lut(dx, dy, hw, vario)
    off_d, cnt_d, lut_d = cu(off, torch.int16), cu(cnt, torch.int32), cu(lut)
    items = n_real * n_path
    rec_n = torch.empty(items, dtype=torch.int32, device=dev)
    rec_idx = torch.empty(items * MAX_POINTS, dtype=torch.int32, device=dev)
    rec_w = torch.empty(items * MAX_POINTS, dtype=torch.float64, device=dev)
    rec_sd = torch.empty(items, dtype=torch.float64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.gmc_sgs_grid_solve(dev.index, H, W, ord_d.data_ptr(), path_d.data_ptr(), n_path, n_real, off_d.data_ptr(),
                                      cnt_d.data_ptr(), int(cnt.shape[0]), int(off.shape[1]), int(hw), int(num_points), lut_d.data_ptr(),
                                      float(vario["sill"]), rec_n.data_ptr(), rec_idx.data_ptr(), rec_w.data_ptr(),
                                      rec_sd.data_ptr(), err.data_ptr(), st))
    z = z0[None].repeat(n_real, 1).contiguous()
    _lib.check(lib.gmc_sgs_grid_values(dev.index, H, W, z.data_ptr(), path_d.data_ptr(), n_path, n_real, rec_n.data_ptr(),
                                       rec_idx.data_ptr(), rec_w.data_ptr(), rec_sd.data_ptr(), noise_d.data_ptr(),
                                       blo.data_ptr() if blo is not None else None, bhi.data_ptr() if bhi is not None else None, st))
    if int(err.item()) & 1:
        raise NotImplementedError(f"a node found no conditioning data within {radii[-1] / 1e3:.0f} km (the search was widened "
                                  f"{len(radii) - 1} times by 100 km, interpolate.py:149-155) - pass a larger radius")
    sim = transform(z.reshape(-1), True).reshape(n_real, H, W)
    return sim if as_tensor else sim.cpu().numpy()


def sgs(xx, yy, grid, variogram, radius=100e3, num_points=20, ktype="ok", sim_mask=None, quiet=False, stencil=None, rcond=None,
        bounds=None, seed=None):
    """One realisation [H, W] — the reference's `interpolate.sgs` (interpolate.py:92).  `quiet` (no progress bar here),
    `stencil` (only its size is used by the reference's search, neighbors.py:28) and `rcond` (the systems are solved
    exactly instead of by lstsq) are accepted and have no effect."""
    return sgs_many(xx, yy, grid, variogram, [seed], radius=radius, num_points=num_points, ktype=ktype, sim_mask=sim_mask,
                    bounds=bounds)[0]
