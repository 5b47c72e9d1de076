"""CPU ORACLE (test infrastructure, never the product path) for the small-scale SGS chain step.

numpy/scipy restatement of the reference's block re-simulation chain: `chain_sgs.run` (MCMC.py:1599-1911), `sgs`
(MCMC.py:91-173), the octant neighbour search (gstatsim_custom/neighbors.py:4-64), ordinary kriging
(gstatsim_custom/_krige.py:5-44, 83-143), the covariance models (gstatsim_custom/covariance.py:4-29) and sklearn's
QuantileTransformer column transform (sklearn/preprocessing/_data.py `_transform_col`, called at MCMC.py:1653-1654, 1767,
1777).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Parity status: PINNED against the unmodified reference by tests/golden/sgs_*.npz (oracle/make_golden.py): with
TIE_ORDER = "numpy" the oracle reproduces the reference trajectories bit-for-bit in the build container.  One
ingredient of the reference is not well defined: it ranks the candidates of an octant with `np.argsort` (default kind:
introsort / SIMD sort), whose order among EQUAL distances is implementation- and CPU-dependent (ties are common on a
regular grid).  TIE_ORDER = "stable" (ties broken by row-major position in the search window) is the deterministic rule
the CUDA kernel implements; the GPU parity tests run the oracle in that mode, and tests/test_oracle_sgs.py shows the two
modes agree wherever no tie straddles the cut-off.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
from scipy.special import gamma, kv
from scipy.stats import norm

from oracle.crf_oracle import mass_conservation_residual, masked_loss

BOUNDS_THRESHOLD = 1e-7
TIE_ORDER = "stable"       # "numpy": np.argsort's default kind, as the reference calls it


# ---- covariance models (functions of the range-normalised lag)          gstatsim_custom/covariance.py:4-29 -------
def covariance(vtype, h, sill, nugget, s=None):
    h = np.array(h, dtype=np.float64, copy=True)
    vt = vtype.lower()
    if vt == "exponential":
        return (sill - nugget) * np.exp(-3 * h)
    if vt == "gaussian":
        return (sill - nugget) * np.exp(-3 * np.square(h))
    if vt == "spherical":                       # includes the reference's units quirk (covariance.py:13-14)
        c = sill - nugget - 1.5 * h + 0.5 * np.power(h, 3)
        c[h > 1] = sill - 1
        return c
    if vt == "matern":
        scale = 0.45246434 * np.exp(-0.70449189 * s) + 1.7863836
        h[h == 0.0] = 1e-8
        c = (sill - nugget) * 2 / gamma(s) * np.power(scale * h * np.sqrt(s), s) * kv(s, 2 * scale * h * np.sqrt(s))
        c[np.isnan(c)] = sill - nugget
        return c
    raise ValueError(vtype)


def rotation_matrix(azimuth, major_range, minor_range):
    """gstatsim_custom/_krige.py:83-102."""
    theta = (azimuth / 180.0) * np.pi
    return np.dot(np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]]),
                  np.array([[1 / major_range, 0], [0, 1 / minor_range]]))


def ok_solve(sim_xy, nearest, vario):
    """Ordinary kriging estimate and variance from the bordered system via lstsq.  _krige.py:5-44."""
    from scipy.spatial.distance import pdist, squareform
    R = rotation_matrix(vario["azimuth"], vario["major_range"], vario["minor_range"])
    xy = nearest[:, :2]
    local_mean = np.mean(nearest[:, 2])
    n = nearest.shape[0]
    Sigma = np.zeros((n + 1, n + 1))
    Sigma[0:n, 0:n] = covariance(vario["vtype"], squareform(pdist(xy @ R)), vario["sill"], vario["nugget"], vario.get("s"))
    Sigma[n, 0:n] = 1
    Sigma[0:n, n] = 1
    rho = np.zeros(n + 1)
    m1, m2 = xy @ R, np.asarray(sim_xy) @ R
    rho[0:n] = covariance(vario["vtype"], np.sqrt(np.square(m1 - m2).sum(axis=1)), vario["sill"], vario["nugget"], vario.get("s"))
    rho[n] = 1
    w = np.linalg.lstsq(Sigma, rho, rcond=None)[0]
    var = vario["sill"] - np.sum(w[0:n] * rho[0:n])
    est = local_mean + np.sum(w[0:n] * (nearest[:, 2] - local_mean))
    return est, var


# ---- octant neighbour search                                          gstatsim_custom/neighbors.py:4-64 ----------
def stencil_half_width(x_row, radius):
    """Half width (cells) of the search window = make_circle_stencil(...).shape[0]//2.  neighbors.py:66-83."""
    dx = np.abs(x_row[1] - x_row[0])
    return math.ceil(radius / dx)


def octant_neighbors(i, j, xx, yy, grid, cond_msk, radius, num_points, hw):
    ni, nj = grid.shape
    ilow, ihigh = max(0, i - hw), min(ni, i + hw + 1)
    jlow, jhigh = max(0, j - hw), min(nj, j + hw + 1)
    sl = (slice(ilow, ihigh), slice(jlow, jhigh))
    g, x, y, c = grid[sl], xx[sl], yy[sl], cond_msk[sl]
    ii, jj = np.meshgrid(np.arange(ilow, ihigh), np.arange(jlow, jhigh), indexing="ij")
    dist = np.sqrt((xx[i, j] - x) ** 2 + (yy[i, j] - y) ** 2)
    ang = np.arctan2(yy[i, j] - y, xx[i, j] - x)
    pts = []
    for b in range(-4, 4, 1):
        m = (dist < radius) & (ang > b / 4 * np.pi) & (ang <= (b + 1) / 4 * np.pi) & c
        order = np.argsort(dist[m], kind="stable") if TIE_ORDER == "stable" else np.argsort(dist[m])
        p = np.array([x[m], y[m], g[m], ii[m], jj[m]]).T[order, :][:num_points // 8, :]
        pts.append(p)
    pts = np.concatenate(pts)
    return pts[~np.isnan(pts[:, 2]), :]


def sgs_block(xx, yy, grid, vario, radius, num_points, sim_mask, rng, record=None, replay=None):
    """MCMC.sgs with ktype='ok' (MCMC.py:91-173): shuffle the block's cells, krige + draw each unconditioned one.

    record: dict that receives path (shuffled [n,2] indices) and z (the unit normals used, NaN for conditioned nodes).
    replay: dict(path, z) to inject instead of drawing from rng."""
    cond = ~np.isnan(grid)
    out = grid.copy()
    ii, jj = np.meshgrid(np.arange(xx.shape[0]), np.arange(xx.shape[1]), indexing="ij")
    inds = np.array([ii[sim_mask].flatten(), jj[sim_mask].flatten()]).T
    if replay is None:
        rng.shuffle(inds)
    else:
        inds = np.array(replay["path"])
    zs = np.full(inds.shape[0], np.nan)
    hw0 = stencil_half_width(xx[0, :], radius)
    for k in range(inds.shape[0]):
        i, j = inds[k]
        if cond[i, j]:
            continue
        rad, hw = radius, hw0
        while True:
            nearest = octant_neighbors(i, j, xx, yy, out, cond, rad, num_points, hw)
            if nearest.shape[0] > 0:
                break
            rad += 100e3
            hw = stencil_half_width(xx[0, :], rad)
        est, var = ok_solve((xx[i, j], yy[i, j]), nearest, vario)
        var = np.abs(var)
        if replay is None:
            draw = rng.normal(est, np.sqrt(var), 1)[0]          # = est + sqrt(var) * z
            zs[k] = (draw - est) / np.sqrt(var) if var > 0 else 0.0
        else:
            zs[k] = replay["z"][k]
            draw = est + np.sqrt(var) * zs[k]
        out[i, j] = draw
        cond[i, j] = True
    if record is not None:
        record.update(path=inds.copy(), z=zs)
    return out


def truncnorm_ppf(u, a, b):
    """What scipy.stats.truncnorm.rvs applies to its one uniform draw (interpolate.py:181).  scipy evaluates it in log
    space; the kernel uses  Phi^-1(Phi(a) + u (Phi(b) - Phi(a)))  (through the survival function when a > 0), which
    agrees to rounding (tests/test_oracle_sgs.py)."""
    from scipy.stats import truncnorm
    return float(truncnorm.ppf(u, a, b))


def truncnorm_ppf_direct(u, a, b):
    """The closed form the kernel evaluates."""
    if a > 0:                                   # both bounds in the upper tail: work with survival probabilities
        sa, sb = norm.cdf(-a), norm.cdf(-b)
        return -norm.ppf(sa - u * (sa - sb))
    pa, pb = norm.cdf(a), norm.cdf(b)
    return norm.ppf(pa + u * (pb - pa))


def sgs_grid(xx, yy, grid, vario, radius, num_points, rng, bounds=None, n_quantiles=500, record=None, replay=None):
    """Whole-grid Sequential Gaussian Simulation with bounds: gstatsim_custom/interpolate.py:92-191 (ktype='ok',
    sim_mask=None, scalar variogram) with its `_preprocess` (:193-260) and utilities.gaussian_transformation (:7-26).

    Returns (simulation in data units, NormalScore used).  record: dict that receives path [N,2], noise [N] (the standard
    normal of an unbounded node / the uniform of a bounded one, NaN for conditioning cells) and the normal-score tables;
    replay: dict(path, noise, quantiles, references) injects them instead of drawing from rng / refitting."""
    from sklearn.preprocessing import QuantileTransformer
    cond = ~np.isnan(grid)                                                                  # :207
    if replay is None:
        qt = QuantileTransformer(n_quantiles=n_quantiles, output_distribution="normal").fit(grid[cond].reshape(-1, 1))
        nst = NormalScore(qt.quantiles_[:, 0].copy(), qt.references_.copy())
    else:
        nst = NormalScore(np.asarray(replay["quantiles"]), np.asarray(replay["references"]))
    out = np.full(grid.shape, np.nan)
    out[cond] = nst.forward(grid[cond])                                                     # utilities.py:21-24
    ii, jj = np.meshgrid(np.arange(xx.shape[0]), np.arange(xx.shape[1]), indexing="ij")
    inds = np.array([ii.flatten(), jj.flatten()]).T                                         # :214-215
    tb = None
    if bounds is not None:                                                                  # :233-252
        tb = []
        for bnd in bounds:
            arr = np.full(xx.shape, float(bnd)) if np.isscalar(bnd) else np.asarray(bnd, dtype=np.float64)
            tb.append(nst.forward(arr.reshape(-1)).reshape(xx.shape))
    if replay is None:
        rng.shuffle(inds)                                                                   # :125
    else:
        inds = np.array(replay["path"])
    noise = np.full(inds.shape[0], np.nan)
    hw0 = stencil_half_width(xx[0, :], radius)
    for k in range(inds.shape[0]):
        i, j = inds[k]
        if cond[i, j]:
            continue
        rad, hw = radius, hw0
        while True:                                                                         # :149-155
            nearest = octant_neighbors(i, j, xx, yy, out, cond, rad, num_points, hw)
            if nearest.shape[0] > 0:
                break
            rad += 100e3
            hw = stencil_half_width(xx[0, :], rad)
        est, var = ok_solve((xx[i, j], yy[i, j]), nearest, vario)
        sd = np.sqrt(np.abs(var))                                                           # :163
        if tb is None:
            if replay is None:
                draw = rng.normal(est, sd, 1)[0]
                noise[k] = (draw - est) / sd if sd > 0 else 0.0
            else:
                noise[k] = replay["noise"][k]
                draw = est + sd * noise[k]
        elif tb[0][i, j] == tb[1][i, j]:                                                    # :177-178
            draw = tb[0][i, j]
        else:                                                                               # :180-181 truncnorm.rvs
            a, b = (tb[0][i, j] - est) / sd, (tb[1][i, j] - est) / sd
            noise[k] = rng.uniform() if replay is None else replay["noise"][k]
            draw = est + sd * truncnorm_ppf(noise[k], a, b)
        out[i, j] = draw
        cond[i, j] = True
    if record is not None:
        record.update(path=inds.copy(), noise=noise, quantiles=nst.quantiles, references=nst.references)
    return nst.inverse(out.reshape(-1)).reshape(xx.shape), nst                              # :185


# ---- QuantileTransformer(output_distribution='normal') column transform      sklearn _data.py _transform_col ------
@dataclass
class NormalScore:
    quantiles: np.ndarray       # nst_trans.quantiles_[:, 0]
    references: np.ndarray      # nst_trans.references_

    def forward(self, x):
        x = np.array(x, dtype=np.float64, copy=True)
        q, r = self.quantiles, self.references
        with np.errstate(invalid="ignore"):
            lo = x - BOUNDS_THRESHOLD < q[0]
            hi = x + BOUNDS_THRESHOLD > q[-1]
        fin = ~np.isnan(x)
        xf = x[fin]
        x[fin] = 0.5 * (np.interp(xf, q, r) - np.interp(-xf, -q[::-1], -r[::-1]))
        x[hi] = 1
        x[lo] = 0
        with np.errstate(invalid="ignore"):
            x = norm.ppf(x)
            cmin = norm.ppf(BOUNDS_THRESHOLD - np.spacing(1))
            cmax = norm.ppf(1 - (BOUNDS_THRESHOLD - np.spacing(1)))
            x = np.clip(x, cmin, cmax)
        return x

    def inverse(self, z):
        q, r = self.quantiles, self.references
        with np.errstate(invalid="ignore"):
            x = norm.cdf(np.array(z, dtype=np.float64, copy=True))
            lo = x - BOUNDS_THRESHOLD < 0
            hi = x + BOUNDS_THRESHOLD > 1
        fin = ~np.isnan(x)
        x[fin] = np.interp(x[fin], r, q)
        x[hi] = q[-1]
        x[lo] = q[0]
        return x


# ---- the chain                                                                     MCMC.py:1599-1829 ------------
@dataclass
class SgsSetup:
    xx: np.ndarray
    yy: np.ndarray
    surf: np.ndarray
    velx: np.ndarray
    vely: np.ndarray
    dhdt: np.ndarray
    smb: np.ndarray
    cond_bed: np.ndarray
    grounded_ice_mask: np.ndarray
    region_mask: np.ndarray
    mc_region_mask: np.ndarray
    resolution: float
    sigma_mc: float
    trend: np.ndarray | None
    nst: NormalScore | None
    vario: dict
    num_points: int
    radius: float
    block: tuple          # (min_x, max_x, min_y, max_y), upper bounds exclusive (MCMC.py:1755-1756)


def sgs_chain_run(su: SgsSetup, initial_bed, n_iter, rng, record=False, replay=None):
    """chain_sgs.run(n_iter, only_save_last_bed=True, plot=False, progress_bar=False).  Returns the reference's outputs
    (bed WITHOUT trend, like the reference's only_save_last_bed branch, MCMC.py:1907-1911) and optionally the tape."""
    H, W = su.xx.shape
    trend = su.trend if su.trend is not None else 0.0
    bed_c = (initial_bed - trend).copy() if su.trend is not None else initial_bed.copy()
    cond_c = (su.cond_bed - trend).copy() if su.trend is not None else su.cond_bed.copy()
    z_cond = su.nst.forward(cond_c.reshape(-1)).reshape(H, W) if su.nst is not None else cond_c.copy()
    res = mass_conservation_residual(bed_c + trend, su.surf, su.velx, su.vely, su.dhdt, su.smb, su.resolution)
    loss_prev = masked_loss(res, su.mc_region_mask, su.sigma_mc)[0]
    loss_cache, step_cache = np.zeros(n_iter), np.zeros(n_iter)
    blocks = np.full((n_iter, 4), np.nan)
    resampled = np.zeros((H, W))
    loss_cache[0] = loss_prev
    tape = [] if record else None
    for it in range(n_iter):
        rp = replay[it] if replay is not None else None
        if rp is None:
            while True:
                ix = rng.integers(low=0, high=H, size=1)[0]
                iy = rng.integers(low=0, high=W, size=1)[0]
                if su.region_mask[ix, iy] == 1:
                    break
            bsx = rng.integers(low=su.block[0], high=su.block[1], size=1)[0]
            bsy = rng.integers(low=su.block[2], high=su.block[3], size=1)[0]
        else:
            ix, iy, bsx, bsy = rp["idx_x"], rp["idx_y"], rp["bsx"], rp["bsy"]
        blocks[it, :] = [ix, iy, bsx, bsy]
        x0, x1 = max(0, int(ix - bsx / 2)), min(H, int(ix + bsx / 2))
        y0, y1 = max(0, int(iy - bsy / 2)), min(W, int(iy + bsy / 2))
        tosim = su.nst.forward(bed_c.reshape(-1)).reshape(H, W) if su.nst is not None else bed_c.copy()
        tosim[x0:x1, y0:y1] = z_cond[x0:x1, y0:y1].copy()
        sim_mask = np.full((H, W), False)
        sim_mask[x0:x1, y0:y1] = True
        rec = {} if record else None
        newsim = sgs_block(su.xx, su.yy, tosim, su.vario, su.radius, su.num_points, sim_mask, rng, rec, rp)
        bed_next = su.nst.inverse(newsim.reshape(-1)).reshape(H, W) if su.nst is not None else newsim.copy()
        res = mass_conservation_residual(bed_next + trend, su.surf, su.velx, su.vely, su.dhdt, su.smb, su.resolution)
        loss_next = masked_loss(res, su.mc_region_mask, su.sigma_mc)[0]
        thick = su.surf - (bed_next + trend)
        if np.sum((thick <= 0)[su.grounded_ice_mask == 1]) > 0:
            loss_next = np.inf
        acc = 1 if loss_prev > loss_next else min(1, np.exp(loss_prev - loss_next))
        u = rng.random() if rp is None else rp["u"]
        ok = bool(u <= acc)
        if ok:
            bed_c = bed_next
            loss_prev = loss_next
            resampled[x0:x1, y0:y1] += 1
        loss_cache[it] = loss_prev
        step_cache[it] = ok
        if record:
            rec.update(idx_x=int(ix), idx_y=int(iy), bsx=int(bsx), bsy=int(bsy), u=float(u), loss_next=float(loss_next))
            tape.append(rec)
    last_bed = bed_c + trend if su.trend is not None else bed_c          # MCMC.py:1897-1900
    return dict(bed=last_bed, bed_c=bed_c, loss=loss_cache, loss_mc=loss_cache.copy(), loss_data=np.zeros(n_iter), steps=step_cache,
                resampled_times=resampled, blocks=blocks, tape=tape)
