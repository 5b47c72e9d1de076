"""CPU ORACLE (test infrastructure, never the product path) for the large-scale chain step.

A numpy restatement of the reference algorithm of gstatsMCMC's random-field Metropolis chain.  Only
`tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of `bench.py` may import this module; the
shipped path (`mcmc_gpu_b200`) never does.

Parity status: PINNED.  `oracle/make_golden.py` ran the unmodified reference (`/root/reference`, numpy 2.3.5)
in the build container and stored its outputs in `tests/golden/*.npz`; `tests/test_oracle_golden.py`
requires this module to reproduce them bit-for-bit from the same seeds (same numpy Generator draw order).

Conventions follow the reference: axis 0 is "x" (rows), axis 1 is "y" (columns) in the chain's block
bookkeeping, float64 everywhere, masks are compared with `== 1` except in the `np.where` merge.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as dc_field

import numpy as np
from scipy.spatial import cKDTree

# --------------------------------------------------------------------------------------------------
# A1  mass-conservation residual                                    reference: Topography.py:592-600
# --------------------------------------------------------------------------------------------------


def _grad_uniform(f: np.ndarray, h: float, axis: int) -> np.ndarray:
    """np.gradient(f, h, axis=axis) for uniform spacing, edge_order=1, spelled out.

    Interior: (f[i+1]-f[i-1]) / (2.*h); first/last: one-sided difference / h.  (numpy
    lib/_function_base_impl.py `gradient`, uniform-spacing branch; used at Topography.py:595-596.)
    """
    f = np.moveaxis(np.asarray(f, dtype=np.float64), axis, 0)
    n = f.shape[0]
    if n < 2:
        raise ValueError("Shape of array too small to calculate a numerical gradient, at least 2 elements are required.")
    out = np.empty_like(f)
    out[1:-1] = (f[2:] - f[:-2]) / (2.0 * h)
    out[0] = (f[1] - f[0]) / h
    out[-1] = (f[-1] - f[-2]) / h
    return np.moveaxis(out, 0, axis)


def mass_conservation_residual(bed, surf, velx, vely, dhdt, smb, resolution):
    """div(H v) + dh/dt - SMB with H = surf - bed.  Topography.py:592-600.

    Evaluation order is part of the spec: ((dx + dy) + dhdt) - smb, axis=1 first operand.
    """
    thick = surf - bed
    dx = _grad_uniform(velx * thick, resolution, axis=1)
    dy = _grad_uniform(vely * thick, resolution, axis=0)
    return dx + dy + dhdt - smb


# --------------------------------------------------------------------------------------------------
# A2  loss                                                                reference: MCMC.py:1021-1044
# --------------------------------------------------------------------------------------------------


def masked_loss(residual, mc_region_mask, sigma_mc):
    """(total, mc, data) = nansum(res[mask==1]**2)/(2 sigma^2), data loss hard-wired to 0.  MCMC.py:1041-1044."""
    loss_mc = np.nansum(np.square(residual[mc_region_mask == 1])) / (2 * sigma_mc ** 2)
    return loss_mc + 0, loss_mc, 0


# --------------------------------------------------------------------------------------------------
# A3/A4  random-field proposal                                   reference: MCMC.py:176-254, 433-778
# --------------------------------------------------------------------------------------------------


@dataclass
class FieldParams:
    """The RandField constructor arguments (MCMC.py:462) plus block/taper setup (:524-566)."""
    range_min_x: float
    range_max_x: float
    range_min_y: float
    range_max_y: float
    scale_min: float
    scale_max: float
    nugget_max: float
    model_name: str
    isotropic: bool
    smoothness: float | None = None
    resolution: float = 1.0
    pairs: np.ndarray | None = None          # [2, n_pairs]: row 0 widths, row 1 heights (MCMC.py:568-581)
    edge_masks: list = dc_field(default_factory=list)
    logistic_param: tuple = (2.0, 0.0, 6.0, 1.0)
    max_dist: float = 1.0


def block_size_pairs(min_x, max_x, min_y, max_y, steps=5):
    """All (width, height) combinations, forced even.  MCMC.py:568-581."""
    widths = np.linspace(min_x, max_x, steps, dtype=int)
    heights = np.linspace(min_y, max_y, steps, dtype=int)
    w, h = np.meshgrid(widths, heights)
    return np.array([(w // 2 * 2).flatten(), (h // 2 * 2).flatten()])


def nearest_masked_distance(xx, yy, mask):
    """Distance of every cell to the nearest mask==True cell.  Utilities.py:21-24."""
    tree = cKDTree(np.array([xx[mask], yy[mask]]).T)
    d = tree.query(np.array([xx.ravel(), yy.ravel()]).T)[0]
    return d.reshape(xx.shape)


def logistic_weight(dist, logistic_param, max_dist):
    """L/(1+exp(-k(min(d/max,1)-x0))) - offset.  MCMC.py:615-619, 708-711."""
    L, x0, k, offset = logistic_param
    scaled = np.where(dist > max_dist, 1, (dist / max_dist))
    return L / (1 + np.exp(-k * (scaled - x0))) - offset


def edge_taper_masks(pairs, logistic_param, max_dist, resolution):
    """One [height,width] taper per block size: logistic of the distance to the block rim.  MCMC.py:583-623."""
    out = []
    for i in range(pairs.shape[1]):
        bw, bh = int(pairs[0, i]), int(pairs[1, i])
        gx, gy = np.meshgrid(range(bw), range(bh))
        gx = gx * resolution
        gy = gy * resolution
        rim = np.zeros((bh, bw))
        rim[0, :] = 1
        rim[bh - 1, :] = 1
        rim[:, 0] = 1
        rim[:, bw - 1] = 1
        d = nearest_masked_distance(gx, gy, rim == 1)
        out.append(logistic_weight(d, logistic_param, max_dist))
    return out


def crf_data_weight(xx, yy, data_mask, logistic_param, max_dist):
    """Conditioning weight, 0 at radar cells: logistic(dist to data) - min.  MCMC.py:689-714."""
    d = nearest_masked_distance(xx, yy, data_mask == 1)
    lg = logistic_weight(d, logistic_param, max_dist)
    return lg - np.min(lg)


def spectral_density(model_name, smoothness, ny, nx, res, range_x, range_y):
    """S(k) on the fftfreq grid.  MCMC.py:209-239."""
    if model_name == "Gaussian":
        len_x, len_y = range_x / np.sqrt(3), range_y / np.sqrt(3)
    elif model_name == "Exponential":
        len_x, len_y = range_x / 3.0, range_y / 3.0
    else:
        len_x, len_y = range_x / 2.0, range_y / 2.0
    kx = np.fft.fftfreq(nx, d=res) * 2 * np.pi
    ky = np.fft.fftfreq(ny, d=res) * 2 * np.pi
    kyv, kxv = np.meshgrid(ky, kx, indexing="ij")
    k = np.sqrt(kxv ** 2 + kyv ** 2) + 1e-10
    a = np.sqrt(len_x * len_y)
    if model_name == "Gaussian":
        return np.exp(-0.5 * (a * k) ** 2)
    if model_name == "Exponential":
        return 1.0 / (1.0 + (a * k) ** 2) ** 1.5
    nu = smoothness or 1.0
    const = (4 * np.pi * math.gamma(nu + 1) * (2 * nu) ** nu) / (math.gamma(nu) * a ** (2 * nu))
    kappa = 2 * nu / (a ** 2)
    return const * ((kappa + 4 * np.pi * k ** 2) ** (-nu - 1))


def field_from_draws(p: FieldParams, shape, scale, nug, range_x, range_y, z_re, z_im, z_nug):
    """Deterministic part of spectral_synthesis_field given its random draws.  MCMC.py:241-251.

    scale is the already-divided value (U/3); z_nug are unit normals (the reference draws N(0, sqrt(nug))
    = sqrt(nug)*z, numpy's `normal(loc, scale)` is loc + scale*z).
    """
    ny, nx = shape
    S = spectral_density(p.model_name, p.smoothness, ny, nx, p.resolution, range_x, range_y)
    spec = (z_re + 1j * z_im) * np.sqrt(S)
    fld = np.fft.ifft2(spec).real
    fld = (fld - np.mean(fld)) / (np.std(fld) + 1e-12)
    return fld * scale + (0 + np.sqrt(nug) * z_nug)


def draw_field_inputs(p: FieldParams, rng: np.random.Generator, shape):
    """Consume `rng` exactly like spectral_synthesis_field does (MCMC.py:200-207, 242, 250)."""
    ny, nx = shape
    scale = rng.uniform(p.scale_min, p.scale_max) / 3.0
    nug = rng.uniform(0.0, p.nugget_max)
    if not p.isotropic:
        range_x = rng.uniform(p.range_min_x, p.range_max_x)
        range_y = rng.uniform(p.range_min_y, p.range_max_y)
    else:
        range_x = range_y = rng.uniform(p.range_min_x, p.range_max_x)
    z_re = rng.normal(size=(ny, nx))
    z_im = rng.normal(size=(ny, nx))
    z_nug = rng.normal(size=(ny, nx))          # == rng.normal(0, sqrt(nug), ...)/sqrt(nug) stream-wise
    return dict(scale=scale, nug=nug, range_x=range_x, range_y=range_y, z_re=z_re, z_im=z_im, z_nug=z_nug)


def spectral_field(p: FieldParams, rng: np.random.Generator, shape):
    """spectral_synthesis_field(RF, shape, res).  MCMC.py:176-254."""
    d = draw_field_inputs(p, rng, shape)
    return field_from_draws(p, shape, **d), d


def rf_block(p: FieldParams, rng: np.random.Generator, record: dict | None = None):
    """RandField.get_rfblock (spectral branch): pick a block size, synthesise, taper.  MCMC.py:742-778."""
    pick = rng.integers(low=0, high=p.pairs.shape[1], size=1)[0]
    bw, bh = int(p.pairs[0, pick]), int(p.pairs[1, pick])
    while True:
        fld, draws = spectral_field(p, rng, (bh, bw))
        if np.sum(np.isnan(fld)) == 0:
            break
    if record is not None:
        record.update(draws, pair=int(pick))
    return fld * p.edge_masks[pick]


# --------------------------------------------------------------------------------------------------
# A6  Metropolis step                                                  reference: MCMC.py:1247-1366
# --------------------------------------------------------------------------------------------------


@dataclass
class ChainSetup:
    """Static (chain-independent) inputs of a chain_crf (MCMC.py:808-872, 950-1018, 1098-1134)."""
    surf: np.ndarray
    velx: np.ndarray
    vely: np.ndarray
    dhdt: np.ndarray
    smb: np.ndarray
    grounded_ice_mask: np.ndarray
    region_mask: np.ndarray
    mc_region_mask: np.ndarray
    resolution: float
    sigma_mc: float
    update_in_region: bool = True
    block_type: str = "CRF_weight"
    crf_weight: np.ndarray | None = None


def block_window(idx_x, idx_y, bh, bw, H, W):
    """Clip the (bh x bw) block centred on (idx_x, idx_y) to the grid.  MCMC.py:1267-1276.

    Returns grid window (x0,x1,y0,y1) and the matching window of the field (mx0,mx1,my0,my1): a block cut by
    the low edge keeps the trailing rows/cols of the field.
    """
    x0 = max(0, int(idx_x - bh / 2))
    x1 = min(H, int(idx_x + bh / 2))
    y0 = max(0, int(idx_y - bw / 2))
    y1 = min(W, int(idx_y + bw / 2))
    mx0 = max(bh - x1, 0)
    mx1 = min(H - x0, bh)
    my0 = max(bw - y1, 0)
    my1 = min(W - y0, bw)
    return (x0, x1, y0, y1), (mx0, mx1, my0, my1)


def crf_step(cs: ChainSetup, bed, mc_res, loss_prev, f, idx_x, idx_y, u):
    """One proposal + accept/reject, executed the way the reference executes it (full-grid copies).

    Returns (bed, mc_res, loss, accepted, loss_next, window).  MCMC.py:1263-1360.
    """
    H, W = bed.shape
    bh, bw = f.shape
    (x0, x1, y0, y1), (mx0, mx1, my0, my1) = block_window(idx_x, idx_y, bh, bw, H, W)
    if cs.block_type == "CRF_weight":
        perturb = f[mx0:mx1, my0:my1] * cs.crf_weight[x0:x1, y0:y1]
    else:
        perturb = f[mx0:mx1, my0:my1]
    cand = bed.copy()
    cand[x0:x1, y0:y1] = cand[x0:x1, y0:y1] + perturb
    gate = cs.region_mask if cs.update_in_region else cs.grounded_ice_mask
    cand = np.where(gate, cand, bed)

    # residual recomputed on the block padded by one cell, written back on the block only (:1292-1315)
    cx0, cx1 = max(0, x0 - 1), min(H, x1 + 1)
    cy0, cy1 = max(0, y0 - 1), min(W, y1 + 1)
    sl = (slice(cx0, cx1), slice(cy0, cy1))
    local = mass_conservation_residual(cand[sl], cs.surf[sl], cs.velx[sl], cs.vely[sl], cs.dhdt[sl], cs.smb[sl],
                                       cs.resolution)
    res_cand = mc_res.copy()
    res_cand[x0:x1, y0:y1] = local[x0 - cx0:x0 - cx0 + (x1 - x0), y0 - cy0:y0 - cy0 + (y1 - y0)]

    loss_next, _, _ = masked_loss(res_cand, cs.mc_region_mask, cs.sigma_mc)

    thick = cs.surf[x0:x1, y0:y1] - cand[x0:x1, y0:y1]
    if np.sum((thick <= 0)[gate[x0:x1, y0:y1] == 1]) > 0:
        loss_next = np.inf

    if loss_prev > loss_next:
        acc = 1
    else:
        acc = min(1, np.exp(loss_prev - loss_next))
    accepted = bool(u <= acc)
    if accepted:
        return cand.copy(), res_cand, loss_next, True, loss_next, (x0, x1, y0, y1)
    return bed, mc_res, loss_prev, False, loss_next, (x0, x1, y0, y1)


def draw_centre(cs: ChainSetup, rng: np.random.Generator, H, W):
    """Block centre: rejection-sample inside region_mask when update_in_region.  MCMC.py:1253-1261."""
    while True:
        ix = rng.integers(low=0, high=H, size=1)[0]
        iy = rng.integers(low=0, high=W, size=1)[0]
        if (not cs.update_in_region) or cs.region_mask[ix, iy] == 1:
            return int(ix), int(iy)


def run_chain(cs: ChainSetup, fp: FieldParams, initial_bed, n_iter, chain_rng, rf_rng, record=False,
              keep_beds=False):
    """chain_crf.run(n_iter, RF, only_save_last_bed=not keep_beds, plot=False, progress_bar=False).

    Same RNG consumption order as the reference: RF stream for the proposal, chain stream for centre and u.
    Returns a dict with the reference's 7 outputs (MCMC.py:1434-1443) and, if record, the per-step proposal
    inputs (f, idx_x, idx_y, u) that make the step deterministic.
    """
    H, W = initial_bed.shape
    bed = initial_bed
    gate = cs.region_mask if cs.update_in_region else cs.grounded_ice_mask
    mc_res = mass_conservation_residual(bed, cs.surf, cs.velx, cs.vely, cs.dhdt, cs.smb, cs.resolution)
    loss_prev, _, _ = masked_loss(mc_res, cs.mc_region_mask, cs.sigma_mc)

    loss_cache = np.zeros(n_iter)
    step_cache = np.zeros(n_iter)
    blocks_cache = np.full((n_iter, 4), np.nan)
    resampled = np.zeros((H, W))
    loss_cache[0] = loss_prev
    beds = np.zeros((n_iter, H, W)) if keep_beds else None
    if keep_beds:
        beds[0] = bed
    tape = [] if record else None

    for i in range(1, n_iter):
        rec = {} if record else None
        f = rf_block(fp, rf_rng, rec)
        ix, iy = draw_centre(cs, chain_rng, H, W)
        blocks_cache[i, :] = [ix, iy, f.shape[0], f.shape[1]]
        u = chain_rng.random()
        bed, mc_res, loss_prev, ok, loss_next, (x0, x1, y0, y1) = crf_step(cs, bed, mc_res, loss_prev, f, ix, iy, u)
        loss_cache[i] = loss_prev
        step_cache[i] = ok
        if ok:
            resampled[x0:x1, y0:y1] += gate[x0:x1, y0:y1]
        if keep_beds:
            beds[i] = bed
        if record:
            rec.update(f=f, idx_x=ix, idx_y=iy, u=u, loss_next=loss_next)
            tape.append(rec)

    return dict(bed=bed, beds=beds, loss_mc=loss_cache.copy(), loss_data=np.zeros(n_iter), loss=loss_cache,
                steps=step_cache, resampled_times=resampled, blocks=blocks_cache, mc_res=mc_res, tape=tape)


# --------------------------------------------------------------------------------------------------
# helpers to build oracle inputs from the synthetic recipe
# --------------------------------------------------------------------------------------------------


def setup_from_grids(g: dict, sigma_mc=5.0, logistic=(2.0, 0.0, 6.0, 1.0), max_dist=30e3, blocks=(50, 80, 50, 80),
                     rf_kw: dict | None = None, update_in_region=True, block_type="CRF_weight", steps=5):
    """Build (ChainSetup, FieldParams) the way the tutorial configures chain_crf + RandField."""
    from mcmc_gpu_b200 import synthetic as syn
    kw = dict(syn.RF_KW if rf_kw is None else rf_kw)
    pairs = block_size_pairs(*blocks, steps=steps)
    fp = FieldParams(**kw, resolution=g["resolution"], pairs=pairs, logistic_param=tuple(logistic), max_dist=max_dist)
    fp.edge_masks = edge_taper_masks(pairs, fp.logistic_param, max_dist, g["resolution"])
    region = g["highvel_mask"] if update_in_region else np.full(g["xx"].shape, 1)
    cs = ChainSetup(surf=g["surf"], velx=g["velx"], vely=g["vely"], dhdt=g["dhdt"], smb=g["smb"],
                    grounded_ice_mask=g["grounded_ice_mask"], region_mask=region, mc_region_mask=region,
                    resolution=g["resolution"], sigma_mc=sigma_mc, update_in_region=update_in_region,
                    block_type=block_type,
                    crf_weight=crf_data_weight(g["xx"], g["yy"], g["data_mask"], fp.logistic_param, max_dist))
    return cs, fp
