"""Host-side tables of the SGS kernel (one-time setup, numpy/scipy): the octant search stencil and the covariance
look-up table over integer cell offsets.  Mirrors gstatsim_custom/neighbors.py:4-83, _krige.py:83-143 and
covariance.py:4-29 of the reference; the per-node work that consumes these tables runs on the GPU (csrc/sgs.cu)."""
from __future__ import annotations

import math

import numpy as np
from scipy.special import gamma, kv


def grid_steps(xx, yy):
    """(dx, dy) of a regular grid; the SGS kernel indexes neighbours by integer offsets, so it needs one."""
    dx = float(xx[0, 1] - xx[0, 0])
    dy = float(yy[1, 0] - yy[0, 0])
    j = np.arange(xx.shape[1])
    i = np.arange(xx.shape[0])
    if not (np.allclose(xx, xx[0, 0] + j[None, :] * dx, rtol=0, atol=1e-6 * abs(dx)) and
            np.allclose(yy, yy[0, 0] + i[:, None] * dy, rtol=0, atol=1e-6 * abs(dy))):
        raise NotImplementedError("the SGS kernel needs a regular grid (xx varying along columns, yy along rows)")
    return dx, dy


def octant_stencil(dx, dy, radius, hw_cap=None):
    """For octant b = -4..3 (neighbors.py:52-60) the window offsets (di, dj) with distance < radius, ordered by
    (distance, di, dj) — i.e. np.argsort(kind='stable') over the row-major window, which is how ties are broken.
    hw_cap bounds the window half width (offsets that can never fall inside the grid need not be listed).
    Returns (offsets int16 [8, lmax, 2], counts int32 [8], hw)."""
    hw = math.ceil(radius / abs(dx))                       # make_circle_stencil: ncells = ceil(rad/dx) (neighbors.py:77-79)
    if hw_cap is not None:
        hw = max(1, min(hw, int(hw_cap)))
    di, dj = np.meshgrid(np.arange(-hw, hw + 1), np.arange(-hw, hw + 1), indexing="ij")
    # reference: distances = sqrt((x0 - x)^2 + (y0 - y)^2), angles = arctan2(y0 - y, x0 - x)   (neighbors.py:48-49)
    ddx, ddy = 0.0 - dj * dx, 0.0 - di * dy      # x0 - x with x0 = 0: keeps +0.0 (arctan2(-0.0, -1) would be -pi, not +pi)
    dist = np.sqrt(ddx ** 2 + ddy ** 2)
    ang = np.arctan2(ddy, ddx)
    lists = []
    for b in range(-4, 4, 1):
        m = (dist < radius) & (ang > b / 4 * np.pi) & (ang <= (b + 1) / 4 * np.pi)
        order = np.argsort(dist[m], kind="stable")         # row-major masked order is (di, dj) ascending
        lists.append(np.stack([di[m][order], dj[m][order]], axis=1))
    lmax = max(1, max(len(x) for x in lists))
    off = np.zeros((8, lmax, 2), dtype=np.int16)
    cnt = np.zeros(8, dtype=np.int32)
    for o, x in enumerate(lists):
        off[o, :len(x)] = x
        cnt[o] = len(x)
    return off, cnt, hw


MAX_LEVELS = 4            # search radii: radius, +100 km, +200 km, +300 km
MAX_HALF_WIDTH = 700      # cells; bounds the offset lists and the covariance table ((4 hw + 1)^2 doubles)


def search_levels(dx, dy, H, W, radius):
    """Octant search tables for `radius` and its widened versions: a node that finds no data within `radius` searches
    again with radius + 100 km (interpolate.py:149-155).  The octant lists are sorted by distance, so every radius level is
    a prefix of the lists built for the widest one; levels stop once the radius covers the grid diagonal (or after
    MAX_LEVELS, or when the tables would get unreasonably large).
    Returns (offsets int16 [8, lmax, 2], counts int32 [levels, 8], half-width of the widest window, radii)."""
    diag = float(np.hypot(abs(dx) * W, abs(dy) * H))
    radii = [float(radius)]
    while radii[-1] < diag and len(radii) < MAX_LEVELS and (radii[-1] + 100e3) / min(abs(dx), abs(dy)) <= MAX_HALF_WIDTH:
        radii.append(radii[-1] + 100e3)
    off, cnt_max, hw = octant_stencil(dx, dy, radii[-1], hw_cap=max(H, W) - 1 if len(radii) > 1 else None)
    dist = np.sqrt((off[..., 1] * dx) ** 2 + (off[..., 0] * dy) ** 2)         # [8, lmax], the reference's expression
    valid = np.arange(off.shape[1])[None, :] < cnt_max[:, None]
    cnt = np.stack([((dist < r) & valid).sum(1) for r in radii]).astype(np.int32)   # [levels, 8]
    return off, cnt, hw, radii


def covariance_model(vtype, h, sill, nugget, s=None):
    """Covariance of the range-normalised lag h (covariance.py:4-22), including the spherical model's quirk."""
    h = np.array(h, dtype=np.float64, copy=True)
    vt = vtype.lower()
    if vt == "exponential":
        return (sill - nugget) * np.exp(-3 * h)
    if vt == "gaussian":
        return (sill - nugget) * np.exp(-3 * np.square(h))
    if vt == "spherical":
        c = sill - nugget - 1.5 * h + 0.5 * np.power(h, 3)
        c[h > 1] = sill - 1
        return c
    if vt == "matern":
        scale = 0.45246434 * np.exp(-0.70449189 * s) + 1.7863836
        h[h == 0.0] = 1e-8
        c = (sill - nugget) * 2 / gamma(s) * np.power(scale * h * np.sqrt(s), s) * kv(s, 2 * scale * h * np.sqrt(s))
        c[np.isnan(c)] = sill - nugget
        return c
    raise ValueError("vario_type argument should be one of the following: Gaussian, Exponential, Spherical, or Matern")


def covariance_lut(dx, dy, hw, vario):
    """cov[(di + 2hw), (dj + 2hw)] for cell offsets di, dj in [-2hw, 2hw]: |offset @ R| through the model
    (make_rotation_matrix / make_sigma / make_rho, _krige.py:83-143)."""
    theta = (vario["azimuth"] / 180.0) * np.pi
    R = np.dot(np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]]),
               np.array([[1 / vario["major_range"], 0], [0, 1 / vario["minor_range"]]]))
    di, dj = np.meshgrid(np.arange(-2 * hw, 2 * hw + 1), np.arange(-2 * hw, 2 * hw + 1), indexing="ij")
    v = np.stack([(dj * dx).ravel(), (di * dy).ravel()], axis=1) @ R
    h = np.sqrt(np.square(v).sum(axis=1))
    return covariance_model(vario["vtype"], h, vario["sill"], vario["nugget"], vario.get("s")).reshape(di.shape)
