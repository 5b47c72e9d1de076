"""Host-side mirror of the reference's physics helper (gstatsMCMC/Topography.py:592-612) over kernel K2.

Only the hot-path functions are provided; the data loaders / gridding / geoid helpers of the reference's
Topography.py are preprocessing I/O and out of scope.
"""
from __future__ import annotations

import numpy as np

from ._lib import Context, GmcShapeError

_ctx_cache: dict = {}


def _context(H, W, C, device):
    key = (H, W, C, str(device))
    ctx = _ctx_cache.get(key)
    if ctx is None:
        if len(_ctx_cache) > 8:
            for old in list(_ctx_cache.values()):
                old.close()
            _ctx_cache.clear()
        ctx = _ctx_cache[key] = Context(H, W, C, device)
    return ctx


def _residual(bed, statics, resolution, device=None):
    import torch
    bed_t = bed if isinstance(bed, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(bed, dtype=np.float64))
    squeeze = bed_t.dim() == 2
    if squeeze:
        bed_t = bed_t[None]
    if bed_t.dim() != 3:
        raise GmcShapeError(f"bed must be [H,W] or [C,H,W], got {tuple(bed_t.shape)}")
    C, H, W = bed_t.shape
    ctx = _context(H, W, C, device if device is not None else (bed_t.device if bed_t.is_cuda else None))
    host = [s.detach().cpu().numpy() if isinstance(s, torch.Tensor) else np.asarray(s) for s in statics]
    for s in host:
        if s.shape != (H, W):
            raise GmcShapeError("operands could not be broadcast together: every field must have the bed's shape")
    ones = np.ones((H, W), dtype=np.uint8)
    ctx.set_static(*host, ones, ones, None, None, float(resolution), 1.0)
    bed_d = bed_t.to(ctx.device, dtype=torch.float64).contiguous()
    out = torch.empty_like(bed_d)
    ctx.residual(bed_d, out)
    return out[0] if squeeze else out


def get_mass_conservation_residual(bed, surf, velx, vely, dhdt, smb, resolution):
    """div(H v) + dh/dt - SMB, H = surf - bed, with np.gradient's stencil (reference Topography.py:592-600).

    numpy in, numpy out; `bed` may also be a [C,H,W] stack (all chains share the other fields).  Bit-identical to the
    reference: per-operation rounding, true division, ((dx + dy) + dhdt) - smb.
    """
    return _residual(bed, (surf, velx, vely, dhdt, smb), resolution).cpu().numpy()


def get_mass_conservation_residual_tensor(bed, surf, velx, vely, dhdt, smb, resolution):
    """torch.Tensor in, CUDA float64 torch.Tensor out (reference Topography.py:602-612 is float32 torch.gradient)."""
    return _residual(bed, (surf, velx, vely, dhdt, smb), float(resolution))


def get_highvel_boundary(velx, vely, velmag_threshold, grounded_ice_mask, ocean_mask, distance_max, xx, yy, smooth_mode=10):
    """High-velocity update region (reference Topography.py:546-571): grounded cells at least `velmag_threshold` fast,
    plus ocean, smoothed by a `smooth_mode` mode filter, then grown by `distance_max` inside the grounded mask.

    The reference runs PIL's ModeFilter and an O(N^2) Python double loop for the nearest-region distance; here both are
    kernels (`mode_filter_kernel`, `min_dist_kernel`).  The distance is the same correctly rounded expression
    sqrt((y-y')^2 + (x-x')^2) minimised over the region, so the returned mask equals the reference's exactly.
    """
    import torch
    from . import _lib
    from .Utilities import min_dist_from_mask
    velx, vely = np.asarray(velx, dtype=np.float64), np.asarray(vely, dtype=np.float64)
    grounded = np.asarray(grounded_ice_mask)
    mask = (grounded.astype(bool)) & (np.sqrt(velx ** 2 + vely ** 2) >= velmag_threshold)
    mask = mask | np.asarray(ocean_mask).astype(bool)
    dev = _lib.require_cuda()
    lib = _lib.load()
    H, W = mask.shape
    img = torch.as_tensor((mask * 255).astype(np.uint8)).to(dev)
    out = torch.empty_like(img)
    _lib.check(lib.gmc_mode_filter_binary(dev.index, img.data_ptr(), out.data_ptr(), H, W, int(smooth_mode),
                                          torch.cuda.current_stream().cuda_stream))
    mask_mat = (out.cpu().numpy() // 255).astype(int)
    hard = (mask_mat == 1) & (grounded == 1)
    if not hard.any():                       # the reference's nanmin over an all-NaN map is NaN: nothing is "close"
        return np.zeros(mask.shape, dtype=grounded.dtype) & grounded
    mask_dist = min_dist_from_mask(np.asarray(xx, dtype=np.float64), np.asarray(yy, dtype=np.float64), hard)
    return (mask_dist < distance_max) & grounded
