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
