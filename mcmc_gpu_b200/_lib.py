"""ctypes binding of libgmc.so (include/gmc.h) and a thin Context wrapper over torch device buffers.

PyTorch is used only to own device memory and streams; every computation is a libgmc kernel.  There is no CPU
fallback: if the shared library is missing or no CUDA device is usable the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgmc.so")

GMC_OK, GMC_EINVAL, GMC_ESHAPE, GMC_ECUDA, GMC_ENCCL, GMC_EUNSUPPORTED, GMC_ESTATE = 0, -1, -2, -3, -4, -5, -6
MODEL_IDS = {"Gaussian": 0, "Exponential": 1, "Matern": 2}


class GmcError(RuntimeError):
    """CUDA / NCCL / call-order failure reported by libgmc."""


class GmcShapeError(Exception):
    """Shape mismatch (the reference raises a bare Exception for these, MCMC.py:844-845)."""


_c_p = C.c_void_p
_i32, _i64, _u64, _f64 = C.c_int32, C.c_int64, C.c_uint64, C.c_double

# name -> (restype, argtypes); kept in one table so tests can check it against include/gmc.h
PROTOTYPES = {
    "gmc_last_error": (C.c_char_p, []),
    "gmc_version": (C.c_int, []),
    "gmc_create": (C.c_int, [C.POINTER(_c_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "gmc_destroy": (C.c_int, [_c_p]),
    "gmc_set_static": (C.c_int, [_c_p] + [_c_p] * 5 + [_c_p, _c_p, _c_p, _i64, _c_p, _f64, _f64]),
    "gmc_set_field_model": (C.c_int, [_c_p, C.c_int, _f64, C.c_int] + [_f64] * 7),
    "gmc_set_generation_method": (C.c_int, [_c_p, C.c_int, C.c_int]),
    "gmc_set_blocks": (C.c_int, [_c_p, C.c_int, _c_p, _c_p, _c_p, _c_p, _f64]),
    "gmc_residual": (C.c_int, [_c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_residual_loss": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_residual_loss_range": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, C.c_int, C.c_int, _c_p]),
    "gmc_loss": (C.c_int, [_c_p, _c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_field_spectral": (C.c_int, [_c_p, C.c_int] + [_c_p] * 9 + [_u64, C.c_int, _c_p, _i64, _c_p]),
    "gmc_field_randmeth": (C.c_int, [_c_p, C.c_int] + [_c_p] * 6 + [C.c_int, _c_p, _c_p, _c_p, _u64, C.c_int, _c_p, _i64, _c_p]),
    "gmc_step_injected": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _i64, _c_p, _c_p, _c_p, C.c_int, C.c_int,
                                    _c_p, _c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_run": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _u64, C.c_int, _c_p, _c_p, _c_p, _i64, _i64, _c_p,
                          C.c_int, C.c_int, _c_p]),
    "gmc_ensemble_moments": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_allreduce_moments": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "gmc_sgs_setup": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, C.c_int, _c_p, _c_p, C.c_int, C.c_int, C.c_int, C.c_int, _c_p, _f64,
                                C.c_int, C.c_int, C.c_int, C.c_int]),
    "gmc_sgs_transform": (C.c_int, [_c_p, _c_p, _c_p, _i64, C.c_int, _c_p]),
    "gmc_sgs_init": (C.c_int, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_sgs_step_injected": (C.c_int, [_c_p] * 10 + [_i64] + [_c_p] * 6 + [C.c_int, _c_p]),
    "gmc_sgs_run": (C.c_int, [_c_p] * 7 + [_u64, C.c_int, _c_p, _c_p, _c_p, _i64, _i64, _c_p, _c_p, C.c_int, _c_p]),
    "gmc_min_dist": (C.c_int, [C.c_int, _c_p, _c_p, _i64, _c_p, _c_p, _i64, _c_p, _c_p]),
    "gmc_nst_transform": (C.c_int, [C.c_int, _c_p, _c_p, C.c_int, _c_p, _c_p, _i64, C.c_int, _c_p]),
    "gmc_sgs_grid_solve": (C.c_int, [C.c_int, C.c_int, C.c_int, _c_p, _c_p, _i64, C.c_int, _c_p, _c_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, _c_p, _f64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "gmc_sgs_grid_values": (C.c_int, [C.c_int, C.c_int, C.c_int, _c_p, _c_p, _i64, C.c_int] + [_c_p] * 8),
    "gmc_mode_filter_binary": (C.c_int, [C.c_int, _c_p, _c_p, C.c_int, C.c_int, C.c_int, _c_p]),
    "gmc_launch_count": (_i64, [_c_p]),
    "gmc_step_kernel_info": (C.c_int, [_c_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "gmc_check": (C.c_int, [_c_p, C.c_int]),
    "gmc_set_step_cta": (C.c_int, [_c_p, C.c_int]),
    "gmc_debug_fp64_peak": (C.c_int, [_c_p, C.POINTER(_f64)]),
    "gmc_debug_phase_timing": (C.c_int, [_c_p, C.c_int, _c_p]),
    "gmc_debug_div_check": (C.c_int, [_c_p, _c_p, _i64, _f64, C.POINTER(_i64)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libgmc.so (built in-tree by `make -C mcmc_gpu_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). mcmc_gpu_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == GMC_OK:
        return
    msg = load().gmc_last_error().decode("utf-8", "replace")
    if rc == GMC_EINVAL:
        raise ValueError(msg)
    if rc == GMC_ESHAPE:
        raise GmcShapeError(msg)
    if rc == GMC_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise GmcError(f"[{rc}] {msg}")


def _ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        if not t.flags["C_CONTIGUOUS"]:
            raise ValueError("numpy arrays handed to libgmc must be C-contiguous")
        return t.ctypes.data
    if not t.is_contiguous():
        raise ValueError("tensors handed to libgmc must be contiguous")
    return t.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def require_cuda(device=None):
    """Return a torch.device for a usable CUDA device or raise (no CPU path exists)."""
    import torch
    if not torch.cuda.is_available():
        raise GmcError("mcmc_gpu_b200 needs a CUDA device (B200, sm_100a); torch.cuda.is_available() is False and "
                       "there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise GmcError(f"mcmc_gpu_b200 only runs on CUDA devices, got {dev}")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class Context:
    """Owns one gmc_ctx: a grid shape on one device shared by up to `max_chains` chains."""

    def __init__(self, H: int, W: int, max_chains: int, device=None):
        import torch
        self.lib = load()
        self.device = require_cuda(device)
        self.H, self.W, self.max_chains = int(H), int(W), int(max_chains)
        h = _c_p()
        with torch.cuda.device(self.device):
            check(self.lib.gmc_create(C.byref(h), self.device.index, self.H, self.W, self.max_chains))
        self._h = h
        self.max_h = self.max_w = 0
        self.pairs = None

    def close(self):
        if getattr(self, "_h", None):
            self.lib.gmc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup -------------------------------------------------------------------------------------------
    def set_static(self, surf, velx, vely, dhdt, smb, gate_mask, mc_mask, centre_cells, crf_weight, resolution,
                   sigma_mc):
        def f64(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != (self.H, self.W):
                raise GmcShapeError(f"static field has shape {a.shape}, expected {(self.H, self.W)}")
            return a

        def u8(a):
            a = np.asarray(a)
            if a.shape != (self.H, self.W):
                raise GmcShapeError(f"mask has shape {a.shape}, expected {(self.H, self.W)}")
            return np.ascontiguousarray(a != 0, dtype=np.uint8)

        arrs = [f64(x) for x in (surf, velx, vely, dhdt, smb)]
        g, m = u8(gate_mask), u8(mc_mask)
        cc = None if centre_cells is None else np.ascontiguousarray(centre_cells, dtype=np.int32)
        w = None if crf_weight is None else f64(crf_weight)
        check(self.lib.gmc_set_static(self._h, *[_ptr(a) for a in arrs], _ptr(g), _ptr(m), _ptr(cc),
                                      0 if cc is None else cc.size, _ptr(w), float(resolution), float(sigma_mc)))

    def set_field_model(self, model_name, smoothness, isotropic, range_min_x, range_max_x, range_min_y, range_max_y,
                        scale_min, scale_max, nugget_max):
        if model_name not in MODEL_IDS:
            raise Exception("please put in a valid model_name, including Gaussian, Exponential, and Matern")
        check(self.lib.gmc_set_field_model(self._h, MODEL_IDS[model_name], float(smoothness or 0.0), int(bool(isotropic)),
                                           float(range_min_x), float(range_max_x), float(range_min_y),
                                           float(range_max_y), float(scale_min), float(scale_max), float(nugget_max)))

    def set_generation_method(self, spectral, n_modes=1000):
        """RandField.set_generation_method (MCMC.py:514): False selects the randomization-method proposal (A5)."""
        check(self.lib.gmc_set_generation_method(self._h, int(bool(spectral)), int(n_modes)))

    def set_blocks(self, pairs, edge_masks, field_resolution):
        pairs = np.asarray(pairs)
        n = pairs.shape[1]
        pw = np.ascontiguousarray(pairs[0], dtype=np.int32)
        ph = np.ascontiguousarray(pairs[1], dtype=np.int32)
        offs = np.zeros(n, dtype=np.int64)
        flat = []
        o = 0
        for i in range(n):
            m = np.ascontiguousarray(edge_masks[i], dtype=np.float64)
            if m.shape != (ph[i], pw[i]):
                raise GmcShapeError(f"edge mask {i} has shape {m.shape}, expected {(ph[i], pw[i])}")
            offs[i] = o
            o += m.size
            flat.append(m.ravel())
        flat = np.ascontiguousarray(np.concatenate(flat))
        check(self.lib.gmc_set_blocks(self._h, n, _ptr(pw), _ptr(ph), _ptr(flat), _ptr(offs), float(field_resolution)))
        self.pairs = pairs.copy()
        self.max_h, self.max_w = int(ph.max()), int(pw.max())

    # ---- compute (all tensors: torch, on self.device, contiguous) -------------------------------------------
    def residual(self, bed, res_out):
        check(self.lib.gmc_residual(self._h, _ptr(bed), _ptr(res_out), bed.shape[0], _stream()))

    def residual_loss(self, bed, res_out, loss_out, ssq_out=None):
        check(self.lib.gmc_residual_loss(self._h, _ptr(bed), _ptr(res_out), _ptr(loss_out), _ptr(ssq_out), bed.shape[0],
                                         _stream()))

    def residual_loss_range(self, bed, res_out, loss_out, ssq_out, chain0):
        check(self.lib.gmc_residual_loss_range(self._h, _ptr(bed), _ptr(res_out), _ptr(loss_out), _ptr(ssq_out), bed.shape[0],
                                               int(chain0), _stream()))

    def loss(self, res, loss_out, ssq_out=None):
        check(self.lib.gmc_loss(self._h, _ptr(res), _ptr(loss_out), _ptr(ssq_out), res.shape[0], _stream()))

    def field_spectral(self, pair, scale, nug, range_x, range_y, f_out, z_re=None, z_im=None, z_nug=None, seeds=None,
                       iteration=0, apply_taper=True):
        check(self.lib.gmc_field_spectral(self._h, pair.shape[0], _ptr(pair), _ptr(scale), _ptr(nug), _ptr(range_x),
                                          _ptr(range_y), _ptr(z_re), _ptr(z_im), _ptr(z_nug), _ptr(seeds),
                                          int(iteration), int(bool(apply_taper)), _ptr(f_out), f_out.shape[1], _stream()))

    def field_randmeth(self, pair, scale, nug, range_x, range_y, angle_deg, f_out, n_modes=1000, modes=None, z_nug=None,
                       seeds=None, iteration=0, apply_taper=True):
        check(self.lib.gmc_field_randmeth(self._h, pair.shape[0], _ptr(pair), _ptr(scale), _ptr(nug), _ptr(range_x),
                                          _ptr(range_y), _ptr(angle_deg), int(n_modes), _ptr(modes), _ptr(z_nug),
                                          _ptr(seeds), int(iteration), int(bool(apply_taper)), _ptr(f_out),
                                          f_out.shape[1], _stream()))

    def step_injected(self, bed, mcres, ssq, f, hw, centre, u, hmax, wmax, accepted_out, loss_out, loss_next_out=None,
                      resampled=None):
        check(self.lib.gmc_step_injected(self._h, _ptr(bed), _ptr(mcres), _ptr(ssq), _ptr(f), f.shape[1], _ptr(hw),
                                         _ptr(centre), _ptr(u), int(hmax), int(wmax), _ptr(accepted_out), _ptr(loss_out),
                                         _ptr(loss_next_out), _ptr(resampled), bed.shape[0], _stream()))

    def run(self, bed, mcres, ssq, seeds, iter0, n_steps, loss_cache=None, step_cache=None, blocks_cache=None,
            cache_offset=0, resampled=None, resync_every=0):
        stride = 0
        for t in (loss_cache, step_cache, blocks_cache):
            if t is not None:
                stride = t.shape[1]
        check(self.lib.gmc_run(self._h, _ptr(bed), _ptr(mcres), _ptr(ssq), _ptr(seeds), int(iter0), int(n_steps),
                               _ptr(loss_cache), _ptr(step_cache), _ptr(blocks_cache), stride, int(cache_offset),
                               _ptr(resampled), int(resync_every), bed.shape[0], _stream()))

    # ---- SGS chain -------------------------------------------------------------------------------------------
    def sgs_setup(self, trend, zcond, grounded, quantiles, references, oct_off, oct_cnt, hw, num_points, lut, sill, blocks):
        f64 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)          # noqa: E731
        trend, zcond, quantiles, references, lut = f64(trend), f64(zcond), f64(quantiles), f64(references), f64(lut)
        grounded = np.ascontiguousarray(np.asarray(grounded) == 1, dtype=np.uint8)
        oct_off = np.ascontiguousarray(oct_off, dtype=np.int16)
        oct_cnt = np.ascontiguousarray(np.atleast_2d(oct_cnt), dtype=np.int32)          # [n_levels][8]
        if oct_cnt.shape[1] != 8:
            raise GmcShapeError(f"octant counts have shape {oct_cnt.shape}, expected [n_levels, 8]")
        check(self.lib.gmc_sgs_setup(self._h, _ptr(trend), _ptr(zcond), _ptr(grounded), _ptr(quantiles), _ptr(references),
                                     0 if quantiles is None else quantiles.size, _ptr(oct_off), _ptr(oct_cnt), oct_cnt.shape[0],
                                     oct_off.shape[1],
                                     int(hw), int(num_points), _ptr(lut), float(sill), *[int(b) for b in blocks]))

    def sgs_transform(self, x, out, inverse=False):
        check(self.lib.gmc_sgs_transform(self._h, _ptr(x), _ptr(out), x.numel(), int(bool(inverse)), _stream()))

    def sgs_init(self, bed, bedc, z, mcres, ssq, nviol, scratch):
        check(self.lib.gmc_sgs_init(self._h, _ptr(bed), _ptr(bedc), _ptr(z), _ptr(mcres), _ptr(ssq), _ptr(nviol), _ptr(scratch),
                                    bed.shape[0], _stream()))

    def sgs_step_injected(self, bedc, z, mcres, ssq, nviol, centre, bs, path, znorm, u, accepted, loss, loss_next=None,
                          resampled=None, err=None):
        check(self.lib.gmc_sgs_step_injected(self._h, _ptr(bedc), _ptr(z), _ptr(mcres), _ptr(ssq), _ptr(nviol), _ptr(centre),
                                             _ptr(bs), _ptr(path), _ptr(znorm), path.shape[1], _ptr(u), _ptr(accepted), _ptr(loss),
                                             _ptr(loss_next), _ptr(resampled), _ptr(err), bedc.shape[0], _stream()))

    def sgs_run(self, bedc, z, mcres, ssq, nviol, seeds, iter0, n_steps, loss_cache=None, step_cache=None, blocks_cache=None,
                cache_offset=0, resampled=None, err=None):
        stride = 0
        for t in (loss_cache, step_cache, blocks_cache):
            if t is not None:
                stride = t.shape[1]
        check(self.lib.gmc_sgs_run(self._h, _ptr(bedc), _ptr(z), _ptr(mcres), _ptr(ssq), _ptr(nviol), _ptr(seeds), int(iter0),
                                   int(n_steps), _ptr(loss_cache), _ptr(step_cache), _ptr(blocks_cache), stride, int(cache_offset),
                                   _ptr(resampled), _ptr(err), bedc.shape[0], _stream()))

    def ensemble_moments(self, bed, ref_bed, sum_out, sumsq_out):
        check(self.lib.gmc_ensemble_moments(self._h, _ptr(bed), _ptr(ref_bed), _ptr(sum_out), _ptr(sumsq_out),
                                            bed.shape[0], _stream()))

    def launch_count(self) -> int:
        return int(self.lib.gmc_launch_count(self._h))

    def phase_timing(self, enable=True, read=False):
        """Debug: enable/zero or read the per-phase cycle counters of the fused step kernel."""
        out = np.zeros(8, dtype=np.int64) if read else None
        check(self.lib.gmc_debug_phase_timing(self._h, int(bool(enable)), _ptr(out)))
        return out

    def div_check(self, x, divisor) -> int:
        """Debug: number of elements of the CUDA tensor x for which x/divisor by reciprocal+FMA differs from IEEE division."""
        n = _i64(0)
        check(self.lib.gmc_debug_div_check(self._h, _ptr(x), x.numel(), float(divisor), C.byref(n)))
        return int(n.value)

    def step_kernel_info(self, n_chains=None):
        """Launch shape of the fused step kernel; with n_chains, the CTA size gmc_run picks for that many chains (512
        threads, one CTA per SM, when there are no more chains than SMs; else 256 threads, two CTAs per SM)."""
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.gmc_step_kernel_info(self._h, C.byref(a), C.byref(b), C.byref(c)))
        info = dict(smem_bytes=a.value, threads=b.value, ctas_per_sm=c.value)
        if n_chains is not None:
            import torch
            sms = torch.cuda.get_device_properties(self.device).multi_processor_count
            slots = c.value * sms
            mode = getattr(self, "_cta_mode", 0)
            split = mode == 3 and 2 * n_chains < slots
            if os.environ.get("GMC_STEP_SPLIT") is not None:
                split = os.environ["GMC_STEP_SPLIT"][:1] == "1" and 2 * n_chains < slots
            wide = (not split) and ((mode == 0 and n_chains <= sms) or mode == 2)
            if os.environ.get("GMC_STEP_WIDE") is not None:
                wide = (not split) and os.environ["GMC_STEP_WIDE"][:1] == "1" and n_chains <= slots
            if split:
                info.update(mode="split: field producer CTA + Metropolis tail CTA per chain", ctas_per_chain=2)
            elif wide:
                info.update(threads=512, ctas_per_sm=1, mode="wide")
            else:
                info.update(mode="fused")
        return info

    def set_step_cta(self, mode):
        """'auto' | 'narrow' | 'wide' | 'split' (gmc_set_step_cta): callers whose launches share the GPU select 'narrow'."""
        self._cta_mode = {"auto": 0, "narrow": 1, "wide": 2, "split": 3}[mode]
        check(self.lib.gmc_set_step_cta(self._h, self._cta_mode))

    def check(self):
        """Synchronise and raise GmcError if a kernel gave up a bounded in-kernel wait (chain state of that launch invalid)."""
        check(self.lib.gmc_check(self._h, 1))

    def check_flag(self):
        """The same without synchronising: for callers that have just waited for the streams / copies they queued."""
        check(self.lib.gmc_check(self._h, 0))

    def fp64_peak_tflops(self) -> float:
        """Measured FP64 FMA rate of this GPU (register-resident DFMA loop), TFLOP/s."""
        v = _f64(0.0)
        check(self.lib.gmc_debug_fp64_peak(self._h, C.byref(v)))
        return float(v.value)

    def stencil_kernel_name(self) -> str:
        """Which full-grid stencil kernel gmc_residual / gmc_residual_loss launch for this grid."""
        if self.W % 2 == 0 and not os.environ.get("GMC_RS_LEGACY"):
            return "residual_tma_kernel (TMA tensor tiles on an mbarrier ring)"
        return "residual_kernel (cp.async ring; odd W or GMC_RS_LEGACY)"
