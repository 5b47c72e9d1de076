"""Host-side mirror of the reference's chain API (gstatsMCMC/MCMC.py) over libgmc's CUDA kernels.

Same class names, constructor/setter signatures, return tuples and error behaviour as the reference for the
large-scale chain path (`RandField` MCMC.py:433, `chain` :780, `chain_crf` :1083, `spectral_synthesis_field` :176);
every array operation of the hot loop runs on the GPU through the C ABI in include/gmc.h.  Differences, all forced by
the design: random numbers come from per-chain counter-based Philox streams on the device instead of
numpy PCG64 on the host (statistically equivalent; bit parity is defined under injected proposals, `replay=`), and
`run_many` / `ChainBatch` add the batched many-chain form that the reference gets from one process per chain.
"""
from __future__ import annotations

import numbers
import os
import sys
import time

import numpy as np

from . import _lib
from ._lib import Context, GmcError, GmcShapeError  # noqa: F401

_MASK64 = (1 << 64) - 1


def _splitmix64(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def philox_key(chain_seed: int, rf_seed: int | None = None) -> int:
    """64-bit Philox key of one chain.  The reference seeds chain.rng and RF.rng with the same integer in its
    drivers (MCMC.py:371,397); when they differ both enter the key."""
    k = _splitmix64(int(chain_seed) & _MASK64)
    if rf_seed is not None and int(rf_seed) != int(chain_seed):
        k ^= _splitmix64((int(rf_seed) & _MASK64) ^ 0xA5A5A5A5A5A5A5A5)
    return k


def keys_tensor(keys, device):
    """uint64 Philox keys -> int64 torch tensor with the same bit patterns (torch has no uint64 arithmetic)."""
    import torch
    return torch.as_tensor(np.array([int(k) & _MASK64 for k in keys], dtype=np.uint64).view(np.int64)).to(device)


def _seed_to_int(rng_seed, what):
    """None | int | numpy Generator -> integer seed (same accepted types as MCMC.py:484-491, 1057-1064)."""
    if rng_seed is None:
        return int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0]), np.random.default_rng()
    if isinstance(rng_seed, numbers.Integral) and not isinstance(rng_seed, bool):
        return int(rng_seed), np.random.default_rng(seed=int(rng_seed))
    if isinstance(rng_seed, np.random.Generator):
        # derive the device key from the generator without disturbing it
        state = rng_seed.bit_generator.state
        probe = np.random.Generator(type(rng_seed.bit_generator)())
        probe.bit_generator.state = state
        return int(probe.integers(0, 2 ** 63 - 1)), rng_seed
    raise ValueError("Seed should be an integer, a NumPy random Generator, or None")


# ------------------------------------------------------------------------------------------------------------------
# RandField                                                                                   reference MCMC.py:433-778
# ------------------------------------------------------------------------------------------------------------------
class RandField:
    """Parameters of the block random-field proposal (reference MCMC.py:433).  Holds configuration only; fields
    are synthesised on the GPU (K1) inside `chain_crf.run`, or one at a time through `get_rfblock()`."""

    def __init__(self, range_min_x, range_max_x, range_min_y, range_max_y, scale_min, scale_max, nugget_max, model_name,
                 isotropic, smoothness=None, rng_seed=None):
        self.rng_seed_int, self.rng = _seed_to_int(rng_seed, "RandField")
        if (range_max_x < range_min_x) or (range_max_y < range_min_y):
            print("the maximum range must be greater to equal to the minimum range")
        self.range_min_x, self.range_max_x = range_min_x, range_max_x
        self.range_min_y, self.range_max_y = range_min_y, range_max_y
        self.scale_min, self.scale_max, self.nugget_max = scale_min, scale_max, nugget_max
        if model_name not in ("Gaussian", "Exponential", "Matern"):
            raise Exception("please put in a valid model_name, including Gaussian, Exponential, and Matern")
        if model_name == "Matern" and smoothness is None:
            raise Exception("a smoothness value must be defined if model name is Matern")
        self.smoothness = smoothness
        self.model_name = model_name
        self.isotropic = isotropic
        self._draws = 0          # Philox iteration counter of stand-alone get_rfblock() calls
        self._ctx = None

    def _config_key(self):
        """Hashable digest of everything a device context derives from this object (model, ranges, block table, tapers)."""
        import hashlib
        h = hashlib.blake2b(digest_size=16)
        pairs = getattr(self, "pairs", None)
        if pairs is not None:
            h.update(np.ascontiguousarray(pairs, dtype=np.int64).tobytes())
        for m in getattr(self, "edge_masks", None) or []:
            a = np.ascontiguousarray(m, dtype=np.float64)
            h.update(repr(a.shape).encode())
            h.update(a.tobytes())
        return (self.model_name, self.smoothness, bool(self.isotropic), self.range_min_x, self.range_max_x, self.range_min_y,
                self.range_max_y, self.scale_min, self.scale_max, self.nugget_max, getattr(self, "resolution", None),
                h.hexdigest())

    def set_generation_method(self, spectral, n_modes=1000):
        """True: FFT spectral synthesis (A3).  False: the gstools randomization method of `get_random_field`
        (MCMC.py:625-687) with `n_modes` wave vectors (gstools SRF default mode_no=1000), generated on the GPU."""
        self.spectral = spectral
        self.n_modes = int(n_modes)
        if self._ctx is not None:
            self._ctx.set_generation_method(bool(spectral), self.n_modes)

    def set_block_sizes(self, min_block_x, max_block_x, min_block_y, max_block_y, steps=5):
        self.min_block_x, self.max_block_x = min_block_x, max_block_x
        self.min_block_y, self.max_block_y = min_block_y, max_block_y
        self.steps = steps
        self.pairs = self.get_block_sizes()
        self._ctx = None

    def get_block_sizes(self):
        """[2, steps^2] int array: row 0 widths, row 1 heights, forced even (MCMC.py:568-581)."""
        width = np.linspace(self.min_block_x, self.max_block_x, self.steps, dtype=int)
        height = np.linspace(self.min_block_y, self.max_block_y, self.steps, dtype=int)
        w, h = np.meshgrid(width, height)
        return np.array([(w // 2 * 2).flatten(), (h // 2 * 2).flatten()])

    def set_weight_param(self, logis_func_L, logis_func_x0, logis_func_k, logis_func_offset, max_dist, resolution):
        if not hasattr(self, "pairs"):
            raise Exception("It seems like the set_block_sizes has not been called yet before calling set_weight_param")
        self.logistic_param = [logis_func_L, logis_func_x0, logis_func_k, logis_func_offset]
        self.max_dist = max_dist
        self.resolution = resolution
        self.edge_masks = self.get_edge_masks()
        self._ctx = None

    def _logistic(self, dist):
        L, x0, k, offset = self.logistic_param
        scaled = np.where(dist > self.max_dist, 1, dist / self.max_dist)
        return L / (1 + np.exp(-k * (scaled - x0))) - offset

    def get_edge_masks(self):
        """Taper of each block size: logistic of the distance to the block rim (MCMC.py:583-623).  The rim is the
        full border, so the nearest-rim distance is the smallest of the four axis distances (what the reference's
        KDTree query returns, without the tree)."""
        if not hasattr(self, "pairs"):
            raise Exception("It seems like the set_block_sizes has not been called yet before calling get_edge_mask")
        masks = []
        for i in range(self.pairs.shape[1]):
            bw, bh = int(self.pairs[0, i]), int(self.pairs[1, i])
            px = np.arange(bw) * self.resolution
            py = np.arange(bh) * self.resolution
            dx = np.minimum(px - px[0], px[-1] - px)
            dy = np.minimum(py - py[0], py[-1] - py)
            masks.append(self._logistic(np.minimum(dy[:, None], dx[None, :])))
        return masks

    def get_crf_weight(self, xx, yy, cond_data_mask):
        """(weight, dist, dist_rescale, dist_logi): conditioning weight, 0 at data cells (MCMC.py:689-714).
        One-time setup, outside the hot path; the nearest-data distance is an exact scan on the GPU (Utilities.py,
        bit-identical to the reference's KD-tree query).  No CPU fallback: without a CUDA device this raises GmcError; a
        host that only prepares inputs can compute the distance map itself and call get_crf_weight_from_dist."""
        sel = np.asarray(cond_data_mask) == 1
        from .Utilities import min_dist_from_mask
        dist = min_dist_from_mask(np.asarray(xx), np.asarray(yy), sel)
        return self.get_crf_weight_from_dist(xx, yy, dist)

    def get_crf_weight_from_dist(self, xx, yy, dist):
        """Same, from a precomputed distance map (MCMC.py:716-740)."""
        scaled = np.where(dist > self.max_dist, 1, dist / self.max_dist)
        logi = self._logistic(dist)
        return logi - np.min(logi), dist, scaled, logi

    # ---- stand-alone field generation on the GPU --------------------------------------------------------------
    def _field_ctx(self):
        if self._ctx is None:
            hmax, wmax = int(self.pairs[1].max()), int(self.pairs[0].max())
            ctx = Context(max(hmax, 2), max(wmax, 2), 1)
            ctx.set_field_model(self.model_name, self.smoothness, self.isotropic, self.range_min_x, self.range_max_x,
                                self.range_min_y, self.range_max_y, self.scale_min, self.scale_max, self.nugget_max)
            ctx.set_blocks(self.pairs, self.edge_masks, self.resolution)
            self._ctx = ctx
        return self._ctx

    def _draw_field_params(self):
        """scale, nug, range1, range2, angle in the reference's draw order (MCMC.py:200-207 / 642-653)."""
        g = self.rng
        scale = g.uniform(self.scale_min, self.scale_max) / 3.0
        nug = g.uniform(0.0, self.nugget_max)
        angle = 0.0
        if not self.isotropic:
            rx = g.uniform(self.range_min_x, self.range_max_x)
            ry = g.uniform(self.range_min_y, self.range_max_y)
            if not getattr(self, "spectral", True):
                angle = g.uniform(0, 180)
        else:
            rx = ry = g.uniform(self.range_min_x, self.range_max_x)
        return scale, nug, rx, ry, angle

    def _generate(self, ctx, pick, bh, bw, apply_taper):
        import torch
        dev = ctx.device
        scale, nug, rx, ry, angle = self._draw_field_params()
        out = torch.empty((1, ctx.max_h * ctx.max_w), dtype=torch.float64, device=dev)

        def t(v, dt=torch.float64):
            return torch.tensor([v], dtype=dt, device=dev)
        seeds = keys_tensor([philox_key(self.rng_seed_int)], dev)
        if getattr(self, "spectral", True):
            ctx.field_spectral(t(pick, torch.int32), t(scale), t(nug), t(rx), t(ry), out, seeds=seeds,
                               iteration=self._draws, apply_taper=apply_taper)
        else:
            ctx.field_randmeth(t(pick, torch.int32), t(scale), t(nug), t(rx), t(ry), t(angle), out,
                               n_modes=getattr(self, "n_modes", 1000), seeds=seeds, iteration=self._draws,
                               apply_taper=apply_taper)
        self._draws += 1
        return out[0, :bh * bw].reshape(bh, bw).cpu().numpy()

    def get_random_field(self, X, Y, n=1):
        """One realisation [len(Y), len(X)] of the randomization-method generator (MCMC.py:625-687: gstools
        SRF(model).structured([X, Y]).T * scale), summed on the GPU (A5).  Like the reference, only the first of the
        `n` realisations is returned, and X, Y are taken as regular axes starting at 0."""
        X, Y = np.asarray(X, dtype=np.float64), np.asarray(Y, dtype=np.float64)
        nx, ny = len(X), len(Y)
        res = float(X[1] - X[0]) if nx > 1 else float(getattr(self, "resolution", 1.0))
        ctx = Context(max(ny, 2), max(nx, 2), 1)
        ctx.set_field_model(self.model_name, self.smoothness, self.isotropic, self.range_min_x, self.range_max_x,
                            self.range_min_y, self.range_max_y, self.scale_min, self.scale_max, self.nugget_max)
        ctx.set_blocks(np.array([[nx], [ny]]), [np.ones((ny, nx))], res)
        spectral, self.spectral = getattr(self, "spectral", True), False
        try:
            return self._generate(ctx, 0, ny, nx, apply_taper=False)
        finally:
            self.spectral = spectral
            ctx.close()

    def get_rfblock(self):
        """One tapered proposal field f = field * edge_mask (MCMC.py:742-778), synthesised by kernel K1 (spectral) or
        the randomization-method kernel (set_generation_method(False))."""
        ctx = self._field_ctx()
        pick = int(self.rng.integers(low=0, high=self.pairs.shape[1], size=1)[0])
        bw, bh = int(self.pairs[0, pick]), int(self.pairs[1, pick])
        return self._generate(ctx, pick, bh, bw, apply_taper=True)


def spectral_synthesis_field(RF, shape, res=1.0, *, draws=None):
    """2-D Gaussian random field by spectral synthesis on the GPU (reference MCMC.py:176-254).

    With `draws=None` the parameters come from RF.rng exactly like the reference (scale, nug, range[, range_y]) and the
    normals from the device Philox stream.  `draws=dict(scale, nug, range_x, range_y, z_re, z_im, z_nug)` injects
    every random input, which makes the result comparable to the reference to ~1e-13 (tests/test_gpu_field.py).
    """
    import torch
    ny, nx = shape
    ctx = Context(max(ny, 2), max(nx, 2), 1)
    ctx.set_field_model(RF.model_name, RF.smoothness, RF.isotropic, RF.range_min_x, RF.range_max_x, RF.range_min_y,
                        RF.range_max_y, RF.scale_min, RF.scale_max, RF.nugget_max)
    ctx.set_blocks(np.array([[nx], [ny]]), [np.ones((ny, nx))], res)
    dev = ctx.device
    if draws is None:
        g = RF.rng
        scale = g.uniform(RF.scale_min, RF.scale_max) / 3.0
        nug = g.uniform(0.0, RF.nugget_max)
        if not RF.isotropic:
            rx = g.uniform(RF.range_min_x, RF.range_max_x)
            ry = g.uniform(RF.range_min_y, RF.range_max_y)
        else:
            rx = ry = g.uniform(RF.range_min_x, RF.range_max_x)
        z = {}
    else:
        scale, nug, rx, ry = draws["scale"], draws["nug"], draws["range_x"], draws["range_y"]
        z = {k: torch.as_tensor(np.ascontiguousarray(draws[k], dtype=np.float64).reshape(1, ny * nx), device=dev)
             for k in ("z_re", "z_im", "z_nug")}

    def t(v, dt=torch.float64):
        return torch.tensor([v], dtype=dt, device=dev)
    out = torch.empty((1, ny * nx), dtype=torch.float64, device=dev)
    seeds = keys_tensor([philox_key(getattr(RF, "rng_seed_int", 0))], dev)
    it = getattr(RF, "_draws", 0)
    ctx.field_spectral(t(0, torch.int32), t(scale), t(nug), t(rx), t(ry), out, seeds=None if z else seeds, iteration=it,
                       apply_taper=False, **z)
    if not z:
        RF._draws = it + 1
    res_np = out.reshape(ny, nx).cpu().numpy()
    ctx.close()
    return res_np


# ------------------------------------------------------------------------------------------------------------------
# chain base class                                                                            reference MCMC.py:780-1081
# ------------------------------------------------------------------------------------------------------------------
class chain:
    def __init_func__(self):
        return

    def __init__(self, xx, yy, initial_bed, surf, velx, vely, dhdt, smb, cond_bed, data_mask, grounded_ice_mask,
                 resolution):
        self.xx, self.yy = xx, yy
        self.initial_bed = initial_bed
        self.surf, self.velx, self.vely, self.dhdt, self.smb = surf, velx, vely, dhdt, smb
        self.cond_bed = cond_bed
        self.data_mask = data_mask
        self.grounded_ice_mask = grounded_ice_mask
        self.resolution = resolution
        self.loss_function_list = []
        self.sample_loc = None
        shp = initial_bed.shape
        if any(a.shape != shp for a in (surf, velx, vely, dhdt, smb, cond_bed, data_mask)):
            raise Exception("the shape of bed, surf, velx, vely, dhdt, smb, radar_bed, data_mask need to be same")
        self._ctx = None
        self._ctx_key = None
        self._philox_iter = 1        # iteration 0 is the initial state (MCMC.py:1193-1199)
        self.__init_func__()

    def set_update_region(self, update_in_region, region_mask=[]):
        self.update_in_region = update_in_region
        if update_in_region is False or update_in_region == False:  # noqa: E712  (reference accepts 0/False)
            print("the update blocks is set to be randomly generated for any locations inside the entire map")
            self.region_mask = np.full(self.xx.shape, 1)
        else:
            if np.shape(region_mask) != self.xx.shape:
                raise ValueError("the region_mask input is invalid. It has to be a 2D numpy array with the shape of the map")
            print("the update blocks is set to be randomly generated for any locations inside the given region")
            self.region_mask = region_mask
        self._ctx = None

    def set_loss_type(self, sigma_mc=-1, massConvInRegion=True):
        if massConvInRegion:
            self.mc_region_mask = self.region_mask
        else:
            self.mc_region_mask = np.full(self.xx.shape, 1)
        self.sigma_mc = sigma_mc
        self._ctx = None

    def set_random_generator(self, rng_seed=None):
        self.rng_seed_int, self.rng = _seed_to_int(rng_seed, "chain")
        if isinstance(rng_seed, np.random.Generator):
            chain.seed = rng_seed        # the reference stores it on the class (MCMC.py:1062)

    def set_sample_points_locations(self, loc):
        self.sample_loc = loc

    # ---- device context ------------------------------------------------------------------------------------------
    @staticmethod
    def _binary(mask, name):
        m = np.asarray(mask)
        if not np.isin(m, (0, 1)).all():
            raise ValueError(f"{name} must contain only 0/1 (the reference mixes `== 1` and truthiness tests, which agree "
                             "only for binary masks)")
        return m.astype(np.uint8)

    def _static_args(self):
        gate = self.region_mask if self.update_in_region else self.grounded_ice_mask
        gate = self._binary(gate, "region_mask/grounded_ice_mask")
        mc = self._binary(self.mc_region_mask, "mc_region_mask")
        centre = np.flatnonzero(self._binary(self.region_mask, "region_mask").ravel() == 1).astype(np.int32) \
            if self.update_in_region else None
        if centre is not None and centre.size == 0:
            raise ValueError("region_mask selects no cell: the reference would loop forever drawing a block centre")
        weight = self.crf_data_weight if getattr(self, "block_type", "RF") == "CRF_weight" else None
        return dict(surf=self.surf, velx=self.velx, vely=self.vely, dhdt=self.dhdt, smb=self.smb, gate_mask=gate, mc_mask=mc,
                    centre_cells=centre, crf_weight=weight, resolution=self.resolution, sigma_mc=self.sigma_mc)

    def _context(self, max_chains=1, RF=None, device=None):
        # the cached context holds the field model, block table and tapers: key it on the RandField's CONTENTS (a setter
        # call on RF, or a re-created RF that happens to reuse the id, must not leave the old configuration in place)
        key = (max_chains, None if RF is None else RF._config_key(), str(device))
        if self._ctx is None or self._ctx_key != key:
            H, W = self.xx.shape
            ctx = Context(H, W, max_chains, device)
            ctx.set_static(**self._static_args())
            if RF is not None:
                if int(np.max(RF.pairs[1])) > H or int(np.max(RF.pairs[0])) > W:
                    # a block cut by BOTH edges of the grid: the reference's slices f[mxmin:mxmax] and bed[bxmin:bxmax]
                    # (MCMC.py:1267-1276) then differ in length and numpy raises on the first such proposal
                    raise ValueError(f"operands could not be broadcast together: a {int(np.max(RF.pairs[1]))}x"
                                     f"{int(np.max(RF.pairs[0]))} proposal block does not fit the {H}x{W} grid")
                ctx.set_field_model(RF.model_name, RF.smoothness, RF.isotropic, RF.range_min_x, RF.range_max_x,
                                    RF.range_min_y, RF.range_max_y, RF.scale_min, RF.scale_max, RF.nugget_max)
                ctx.set_blocks(RF.pairs, RF.edge_masks, RF.resolution)
            self._ctx, self._ctx_key = ctx, key
        if RF is not None:
            self._ctx.set_generation_method(bool(getattr(RF, "spectral", True)), getattr(RF, "n_modes", 1000))
        return self._ctx

    def loss(self, massConvResidual, dataDiff):
        """(total, mc, data) = nansum(res[mc_region_mask==1]^2)/(2 sigma_mc^2), data loss 0 (MCMC.py:1021-1044); K3."""
        import torch
        ctx = self._context(1, getattr(self, "_last_rf", None))
        res = torch.as_tensor(np.ascontiguousarray(massConvResidual, dtype=np.float64), device=ctx.device)[None]
        out = torch.empty(1, dtype=torch.float64, device=ctx.device)
        ctx.loss(res.contiguous(), out)
        loss_mc = float(out.item())
        return loss_mc + 0, loss_mc, 0


# ------------------------------------------------------------------------------------------------------------------
# chain_crf                                                                                  reference MCMC.py:1083-1443
# ------------------------------------------------------------------------------------------------------------------
class chain_crf(chain):
    def __init_func__(self):
        print("before running the chain, please set where the block update will be using the object's function "
              "set_update_region(update_in_region, region_mask)")
        print("then please set up the loss function using either set_loss_type or set_loss_func")
        print("an RandField object also need to be created correctly and passed in set_crf_data_weight(RF) and in run(n_iter, RF)")

    def set_update_type(self, block_type):
        if block_type == "CRF_rbf":
            print("The update block is set to conditional random field generated by rbf method (not implemented yet)")
        elif block_type == "CRF_weight":
            print("The update block is set to conditional random field generated by calculating weights with logistic function")
        elif block_type == "RF":
            print("The update block is set to Random Field")
        else:
            raise ValueError("The block_type argument should be one of the following: CRF_weight, CRF_rbf, RF")
        self.block_type = block_type
        self._ctx = None

    def set_crf_data_weight(self, RF):
        self.crf_data_weight = RF.get_crf_weight(self.xx, self.yy, self.data_mask)[0]
        self._ctx = None

    # ---- the hot loop ------------------------------------------------------------------------------------------
    def run(self, n_iter, RF, only_save_last_bed=False, info_per_iter=1000, plot=True, progress_bar=True, *,
            replay=None, resync_every=4096):
        """Run n_iter-1 Metropolis proposals on the GPU; same return tuple as the reference (MCMC.py:1137, 1434-1443).

        replay: optional sequence of dict(f, idx_x, idx_y, u) — the recorded proposals of a reference run; the chain
        then replays them through gmc_step_injected and reproduces the reference trajectory bit-for-bit.
        plot / progress_bar are accepted for signature compatibility (UI is out of scope): a one-line summary is printed
        every `info_per_iter` iterations when progress_bar is False, like the reference's status line.
        """
        if not isinstance(RF, RandField):
            raise TypeError('The arugment "RF" has to be an object of the class RandField')
        if not hasattr(self, "rng_seed_int"):
            self.set_random_generator(None)
        batch = ChainBatch(self, RF, np.ascontiguousarray(self.initial_bed, dtype=np.float64)[None],
                           [philox_key(self.rng_seed_int, RF.rng_seed_int)], iter0=self._philox_iter,
                           track_resampled=True)
        self._last_rf = RF
        H, W = self.xx.shape
        loss_cache = np.zeros(n_iter)
        step_cache = np.zeros(n_iter)
        blocks_cache = np.full((n_iter, 4), np.nan)
        bed_cache = None if only_save_last_bed else np.zeros((n_iter, H, W))
        sample_values = sample_ij = None
        if self.sample_loc is not None:
            sample_values = np.zeros((self.sample_loc.shape[0], n_iter))
            sample_ij = np.zeros(self.sample_loc.shape, dtype=np.int64)
            for k in range(self.sample_loc.shape[0]):
                si, sj = np.where((self.xx == self.sample_loc[k, 0]) & (self.yy == self.sample_loc[k, 1]))
                sample_ij[k, :] = [int(si[0]), int(sj[0])]
            sample_values[:, 0] = np.asarray(self.initial_bed)[sample_ij[:, 0], sample_ij[:, 1]]
        loss_cache[0] = batch.loss()[0]
        if bed_cache is not None:
            bed_cache[0] = self.initial_bed

        per_step = (bed_cache is not None) or (sample_values is not None) or (replay is not None)
        t0 = time.time()
        done = 1
        while done < n_iter:
            if replay is not None:
                p = replay[done - 1]
                acc, loss_now = batch.step_injected([p["f"]], [(p["idx_x"], p["idx_y"])], [p["u"]])
                loss_cache[done], step_cache[done] = loss_now[0], acc[0]
                blocks_cache[done] = [p["idx_x"], p["idx_y"], p["f"].shape[0], p["f"].shape[1]]
                n = 1
            else:
                n = 1 if per_step else min(n_iter - done, max(int(info_per_iter), 1) if not progress_bar else n_iter)
                lc, st, bl = batch.advance(n, resync_every=resync_every)
                loss_cache[done:done + n], step_cache[done:done + n], blocks_cache[done:done + n] = lc[0], st[0], bl[0]
            done += n
            if bed_cache is not None or sample_values is not None:
                bed_now = batch.beds()[0]
                if bed_cache is not None:
                    bed_cache[done - 1] = bed_now
                if sample_values is not None:
                    sample_values[:, done - 1] = bed_now[sample_ij[:, 0], sample_ij[:, 1]]
            if not progress_bar and ((done - 1) % max(int(info_per_iter), 1) == 0 or done == n_iter):
                el = max(time.time() - t0, 1e-9)
                print(f"Chain {getattr(self, 'chain_id', 'Unknown')} ({str(getattr(self, 'seed', 'Unknown'))[:6]}): "
                      f"{100.0 * (done - 1) / max(n_iter - 1, 1):3.0f}% | it/s: {(done - 1) / el:8.2f} | n: {n_iter} | "
                      f"loss: {loss_cache[done - 1]:.3e} | acc: {np.sum(step_cache) / done:.4f}")
                sys.stdout.flush()
        self._philox_iter = batch.iteration
        bed_c = batch.beds()[0]
        resampled = batch.resampled_times()[0]
        loss_data_cache = np.zeros(n_iter)
        loss_mc_cache = loss_cache.copy()
        first = bed_c if only_save_last_bed else bed_cache
        out = (first, loss_mc_cache, loss_data_cache, loss_cache, step_cache, resampled, blocks_cache)
        batch.close()
        return out + ((sample_values,) if sample_values is not None else ())

    def run_many(self, n_iter, RF, initial_beds, rng_seeds, device=None, resync_every=4096, track_resampled=True,
                 as_arrays=False, batch=None, out=None, pipeline_groups=None, wait=True):
        """Batched form: C independent chains (one per initial bed / seed) stepped concurrently on one GPU.

        Returns a list of the reference's 7-tuples (only_save_last_bed=True form), one per chain — what
        largeScaleChain_mp collects from its worker processes (largeScaleChain_multiprocessing.py:78-79) — or, with
        as_arrays=True, a dict of stacked arrays (bed[C,H,W], loss[C,n], steps[C,n], blocks[C,n,4], resampled[C,H,W]).
        initial_beds: numpy [C,H,W] or a (pinned) CPU / CUDA torch tensor.  `batch` reuses the device buffers of a
        previous call (and stays open: only a batch created here is closed here); `out` = dict of pinned CPU tensors
        (bed, loss, steps, blocks and optionally resampled: int32 coverage counts) to receive the results.
        wait=False (pinned beds + batch + out, as_arrays=True): the uploads, kernels and downloads are only queued and a
        PendingRun is returned; its wait() blocks until the results are in `out`.  Two batches used alternately keep two
        steps in flight, so the transfers of one overlap the compute of the other.
        """
        if not isinstance(RF, RandField):
            raise TypeError('The arugment "RF" has to be an object of the class RandField')
        keys = [philox_key(s, s) for s in rng_seeds]
        import torch
        pipelined = (batch is not None and out is not None and isinstance(initial_beds, torch.Tensor) and initial_beds.is_pinned()
                     and (pipeline_groups is None or pipeline_groups > 1))
        if not wait and not (pipelined and as_arrays):
            raise ValueError("wait=False needs pinned initial beds, a reusable batch, pinned `out` tensors and as_arrays=True")
        own_batch = batch is None
        if pipelined:
            # pinned host buffers on both sides: overlap the transfers with compute
            if pipeline_groups is None:
                # finer ranges shorten the exposed head (first upload) and tail (last download) of the pipeline:
                # 4 ranges 50.5 ms, 8: 49.6 ms, 16: 49.1 ms per 256 x 1000-iteration step (profiles/README.md)
                pipeline_groups = int(os.environ.get("GMC_PIPELINE_GROUPS", "16"))
            pending = batch.run_pipelined(initial_beds, keys, n_iter - 1, out, groups=pipeline_groups, resync_every=resync_every,
                                          wait=False)
            if not wait:
                return pending
            res = pending.wait()
            res["batch"] = batch
        else:
            if batch is None:
                batch = ChainBatch(self, RF, initial_beds, keys, iter0=1, device=device, track_resampled=track_resampled)
            else:
                batch.reset(initial_beds, keys, iter0=1)
            res = batch.advance_into(n_iter - 1, resync_every=resync_every, out=out)
            res["batch"] = batch
        if as_arrays:
            return res
        C = batch.C
        final, lc, st, bl = res["bed"], res["loss"], res["steps"], res["blocks"]
        res_t = batch.resampled_times() if batch.resampled is not None else [np.zeros((batch.H, batch.W))] * C
        outl = []
        for c in range(C):
            loss = np.array(lc[c], dtype=np.float64)
            blocks = np.array(bl[c], dtype=np.float64)
            blocks[0, :] = np.nan                     # MCMC.py:1169: row 0 of blocks_cache stays NaN
            outl.append((np.array(final[c]), loss.copy(), np.zeros(n_iter), loss, np.array(st[c], dtype=np.float64), res_t[c],
                         blocks))
        if own_batch:
            batch.close()
        return outl


# ------------------------------------------------------------------------------------------------------------------
# chain_sgs                                                                                  reference MCMC.py:1445-1911
# ------------------------------------------------------------------------------------------------------------------
class chain_sgs(chain):
    """Small-scale chain: block re-simulation by Sequential Gaussian Simulation (reference MCMC.py:1445).  Same setters
    and `run` signature/return tuple; every per-node operation (octant search, kriging solve, normal-score transform,
    residual, loss, Metropolis) runs in kernel K6 (csrc/sgs.cu)."""

    def __init_func__(self):
        print("before running the chain, please set where the block update will be using the object's function "
              "set_update_in_region(region_mask) and set_update_region(update_in_region)")
        print("please also set up the sgs parameters using set_sgs_param(self, block_size, sgs_param)")
        print("then please set up the loss function using either set_loss_type or set_loss_func")

    def set_normal_transformation(self, nst_trans, do_transform=True):
        self.do_transform = do_transform
        self.nst_trans = nst_trans if do_transform else None
        self._ctx = None

    def set_trend(self, trend=None, detrend_map=True):
        if detrend_map == True:  # noqa: E712
            if len(trend) != len(self.xx) or trend.shape != self.xx.shape:
                raise ValueError("if detrend_map is set to True, then the trend of the topography, which is a 2D numpy array, must be provided")
            self.trend = trend
        else:
            self.trend = None
        self.detrend_map = detrend_map
        self._ctx = None

    def set_variogram(self, vario_type, vario_range, vario_sill, vario_nugget, isotropic=True, vario_smoothness=None,
                      vario_azimuth=None):
        if vario_type in ("Gaussian", "Exponential", "Spherical"):
            print("the variogram is set to type", vario_type)
        elif vario_type == "Matern":
            if (vario_smoothness is None) or (vario_smoothness <= 0):
                raise ValueError("vario_smoothness argument should be a positive float when the vario_type is Matern")
            print("the variogram is set to type", vario_type)
        else:
            raise ValueError("vario_type argument should be one of the following: Gaussian, Exponential, Spherical, or Matern")
        self.vario_type = vario_type
        if isotropic:
            self.vario_param = [0, vario_nugget, vario_range, vario_range, vario_sill, vario_type, vario_smoothness]
        else:
            if len(vario_range) == 2:
                print("set to anistropic variogram with major range and minor range to be ", vario_range)
                self.vario_param = [vario_azimuth, vario_nugget, vario_range[0], vario_range[1], vario_sill, vario_type,
                                    vario_smoothness]
            else:
                raise ValueError("vario_range need to be a list with two floats to specifying for major range and minor range "
                                 "of the variogram when isotropic is set to False")
        self._ctx = None

    def set_sgs_param(self, sgs_num_nearest_neighbors, sgs_searching_radius, sgs_rand_dropout_on=False, dropout_rate=0):
        if sgs_rand_dropout_on == False:  # noqa: E712
            dropout_rate = 0
            print("because the sgs_rand_dropout_on is set to False, the dropout_rate is automatically set to 0")
        self.sgs_param = [sgs_num_nearest_neighbors, sgs_searching_radius, sgs_rand_dropout_on, dropout_rate]
        self._ctx = None

    def set_block_sizes(self, block_min_x, block_max_x, block_min_y, block_max_y):
        self.block_min_x, self.block_max_x = block_min_x, block_max_x
        self.block_min_y, self.block_max_y = block_min_y, block_max_y
        self._ctx = None

    # ---- device context ------------------------------------------------------------------------------------------
    def _vario_dict(self):
        """The dict the reference hands to sgs() (MCMC.py:1682-1702)."""
        p = self.vario_param
        v = dict(azimuth=p[0], nugget=p[1], major_range=p[2], minor_range=p[3], sill=p[4], vtype=p[5])
        if p[5] == "Matern":
            v["s"] = p[6]
        return v

    def _sgs_context(self, max_chains, device=None, widen=False):
        from . import sgs_tables as T
        key = ("sgs", max_chains, str(device), bool(widen))
        if self._ctx is not None and self._ctx_key == key:
            return self._ctx
        H, W = self.xx.shape
        region = self._binary(self.region_mask, "region_mask")
        centre = np.flatnonzero(region.ravel() == 1).astype(np.int32)        # the reference always samples inside region_mask
        if centre.size == 0:
            raise ValueError("region_mask selects no cell: the reference would loop forever drawing a block centre")
        ctx = Context(H, W, max_chains, device)
        ones = np.ones((H, W), dtype=np.uint8)
        ctx.set_static(self.surf, self.velx, self.vely, self.dhdt, self.smb, ones, self._binary(self.mc_region_mask, "mc_region_mask"),
                       centre, None, self.resolution, self.sigma_mc)
        dx, dy = T.grid_steps(np.asarray(self.xx), np.asarray(self.yy))
        vario = self._vario_dict()
        # The reference widens the search radius by 100 km for a node that finds no conditioned cell (MCMC.py:149-155).
        # Every cell outside the block is conditioned unless the bed holds NaN, so with a radius longer than the largest
        # block edge the first level always finds data and the (much larger) tables of the wider radii are not built.
        radius = float(self.sgs_param[1])
        bmax = max(self.block_max_x, self.block_max_y)
        self._sgs_levels = bool(widen) or radius / min(abs(dx), abs(dy)) <= bmax + 1
        if self._sgs_levels:
            off, cnt, hw, _radii = T.search_levels(dx, dy, H, W, radius)
        else:
            off, cnt, hw = T.octant_stencil(dx, dy, radius)
        lut = T.covariance_lut(dx, dy, hw, vario)
        trend = np.asarray(self.trend, dtype=np.float64) if self.detrend_map else None
        cond_c = np.asarray(self.cond_bed, dtype=np.float64) - (trend if trend is not None else 0.0)
        if self.do_transform:
            zcond = self.nst_trans.transform(cond_c.reshape(-1, 1)).reshape(H, W)     # MCMC.py:1653-1661
            quant, refs = self.nst_trans.quantiles_[:, 0], self.nst_trans.references_
            if getattr(self.nst_trans, "output_distribution", "normal") != "normal":
                raise NotImplementedError("only QuantileTransformer(output_distribution='normal') is supported")
        else:
            zcond, quant, refs = cond_c, None, None
        ctx.sgs_setup(trend, zcond, self.grounded_ice_mask, quant, refs, off, cnt, hw, self.sgs_param[0], lut, vario["sill"],
                      (self.block_min_x, self.block_max_x, self.block_min_y, self.block_max_y))
        self._ctx, self._ctx_key = ctx, key
        return ctx

    def run(self, n_iter, only_save_last_bed=False, info_per_iter=100, plot=True, progress_bar=True, *, replay=None):
        """n_iter block re-simulation proposals on the GPU; the reference's return tuple (MCMC.py:1599, 1897-1911).

        replay: optional sequence of dict(idx_x, idx_y, bsx, bsy, path[n,2], z[n], u) — the recorded random inputs of a
        reference run, replayed through gmc_sgs_step_injected."""
        if not hasattr(self, "rng_seed_int"):
            self.set_random_generator(None)
        H, W = self.xx.shape
        batch = SgsBatch(self, np.ascontiguousarray(self.initial_bed, dtype=np.float64)[None], [philox_key(self.rng_seed_int)],
                         iter0=getattr(self, "_philox_iter_sgs", 0))
        loss_cache, step_cache = np.zeros(n_iter), np.zeros(n_iter)
        blocks_cache = np.full((n_iter, 4), np.nan)
        bed_cache = None if only_save_last_bed else np.zeros((n_iter, H, W))
        sample_values = sample_ij = None
        if self.sample_loc is not None:
            sample_values = np.zeros((self.sample_loc.shape[0], n_iter))
            sample_ij = np.zeros(self.sample_loc.shape, dtype=np.int64)
            for k in range(self.sample_loc.shape[0]):
                si, sj = np.where((self.xx == self.sample_loc[k, 0]) & (self.yy == self.sample_loc[k, 1]))
                sample_ij[k, :] = [int(si[0]), int(sj[0])]
        per_step = (bed_cache is not None) or (sample_values is not None) or (replay is not None)
        done = 0
        while done < n_iter:
            if replay is not None:
                p = replay[done]
                acc, loss_now = batch.step_injected([p])
                loss_cache[done], step_cache[done] = loss_now[0], acc[0]
                blocks_cache[done] = [p["idx_x"], p["idx_y"], p["bsx"], p["bsy"]]
                n = 1
            else:
                n = 1 if per_step else n_iter - done
                lc, st, bl = batch.advance(n)
                loss_cache[done:done + n], step_cache[done:done + n], blocks_cache[done:done + n] = lc[0], st[0], bl[0]
            done += n
            if per_step and (bed_cache is not None or sample_values is not None):
                bed_now = batch.beds(with_trend=True)[0]
                if bed_cache is not None:
                    bed_cache[done - 1] = bed_now
                if sample_values is not None:
                    sample_values[:, done - 1] = batch.beds(with_trend=False)[0][sample_ij[:, 0], sample_ij[:, 1]]
        self._philox_iter_sgs = batch.iteration
        last_bed = batch.beds(with_trend=True)[0]
        resampled = batch.resampled_times()[0]
        first = last_bed if only_save_last_bed else bed_cache
        out = (first, loss_cache.copy(), np.zeros(n_iter), loss_cache, step_cache, resampled, blocks_cache)
        batch.close()
        return out + ((sample_values,) if sample_values is not None else ())


class SgsBatch:
    """Device-resident state of C small-scale chains: detrended bed, its normal score, residual, nansum, guard count."""

    def __init__(self, chain_obj, initial_beds, keys, iter0=0, device=None):
        import torch
        self.torch = torch
        beds = np.ascontiguousarray(initial_beds, dtype=np.float64)
        if beds.ndim != 3 or beds.shape[1:] != chain_obj.xx.shape:
            raise GmcShapeError(f"initial beds have shape {beds.shape}, expected [C,{chain_obj.xx.shape[0]},{chain_obj.xx.shape[1]}]")
        self.C, self.H, self.W = beds.shape
        self.chain = chain_obj
        # Precondition of the resident-normal-score formulation: the reference maps the WHOLE grid through
        # inverse_transform(transform(.)) on every proposal (MCMC.py:1767,1777), which clamps every detrended cell outside
        # the transformer's fitted range [quantiles_[0], quantiles_[-1]] from its first accepted step on; this kernel
        # round-trips only the accepted block.  The two agree (to the transform's own <= 1e-12 round-trip drift) exactly
        # when no detrended cell lies outside that range - true for a transformer fitted on initial_bed - trend as in the
        # reference's scripts, not for one fitted on the conditioning data alone.
        check_range = getattr(chain_obj, "do_transform", False) and getattr(chain_obj, "nst_trans", None) is not None
        from ._lib import require_cuda
        dev = require_cuda(device)
        full = torch.as_tensor(beds).to(dev)
        # NaN cells in the beds are unconditioned cells outside the block: a node may then find nothing within the radius
        # (looked for on the device: a host pass over C x H x W doubles costs more than the upload)
        self.ctx = chain_obj._sgs_context(self.C, device, widen=bool(torch.isnan(full).any().item()))
        dev = self.dev = self.ctx.device
        self.bedc = torch.empty_like(full)
        self.z = torch.empty_like(full)
        self.mcres = torch.empty_like(full)
        self.ssq = torch.empty(self.C, dtype=torch.float64, device=dev)
        self.nviol = torch.zeros(self.C, dtype=torch.int32, device=dev)
        self.resampled = torch.zeros(beds.shape, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seeds = keys_tensor(keys, dev)
        self.iteration = int(iter0)
        self.trend = None if not chain_obj.detrend_map else torch.as_tensor(np.ascontiguousarray(chain_obj.trend, dtype=np.float64)).to(dev)
        if check_range:                                            # counted on the device: the beds are already there
            q = np.asarray(chain_obj.nst_trans.quantiles_[:, 0], dtype=np.float64)
            base = full - self.trend if self.trend is not None else full
            n_out = int(((base < float(q[0])) | (base > float(q[-1]))).sum().item())
            del base
            if n_out:
                import warnings
                warnings.warn(f"chain_sgs: {n_out} detrended bed cells lie outside the normal-score transformer's fitted range "
                              f"[{q[0]:.6g}, {q[-1]:.6g}]; the reference clamps such cells to that range at its first accepted "
                              "step (MCMC.py:1767-1806) while this kernel leaves cells outside the proposed blocks untouched, "
                              "so the trajectories differ there.  Fit the transformer on initial_bed - trend (as the reference's "
                              "drivers do) or clip the initial beds to the fitted range.", RuntimeWarning, stacklevel=3)
        scratch = torch.empty_like(full)
        self.ctx.sgs_init(full, self.bedc, self.z, self.mcres, self.ssq, self.nviol, scratch)

    def close(self):
        self.bedc = self.z = self.mcres = self.resampled = None

    def _check_err(self):
        if int(self.err.item()) != 0:
            raise GmcError("SGS: a node found no conditioned neighbour inside the search radius nor inside the widened radii "
                           "(radius + 100 km steps, MCMC.py:149-155, up to the grid diagonal / sgs_tables.MAX_LEVELS)")

    def loss(self):
        den = self.torch.full_like(self.ssq, 2 * self.chain.sigma_mc ** 2)
        return (self.ssq / den).cpu().numpy()

    def beds(self, with_trend=True, out=None):
        """Final beds on the host; `out` (a pinned [C,H,W] float64 tensor) receives them without a pageable staging copy."""
        b = self.bedc + self.trend if (with_trend and self.trend is not None) else self.bedc
        if out is not None:
            out.copy_(b)
            return out.numpy()
        return b.cpu().numpy()

    def resampled_times(self):
        return self.resampled.cpu().numpy().astype(np.float64)

    def advance(self, n_steps):
        torch = self.torch
        lc = torch.empty((self.C, n_steps), dtype=torch.float64, device=self.dev)
        st = torch.empty((self.C, n_steps), dtype=torch.uint8, device=self.dev)
        bl = torch.empty((self.C, n_steps, 4), dtype=torch.int32, device=self.dev)
        self.ctx.sgs_run(self.bedc, self.z, self.mcres, self.ssq, self.nviol, self.seeds, self.iteration, n_steps, lc, st, bl, 0,
                         self.resampled, self.err)
        self.iteration += n_steps
        self._check_err()
        return lc.cpu().numpy(), st.cpu().numpy(), bl.cpu().numpy()

    def step_injected(self, tapes):
        """One step per chain from recorded inputs (dict(idx_x, idx_y, bsx, bsy, path[n,2] grid indices, z[n], u))."""
        torch = self.torch
        H, W = self.H, self.W
        nmax = max(len(t["path"]) for t in tapes)
        path = np.zeros((self.C, nmax), dtype=np.int32)
        zn = np.zeros((self.C, nmax))
        centre = np.zeros((self.C, 2), dtype=np.int32)
        bs = np.zeros((self.C, 2), dtype=np.int32)
        us = np.zeros(self.C)
        for c, t in enumerate(tapes):
            x0, x1 = max(0, int(t["idx_x"] - t["bsx"] / 2)), min(H, int(t["idx_x"] + t["bsx"] / 2))
            y0, y1 = max(0, int(t["idx_y"] - t["bsy"] / 2)), min(W, int(t["idx_y"] + t["bsy"] / 2))
            p = np.asarray(t["path"])
            path[c, :len(p)] = (p[:, 0] - x0) * (y1 - y0) + (p[:, 1] - y0)
            zn[c, :len(p)] = np.nan_to_num(np.asarray(t["z"], dtype=np.float64))
            centre[c] = (t["idx_x"], t["idx_y"])
            bs[c] = (t["bsx"], t["bsy"])
            us[c] = t["u"]
        dev = self.dev
        acc = torch.empty(self.C, dtype=torch.uint8, device=dev)
        loss = torch.empty(self.C, dtype=torch.float64, device=dev)
        cu = lambda a: torch.as_tensor(a).to(dev)                                         # noqa: E731
        self.ctx.sgs_step_injected(self.bedc, self.z, self.mcres, self.ssq, self.nviol, cu(centre), cu(bs), cu(path), cu(zn), cu(us),
                                   acc, loss, None, self.resampled, self.err)
        self.iteration += 1
        self._check_err()
        return acc.cpu().numpy().astype(bool), loss.cpu().numpy()


class PendingRun:
    """Handle of a queued ChainBatch.run_pipelined: wait() blocks until every range's results are in the pinned `out`
    tensors, checks the device-error flag and returns `out`.  The batch must not be reused before wait() returns."""

    def __init__(self, batch, out):
        self.batch, self.out = batch, out
        torch = batch.torch
        self._events = []
        for strm in batch._pre_streams:                          # the downloads of a range are queued on its high-priority stream
            ev = torch.cuda.Event()
            ev.record(strm)
            self._events.append(ev)
        self._done = False

    def wait(self):
        if not self._done:
            for ev in self._events:
                ev.synchronize()
            main = self.batch.torch.cuda.current_stream()
            for strm in self.batch._pre_streams:
                main.wait_stream(strm)
            self.batch.ctx.check_flag()
            self._done = True
        return dict(self.out)


class ChainBatch:
    """Device-resident state of C chains sharing one chain_crf configuration: bed[C,H,W], mcres[C,H,W], ssq[C]."""

    def __init__(self, chain_obj, RF, initial_beds, keys, iter0=1, device=None, track_resampled=False):
        import torch
        self.torch = torch
        shape = tuple(initial_beds.shape)
        if len(shape) != 3 or shape[1:] != chain_obj.xx.shape:
            raise GmcShapeError(f"initial beds have shape {shape}, expected [C,{chain_obj.xx.shape[0]},{chain_obj.xx.shape[1]}]")
        self.C, self.H, self.W = shape
        self.chain = chain_obj
        self.ctx = chain_obj._context(self.C, RF, device)
        dev = self.ctx.device
        self.dev = dev
        self.gate = np.asarray(chain_obj.region_mask if chain_obj.update_in_region else chain_obj.grounded_ice_mask)
        self.bed = torch.empty(shape, dtype=torch.float64, device=dev)
        self.mcres = torch.empty_like(self.bed)
        self.ssq = torch.empty(self.C, dtype=torch.float64, device=dev)
        self._loss = torch.empty(self.C, dtype=torch.float64, device=dev)
        self.resampled = torch.zeros(shape, dtype=torch.int32, device=dev) if track_resampled else None
        self._cache = None
        self.reset(initial_beds, keys, iter0)

    def reset(self, initial_beds, keys, iter0=1):
        """(Re)load C initial beds (numpy, pinned CPU tensor or CUDA tensor) and recompute residual + loss (K2/K3)."""
        torch = self.torch
        if len(keys) != self.C:
            raise ValueError("one seed per chain is required")
        src = initial_beds if isinstance(initial_beds, torch.Tensor) else \
            torch.as_tensor(np.ascontiguousarray(initial_beds, dtype=np.float64))
        if tuple(src.shape) != (self.C, self.H, self.W) or src.dtype != torch.float64:
            raise GmcShapeError(f"initial beds must be float64 [{self.C},{self.H},{self.W}]")
        self.bed.copy_(src, non_blocking=True)
        self.seeds = keys_tensor(keys, self.dev)
        if self.resampled is not None:
            self.resampled.zero_()
        self.iteration = int(iter0)
        self.ctx.residual_loss(self.bed, self.mcres, self._loss, self.ssq)

    def close(self):
        self.bed = self.mcres = self.resampled = self._cache = self._counts16 = None

    def _loss_now(self):
        # tensor / tensor is an IEEE division on the device (tensor / python scalar multiplies by a reciprocal)
        den = self.torch.full_like(self.ssq, 2 * self.chain.sigma_mc ** 2)
        return self.ssq / den

    def loss(self):
        return self._loss_now().cpu().numpy()

    def beds(self):
        return self.bed.cpu().numpy()

    def residuals(self):
        return self.mcres.cpu().numpy()

    def resampled_times(self):
        """float64 [C,H,W]: accepted-block coverage counts times the mask value (MCMC.py:1349-1352)."""
        if self.resampled is None:
            raise GmcError("ChainBatch was created with track_resampled=False")
        return self.resampled.cpu().numpy().astype(np.float64) * self.gate[None].astype(np.float64)

    def _copy_counts(self, dst, a, b, n_steps):
        """Coverage counts of chains [a, b) -> the pinned tensor `dst` (queued on the current stream).  `dst` may be int32
        (the device dtype) or int16: a count cannot exceed the number of proposals, so for runs of up to 32767 proposals the
        narrower type halves the largest download after the beds (the host link is what bounds the end-to-end rate on 8 GPUs)."""
        torch = self.torch
        if dst.dtype == self.resampled.dtype:
            dst[a:b].copy_(self.resampled[a:b], non_blocking=True)
            return
        if dst.dtype != torch.int16:
            raise ValueError("out['resampled'] must be an int32 or int16 tensor")
        if n_steps > 32767:
            raise ValueError("int16 coverage counts need at most 32767 proposals per run; pass an int32 tensor")
        if getattr(self, "_counts16", None) is None:
            self._counts16 = torch.empty((self.C, self.H, self.W), dtype=torch.int16, device=self.dev)
        self._counts16[a:b].copy_(self.resampled[a:b])                      # narrowing on the device
        dst[a:b].copy_(self._counts16[a:b], non_blocking=True)

    def _device_caches(self, n):
        torch = self.torch
        if self._cache is None or self._cache[0].shape[1] != n:
            self._cache = (torch.empty((self.C, n), dtype=torch.float64, device=self.dev),
                           torch.empty((self.C, n), dtype=torch.uint8, device=self.dev),
                           torch.empty((self.C, n, 4), dtype=torch.int32, device=self.dev))
        return self._cache

    def advance(self, n_steps, resync_every=4096, want_caches=True):
        """n_steps fused free-running iterations (kernel K1+K4).  Returns (loss[C,n], accepted[C,n], blocks[C,n,4])."""
        self.ctx.set_step_cta(getattr(self, "step_cta", "auto"))     # "auto" | "narrow" | "wide" | "split" (gmc_set_step_cta)
        lc = st = bl = None
        if want_caches:
            lc, st, bl = self._device_caches(n_steps)
        self.ctx.run(self.bed, self.mcres, self.ssq, self.seeds, self.iteration, n_steps, lc, st, bl, 0, self.resampled,
                     resync_every)
        self.iteration += n_steps
        if not want_caches:
            return None
        res = lc.cpu().numpy(), st.cpu().numpy(), bl.cpu().numpy()
        self.ctx.check_flag()
        return res

    def advance_into(self, n_steps, resync_every=4096, out=None):
        """The reference's per-chain outputs for a run of n_steps+1 iterations (index 0 = initial state, MCMC.py:1193-1199)
        as stacked arrays; results land in the pinned tensors of `out` when given (one D2H copy each)."""
        torch = self.torch
        n = n_steps + 1
        self.ctx.set_step_cta(getattr(self, "step_cta", "auto"))
        lc, st, bl = self._device_caches(n)
        lc[:, 0] = self._loss_now()
        st[:, 0] = 0
        bl[:, 0] = -1
        self.ctx.run(self.bed, self.mcres, self.ssq, self.seeds, self.iteration, n_steps, lc, st, bl, 1, self.resampled,
                     resync_every)
        self.iteration += n_steps
        if out is not None:
            out["bed"].copy_(self.bed, non_blocking=True)
            out["loss"].copy_(lc, non_blocking=True)
            out["steps"].copy_(st, non_blocking=True)
            out["blocks"].copy_(bl, non_blocking=True)
            if "resampled" in out and self.resampled is not None:
                self._copy_counts(out["resampled"], 0, self.C, n_steps)
            torch.cuda.current_stream().synchronize()
            self.ctx.check_flag()
            return dict(out)
        blocks = bl.cpu().numpy().astype(np.float64)
        blocks[:, 0, :] = np.nan                      # MCMC.py:1169: row 0 of blocks_cache stays NaN
        return dict(bed=self.bed.cpu().numpy(), loss=lc.cpu().numpy(), steps=st.cpu().numpy(), blocks=blocks)

    def run_pipelined(self, host_beds, keys, n_steps, out, groups=4, iter0=1, resync_every=4096, wait=True):
        """End-to-end run with host<->device transfers overlapped with compute: the chains are split into `groups` ranges,
        each on its own CUDA stream (H2D of the beds -> residual+loss -> n_steps fused iterations -> D2H of the results).
        The ranges start staggered by their H2D copies, so the D2H of one range overlaps the compute of the others.
        host_beds / out[...] are pinned CPU tensors (out may hold "resampled": int32 [C,H,W]); results are identical to
        `reset` + `advance_into` (chains are independent and counter-addressed).  wait=False returns a PendingRun as soon
        as the work is queued (one cudaMemcpyAsync per buffer and range; nothing batched)."""
        torch = self.torch
        if tuple(host_beds.shape) != (self.C, self.H, self.W) or host_beds.dtype != torch.float64:
            raise GmcShapeError(f"initial beds must be float64 [{self.C},{self.H},{self.W}]")
        if "resampled" in out and self.resampled is None:
            raise GmcError("out['resampled'] given but the ChainBatch was created with track_resampled=False")
        n = n_steps + 1
        lc, st, bl = self._device_caches(n)
        self.seeds = keys_tensor(keys, self.dev)
        den = torch.full((1,), 2 * self.chain.sigma_mc ** 2, dtype=torch.float64, device=self.dev)
        groups = max(1, min(int(groups), self.C))
        if not hasattr(self, "_streams") or len(self._streams) != groups:
            # Two streams per range.  The step kernels of the steps in flight fill every SM's register file, so the small
            # kernels in front of a range's step kernel (zeroing, the residual stencil of the uploaded beds, the row-0
            # fills) can only run in slots that a retiring step CTA frees; on HIGH-priority streams they are dispatched
            # before the pending step CTAs of other ranges, so the next step's kernels are ready long before the current
            # step ends (measured with GMC_TRACE_PIPELINE: without priorities they became ready 48 ms after their upload
            # was queued and the pipeline ran at 48.6 ms per step instead of ~40).
            lo, hi = 0, 0
            try:
                lo, hi = torch.cuda.Stream.priority_range()          # (least, greatest): e.g. (0, -5)
            except Exception:
                pass
            self._streams = [torch.cuda.Stream(device=self.dev, priority=lo) for _ in range(groups)]
            self._pre_streams = [torch.cuda.Stream(device=self.dev, priority=hi) for _ in range(groups)]
        main = torch.cuda.current_stream()
        bounds = [(g * self.C) // groups for g in range(groups + 1)]
        # several launches share the GPU (chain ranges, or further steps in flight when wait=False): 512-thread CTAs of
        # different launches cannot co-reside on an SM, so keep the 256-thread step kernel
        self.ctx.set_step_cta("narrow" if (groups > 1 or not wait) else "auto")
        # debug (GMC_TRACE_PIPELINE=1): timing events per range - upload queued / kernel queued / kernel done / download done
        trace = self._trace = [] if os.environ.get("GMC_TRACE_PIPELINE") else None

        def mark(g, what):
            if trace is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                trace.append((g, what, ev))
        for g, strm in enumerate(self._streams):
            a, b = bounds[g], bounds[g + 1]
            if a == b:
                continue
            pre = self._pre_streams[g]
            pre.wait_stream(main)
            with torch.cuda.stream(pre):
                mark(g, 0)
                self.bed[a:b].copy_(host_beds[a:b], non_blocking=True)
                if self.resampled is not None:
                    self.resampled[a:b].zero_()
                self.ctx.residual_loss_range(self.bed[a:b], self.mcres[a:b], self._loss[a:b], self.ssq[a:b], a)
                lc[a:b, 0] = self.ssq[a:b] / den
                st[a:b, 0] = 0
                bl[a:b, 0] = -1
                mark(g, 1)
            strm.wait_stream(pre)
            with torch.cuda.stream(strm):                         # low priority: the step kernel only
                self.ctx.run(self.bed[a:b], self.mcres[a:b], self.ssq[a:b], self.seeds[a:b], iter0, n_steps, lc[a:b], st[a:b],
                             bl[a:b], 1, None if self.resampled is None else self.resampled[a:b], resync_every)
                mark(g, 2)
            pre.wait_stream(strm)
            with torch.cuda.stream(pre):                          # high priority again: narrowing kernel and the downloads
                out["bed"][a:b].copy_(self.bed[a:b], non_blocking=True)
                out["loss"][a:b].copy_(lc[a:b], non_blocking=True)
                out["steps"][a:b].copy_(st[a:b], non_blocking=True)
                out["blocks"][a:b].copy_(bl[a:b], non_blocking=True)
                if "resampled" in out:
                    self._copy_counts(out["resampled"], a, b, n_steps)
                mark(g, 3)
        self.iteration = int(iter0) + n_steps
        pending = PendingRun(self, out)
        return pending.wait() if wait else pending

    def step_injected(self, fields, centres, us):
        """One step per chain with injected proposals.  Returns (accepted[C] bool, loss[C])."""
        torch = self.torch
        hmax = max(f.shape[0] for f in fields)
        wmax = max(f.shape[1] for f in fields)
        stride = hmax * wmax
        fbuf = np.zeros((self.C, stride))
        hw = np.zeros((self.C, 2), dtype=np.int32)
        for c, f in enumerate(fields):
            fbuf[c, :f.size] = np.ascontiguousarray(f, dtype=np.float64).ravel()
            hw[c] = f.shape
        dev = self.dev
        acc = torch.empty(self.C, dtype=torch.uint8, device=dev)
        loss = torch.empty(self.C, dtype=torch.float64, device=dev)
        self.ctx.step_injected(self.bed, self.mcres, self.ssq, torch.as_tensor(fbuf).to(dev), torch.as_tensor(hw).to(dev),
                               torch.as_tensor(np.asarray(centres, dtype=np.int32).reshape(self.C, 2)).to(dev),
                               torch.as_tensor(np.asarray(us, dtype=np.float64)).to(dev), hmax, wmax, acc, loss, None,
                               self.resampled)
        self.iteration += 1
        return acc.cpu().numpy().astype(bool), loss.cpu().numpy()
