"""numpy emulation of libgmc's device RNG (csrc/common.cuh): Philox4x32-10, the uniform/normal transforms and the
per-iteration draw layout of run_kernel.  Test infrastructure: lets the free-running kernel be compared draw-for-draw
with the oracle."""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
STREAM_RF_SCALARS, STREAM_NOISE, STREAM_NUGGET, STREAM_CHAIN = 0, 1, 2, 3
STREAM_RM_MODE, STREAM_RM_AMP = 8, 9
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32(key: int, c0, c1, c2, c3):
    """Vectorised Philox4x32-10.  key: 64-bit int; c0..c3 broadcastable uint32 arrays.  Returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) & _M32 for x in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = key & 0xFFFFFFFF, (key >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _M32
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def _k53(hi, lo):
    return ((hi.astype(np.uint64) >> np.uint64(5)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))


def u01_open(hi, lo):
    k52 = ((hi.astype(np.uint64) >> np.uint64(6)) << np.uint64(26)) | (lo.astype(np.uint64) >> np.uint64(6))
    return (k52.astype(np.float64) + 0.5) * 2.0 ** -52


def u01_halfopen(hi, lo):
    return _k53(hi, lo).astype(np.float64) * 2.0 ** -53


def bounded(hi, lo, n):
    v = (int(hi) << 32) | int(lo)
    return (v * int(n)) >> 64


def box_muller(r):
    u1 = u01_open(r[0], r[1])
    u2 = u01_open(r[2], r[3])
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * u2            # sincospi(2 u2)
    return rad * np.cos(np.pi * ang), rad * np.sin(np.pi * ang)


def hermitian_noise(key, it, h, w):
    """Full-plane noise arrays (A, B) equivalent to the kernel's direct draw of the Hermitian half plane (step.cu,
    synth_field step 1): Re(ifft2((A + iB) sqrt(S))) equals the kernel's field up to rounding.

    Kernel: for kx in (0, w/2): X_h(ky,kx) = sqrt(S) (z0 + i z1)/sqrt(2) with (z0,z1) = BoxMuller(Philox(ky*w+kx));
    columns kx in {0, w/2}: same for 0 < ky < h/2 and X_h(h-ky) = conj(X_h(ky)); the four self-conjugate points are real,
    sqrt(S) z0.  Reference algebra: X_h(k) = sqrt(S)/2 ((A_k + A_-k) + i (B_k - B_-k)).
    """
    lo, hi = it & 0xFFFFFFFF, (it >> 32) & 0xFFFFFFFF
    e = np.arange(h * w, dtype=np.uint32)
    z0, z1 = box_muller(philox4x32(key, e, lo, hi, STREAM_NOISE))
    z0, z1 = z0.reshape(h, w), z1.reshape(h, w)
    A = np.zeros((h, w))
    B = np.zeros((h, w))
    n2 = w // 2
    r2 = np.sqrt(0.5)
    for ky in range(h):
        nky = (h - ky) % h
        for kx in range(n2 + 1):
            nkx = (w - kx) % w
            self_y = ky in (0, h // 2)
            if kx in (0, n2):
                if self_y:
                    A[ky, kx] = z0[ky, kx]
                elif ky < h // 2:
                    p, q = z0[ky, kx] * r2, z1[ky, kx] * r2
                    A[ky, kx] = A[nky, kx] = p
                    B[ky, kx], B[nky, kx] = q, -q
            else:
                p, q = z0[ky, kx] * r2, z1[ky, kx] * r2
                A[ky, kx] = A[nky, nkx] = p
                B[ky, kx], B[nky, nkx] = q, -q
    return A, B


def randmeth_modes(key, it, n_modes, model, range_x, range_y, angle_deg, nu):
    """(kx, ky, z1, z2) of the randomization-method field of iteration `it` (step.cu rm_mode): the radius quantile and
    direction come from one Philox block of stream RM_MODE, the two amplitudes from one block of stream RM_AMP."""
    from oracle import randmeth_oracle as R
    lo, hi = it & 0xFFFFFFFF, (it >> 32) & 0xFFFFFFFF
    m = np.arange(n_modes, dtype=np.uint32)
    a = philox4x32(key, m, lo, hi, STREAM_RM_MODE)
    kx, ky = R.grid_wave_vectors(model, u01_open(a[0], a[1]), u01_open(a[2], a[3]), range_x, range_y, angle_deg, nu)
    z1, z2 = box_muller(philox4x32(key, m, lo, hi, STREAM_RM_AMP))
    return kx, ky, z1, z2


def step_draws(key, it, n_pairs, pairs, fm, H, W, centre_cells, randmeth=False):
    """All random inputs of iteration `it` of one chain, in oracle terms.  randmeth: the proposal is the
    randomization-method field (anisotropy angle drawn, no spectral noise planes)."""
    lo, hi = it & 0xFFFFFFFF, (it >> 32) & 0xFFFFFFFF
    r0 = [int(x) for x in philox4x32(key, 0, lo, hi, STREAM_RF_SCALARS)]
    r1 = [int(x) for x in philox4x32(key, 1, lo, hi, STREAM_RF_SCALARS)]
    pick = bounded(r0[0], r0[1], n_pairs)
    f64 = np.float64
    u = lambda a, b: float(u01_halfopen(np.uint32(a), np.uint32(b)))       # noqa: E731
    scale = (f64(fm["scale_min"]) + (f64(fm["scale_max"]) - f64(fm["scale_min"])) * u(r0[2], r0[3])) / 3.0
    nug = 0.0 + f64(fm["nugget_max"]) * u(r1[0], r1[1])
    range_x = f64(fm["range_min_x"]) + (f64(fm["range_max_x"]) - f64(fm["range_min_x"])) * u(r1[2], r1[3])
    angle = 0.0
    if fm["isotropic"]:
        range_y = range_x
    else:
        r2 = [int(x) for x in philox4x32(key, 2, lo, hi, STREAM_RF_SCALARS)]
        range_y = f64(fm["range_min_y"]) + (f64(fm["range_max_y"]) - f64(fm["range_min_y"])) * u(r2[0], r2[1])
        angle = 180.0 * u(r2[2], r2[3])
    bw, bh = int(pairs[0, pick]), int(pairs[1, pick])
    e = np.arange(bh * bw, dtype=np.uint32)
    if randmeth:
        z_re = z_im = np.zeros(bh * bw)
    else:
        z_re, z_im = hermitian_noise(key, it, bh, bw)
    z_nug, _ = box_muller(philox4x32(key, e, lo, hi, STREAM_NUGGET))
    c0 = [int(x) for x in philox4x32(key, 0, lo, hi, STREAM_CHAIN)]
    c1 = [int(x) for x in philox4x32(key, 1, lo, hi, STREAM_CHAIN)]
    if centre_cells is not None:
        cell = int(centre_cells[bounded(c0[0], c0[1], len(centre_cells))])
        ix, iy = divmod(cell, W)
    else:
        ix, iy = bounded(c0[0], c0[1], H), bounded(c0[2], c0[3], W)
    return dict(pair=pick, scale=float(scale), nug=float(nug), range_x=float(range_x), range_y=float(range_y), angle=float(angle),
                z_re=z_re.reshape(bh, bw), z_im=z_im.reshape(bh, bw), z_nug=z_nug.reshape(bh, bw), idx_x=ix, idx_y=iy,
                u=u(c1[0], c1[1]))
