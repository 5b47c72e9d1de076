// stencil_tma.cuh — K2/K3 on the Blackwell copy engine: the bed tiles of the chains stream through shared memory as TMA
// tensor tiles (cp.async.bulk.tensor.3d over bed[C][H][W]) on a full/empty mbarrier ring, the residual tile leaves through
// 16-byte streaming stores or, as a run-time option, TMA tensor stores (GMC_RS_TMA_STORE=1).  Included by residual.cu; same arithmetic, same
// rounding rules and the same partial-sum layout as residual_kernel there, which remains the path for odd W / unaligned
// bases (TMA needs 16-byte global strides).
//
// Geometry: a CTA (4 warps) owns an 8 x 64 cell tile and walks the chains c = z, z+G, ...; per chain ONE thread issues ONE
// tensor copy of the (8+2) x 68 box that starts at grid (row ty0-1, column tx0-2) — the two extra columns keep every lane's
// cell pair 16-byte aligned in shared memory; cells outside the grid arrive as zeros (TMA out-of-bound fill), so the ring
// needs no predicates and no clearing.  A warp owns 2 rows, a lane 2 columns; the lane's chain-independent operands
// (surf, velx, vely, dhdt, smb of its cells and halo rows) are read once from global memory into registers.
// Division: q = RN(x r), q' = fma(fma(-d, q, x), r, q) as in residual.cu (div_const); here the exponent-range guard is
// folded into ONE unsigned max per quotient (2*hi - LO as unsigned wraps for zero/tiny and exceeds the window for
// huge/inf/nan), checked once per iteration; zeros, infinities and NaNs take q itself (exact), anything else the true
// division, in a cold block.
#pragma once

#include <cuda.h>   // CUtensorMap and the encoder's enums (the encoder itself is resolved at run time, see r2_encode)

#ifndef R2_STAGES
#define R2_STAGES 8            // ring depth (tensor tiles in flight per CTA = R2_STAGES - R2_LAG)
#endif
#ifndef R2_LOSS_STAGES
#define R2_LOSS_STAGES 6       // ring depth of the loss-only variant (smaller ring, more CTAs per SM)
#endif
#ifndef R2_LOSS_MIN_CTAS
#define R2_LOSS_MIN_CTAS 5     // loss-only variant: 20 chain-independent doubles per lane, <= 96 registers
#endif
#ifndef R2_LIN4
#define R2_LIN4 1             // loss-only variant: a lane owns 4 rows x 1 column (1) or 2 rows x 2 columns (0)
#endif
#ifndef R2_LAG
#define R2_LAG 2               // the producer refills the stage consumed R2_LAG iterations ago
#endif
#ifndef R2_MIN_CTAS
#define R2_MIN_CTAS 3          // <= 168 registers: the lane's 40 chain-independent doubles stay in registers
#endif
#define R2_WARPS 4
#define R2_RW 2
#define R2_TH (R2_WARPS * R2_RW)                 // 8 rows per CTA tile
#define R2_TW 64
#define R2_BOXW (R2_TW + 4)                      // 68 columns: [pad, haloL, 64 cells, haloR, pad]
#define R2_BOXH (R2_TH + 2)
#ifndef R2_FETCH_ROWS
#define R2_FETCH_ROWS R2_BOXH    // timing experiments only: fewer rows per tensor copy (wrong results) shows what the halo traffic costs
#endif
#define R2_BOX_BYTES (R2_FETCH_ROWS * R2_BOXW * 8)     // 5440
#define R2_STAGE_BYTES ((R2_BOXH * R2_BOXW * 8 + 127) / 128 * 128)
#define R2_THREADS (R2_WARPS * 32)
#define R2_OUT_BUFS 4                            // per-warp staging buffers of the tensor stores
#define R2_LS_IT 8                               // loss partials are transposed through shared memory every 8 chains
#ifdef R2_LS_OCTETS                              // A/B: the previous scratch layout (pitch 33, eight contiguous entries per lane)
#define R2_LS_PITCH 33
#define R2_LS_IDX(rr, qd, j) ((rr) * 33 + (qd) * 8 + (j))
#else
#define R2_LS_IDX(rr, qd, j) ((rr) * R2_LS_PITCH + (qd) + 4 * (j))
#define R2_LS_PITCH 36                           // row pitch of that scratch (doubles), = 4 mod 16: lane (rr, qd) sums the entries qd + 4 j of
                                                 // row rr, so the 16 lanes of a half-warp read 16 different 8-byte banks (4 rr + qd); with
                                                 // pitch 33 and contiguous octets per lane every read was a 2-way bank conflict
#endif
#define R2_DIV_LO2 (2u * 0x05d00000u)            // 2 * hi word of 2^-930
#define R2_DIV_WIN (2u * (0x7a100000u - 0x05d00000u))

struct R2Layout {                                // byte offsets inside the dynamic shared memory (128-byte aligned base)
    int bars, out, lsum, total;
};
__host__ __device__ constexpr int r2_stages(bool write_res) { return write_res ? R2_STAGES : R2_LOSS_STAGES; }
__host__ __device__ inline R2Layout r2_layout(bool tma_store, bool do_loss, int n_stages) {
    const bool write_res = tma_store;
    R2Layout L;
    int off = n_stages * R2_STAGE_BYTES;
    L.bars = off;
    off += 2 * n_stages * 8;
    off = (off + 127) / 128 * 128;
    L.out = off;
    if (write_res) off += R2_WARPS * R2_OUT_BUFS * R2_RW * R2_TW * 8;
    L.lsum = off;
    if (do_loss) off += R2_WARPS * R2_LS_IT * R2_LS_PITCH * 8;
    L.total = off;
    return L;
}

__device__ __forceinline__ unsigned r2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void r2_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void r2_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void r2_mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void r2_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tR2_WAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra R2_DONE_%=;\n\t"
        "bra R2_WAIT_%=;\n\tR2_DONE_%=:\n\t}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void r2_tma_load3(unsigned dst, const CUtensorMap* tm, int x, int y, int z, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(tm), "r"(x), "r"(y), "r"(z), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void r2_tma_store3(const CUtensorMap* tm, int x, int y, int z, unsigned src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(x), "r"(y),
                 "r"(z), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void r2_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void r2_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void r2_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// cold: a quotient whose fast-path guard fired.  +-0 / d, +-inf / d and nan / d equal x * (1/d) exactly (checked inline:
// exact zeros are common in real velocity grids); tiny and huge finite quotients take the true division, out of line.
static __device__ __noinline__ double r2_div_true(double x, double d) { return __ddiv_rn(x, d); }
__device__ __forceinline__ double r2_div_fix(double x, double d, double q) {
    const unsigned e = (unsigned)(__double2hiint(x) << 1);
    if ((e | (unsigned)__double2loint(x)) == 0u || e >= 0xffe00000u) return q;
    return r2_div_true(x, d);
}

struct R2Lane {                                   // loop-invariant state of one lane
    double2 sf[R2_RW + 2], vy[R2_RW + 2];         // surf, vely: rows i0-1 .. i0+2, the lane's two columns
    double2 vx[R2_RW], dh[R2_RW], sm[R2_RW];      // velx, dhdt, smb: rows i0, i0+1
    double sfl[R2_RW], vxl[R2_RW], sfr[R2_RW], vxr[R2_RW];   // surf, velx of the columns left / right of the lane's pair
};

// One chain of one warp: 2 rows x 64 columns from the staged box at shared address `sa` (this lane's first cell of the warp's
// halo row).  No shuffles and no lane specialisation: every lane reads its pair (16 B) of four rows plus the two
// neighbouring cells (8 B each) of its two rows and forms their x-fluxes itself.
// EXACT = false (the loss-only variant, nothing is written back): the residual only feeds the masked sum of squares, whose
// contract is 1e-9 relative, so the quotients are plain products with the rounded reciprocal (<= 1 ulp each, no FMA
// correction, no range guard), (dhdt - smb) is taken pre-combined from L.dh, and where both quotients share the divisor
// 2 res (interior CTAs) the cell's residual is ONE fma: (numx + numy) * r + (dhdt - smb).  55 instead of 87 FP64
// instructions per warp-iteration; the loss agrees with the exact path to ~1e-15 relative.
template <bool DO_LOSS, bool EDGE, bool EXACT>
__device__ __forceinline__ void r2_rows(const R2Lane& L, const GmcDev& d, unsigned sa, double r_res, double r_two_res, bool xl_edge,
                                        bool xr_edge, int k_top, int k_bot, const bool (&vrow)[R2_RW], unsigned mcbits,
                                        double2 (&r)[R2_RW], double& acc, unsigned release_bar, int lane) {
    double2 bd[R2_RW + 2];
    double bl[R2_RW], br[R2_RW];
#pragma unroll
    for (int k = 0; k < R2_RW + 2; ++k)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(bd[k].x), "=d"(bd[k].y) : "r"(sa + k * (R2_BOXW * 8)));
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bl[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) - 8));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(br[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) + 16));
    }
    __syncwarp();
    if (lane == 0) r2_mbar_arrive(release_bar);                    // every lane has its operands: the stage may be refilled

    double fy0[R2_RW + 2], fy1[R2_RW + 2], fx0[R2_RW], fx1[R2_RW];
#pragma unroll
    for (int k = 0; k < R2_RW + 2; ++k) {
        const double t0 = sub_rn(L.sf[k].x, bd[k].x), t1 = sub_rn(L.sf[k].y, bd[k].y);
        fy0[k] = mul_rn(L.vy[k].x, t0);
        fy1[k] = mul_rn(L.vy[k].y, t1);
        if (k >= 1 && k <= R2_RW) {
            fx0[k - 1] = mul_rn(L.vx[k - 1].x, t0);
            fx1[k - 1] = mul_rn(L.vx[k - 1].y, t1);
        }
    }
    if (EDGE) {                                    // np.gradient's one-sided rows: duplicate the edge flux, divide by res
#pragma unroll
        for (int k = 0; k < R2_RW; ++k) {
            if (k == k_top) { fy0[k] = fy0[k + 1]; fy1[k] = fy1[k + 1]; }
            if (k == k_bot) { fy0[k + 2] = fy0[k + 1]; fy1[k + 2] = fy1[k + 1]; }
        }
    }
    double num[R2_RW][4], quo[R2_RW][4];
    unsigned guard = 0;
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        double fl = mul_rn(L.vxl[k], sub_rn(L.sfl[k], bl[k]));
        double fr = mul_rn(L.vxr[k], sub_rn(L.sfr[k], br[k]));
        double rdx0 = r_two_res, rdx1 = r_two_res, rdy = r_two_res, dnx0 = d.two_res, dnx1 = d.two_res, dny = d.two_res;
        if (EDGE) {
            if (xl_edge) { fl = fx0[k]; rdx0 = r_res; dnx0 = d.res; }          // column 0:   (f[1] - f[0]) / res
            if (xr_edge) { fr = fx1[k]; rdx1 = r_res; dnx1 = d.res; }          // column W-1: (f[W-1] - f[W-2]) / res
            if (k == k_top || k == k_bot) { rdy = r_res; dny = d.res; }
        }
        num[k][0] = sub_rn(fx1[k], fl);
        num[k][1] = sub_rn(fr, fx0[k]);
        num[k][2] = sub_rn(fy0[k + 2], fy0[k]);
        num[k][3] = sub_rn(fy1[k + 2], fy1[k]);
        const double rd[4] = {rdx0, rdx1, rdy, rdy}, dn[4] = {dnx0, dnx1, dny, dny};
        if (!EXACT) {
            if (EDGE) {                                            // per-cell divisors: two fmas per cell, see below
                quo[k][0] = rdx0; quo[k][1] = rdx1; quo[k][2] = quo[k][3] = rdy;
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double q = mul_rn(num[k][j], rd[j]);
            quo[k][j] = fma(fma(-dn[j], q, num[k][j]), rd[j], q);
            const unsigned e = (unsigned)(__double2hiint(q) << 1) - R2_DIV_LO2;
            if (!EDGE || vrow[k]) guard = max(guard, e);
        }
    }
    if (EXACT && guard > R2_DIV_WIN) {                             // zero, tiny, huge, inf or nan somewhere (cold)
#pragma unroll
        for (int k = 0; k < R2_RW; ++k) {
            double dnx0 = d.two_res, dnx1 = d.two_res, dny = d.two_res, rdx0 = r_two_res, rdx1 = r_two_res, rdy = r_two_res;
            if (EDGE) {
                if (xl_edge) { dnx0 = d.res; rdx0 = r_res; }
                if (xr_edge) { dnx1 = d.res; rdx1 = r_res; }
                if (k == k_top || k == k_bot) { dny = d.res; rdy = r_res; }
            }
            const double dn[4] = {dnx0, dnx1, dny, dny}, rd[4] = {rdx0, rdx1, rdy, rdy};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double q = mul_rn(num[k][j], rd[j]);
                if ((unsigned)(__double2hiint(q) << 1) - R2_DIV_LO2 > R2_DIV_WIN) quo[k][j] = r2_div_fix(num[k][j], dn[j], q);
            }
        }
    }
    double v[R2_RW][2];
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        if (EXACT) {
            r[k].x = sub_rn(add_rn(add_rn(quo[k][0], quo[k][2]), L.dh[k].x), L.sm[k].x);
            r[k].y = sub_rn(add_rn(add_rn(quo[k][1], quo[k][3]), L.dh[k].y), L.sm[k].y);
        } else if (EDGE) {                                         // L.dh holds dhdt - smb; quo holds the reciprocals
            r[k].x = fma(num[k][0], quo[k][0], fma(num[k][2], quo[k][2], L.dh[k].x));
            r[k].y = fma(num[k][1], quo[k][1], fma(num[k][3], quo[k][3], L.dh[k].y));
        } else {                                                   // one divisor: (numx + numy) / (2 res) + (dhdt - smb)
            r[k].x = fma(num[k][0] + num[k][2], r_two_res, L.dh[k].x);
            r[k].y = fma(num[k][1] + num[k][3], r_two_res, L.dh[k].y);
        }
        if (DO_LOSS) {
            // four independent masked squares, then a tree: a sequential `if (...) acc += r*r` chain compiled to a serial
            // select-and-move sequence twice as long (bits of rows / columns outside the grid are 0; nan cells count 0)
            const double s0 = mul_rn(r[k].x, r[k].x), s1 = mul_rn(r[k].y, r[k].y);
            v[k][0] = (((mcbits >> (2 * k)) & 1u) && s0 == s0) ? s0 : 0.0;
            v[k][1] = (((mcbits >> (2 * k + 1)) & 1u) && s1 == s1) ? s1 : 0.0;
        }
    }
    acc = 0.0;
    if (DO_LOSS) {
        static_assert(R2_RW == 2, "the loss tree is written for two rows per warp");
        acc = add_rn(add_rn(v[0][0], v[0][1]), add_rn(v[1][0], v[1][1]));
    }
}

// ---- loss-only variant (nothing written back): the residual as a 5-point LINEAR form of the bed ---------------------
// With f = v (surf - bed), the reference's cell residual (fxR - fxL) rdx + (fyD - fyU) rdy + dhdt - smb is
//     r = K + cR bed_R + cL bed_L + cD bed_D + cU bed_U,
//     K  = dhdt - smb + rdx (vxR surfR - vxL surfL) + rdy (vyD surfD - vyU surfU),
//     cR = -rdx vxR,  cL = rdx vxL,  cD = -rdy vyD,  cU = rdy vyU
// (R/L/D/U = the neighbours np.gradient uses: the cell itself on the one-sided edge rows / columns; rdx, rdy = 1/(2 res)
// or 1/res there).  K and the four coefficients do not depend on the chain: a lane forms them once for its four cells (20
// doubles instead of the 40 operands of the flux form), and a chain costs 4 fma + 1 square per cell — 24 FP64 instructions
// per warp-iteration instead of 55.  The residual of this form differs from the flux form by rounding only (a few ulp of
// the flux magnitude); it feeds nothing but the masked sum of squares, whose contract is 1e-9 relative (measured: ~1e-15).
// Cells outside the loss mask (and outside the grid) get K = c = 0, so their square is exactly 0 and the common case needs
// no per-cell select; a NaN anywhere in the lane's four squares takes the cold per-cell path (nan cells count 0).
struct R2Lin {
    double K[R2_RW][2], cL[R2_RW][2], cR[R2_RW][2], cU[R2_RW][2], cD[R2_RW][2];
};

template <bool EDGE>
__device__ __forceinline__ void r2_lin_setup(R2Lin& Q, const R2Lane& L, double r_res, double r_two_res, bool xl_edge, bool xr_edge,
                                             int k_top, int k_bot, unsigned mcbits) {
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        // neighbours of the lane's two cells in row k: [surf, vel] left / right (x) and up / down (y)
        double sL[2] = {L.sfl[k], L.sf[k + 1].x}, vL[2] = {L.vxl[k], L.vx[k].x};
        double sR[2] = {L.sf[k + 1].y, L.sfr[k]}, vR[2] = {L.vx[k].y, L.vxr[k]};
        double sU[2] = {L.sf[k].x, L.sf[k].y}, vU[2] = {L.vy[k].x, L.vy[k].y};
        double sD[2] = {L.sf[k + 2].x, L.sf[k + 2].y}, vD[2] = {L.vy[k + 2].x, L.vy[k + 2].y};
        double rdx[2] = {r_two_res, r_two_res}, rdy = r_two_res;
        if (EDGE) {
            if (xl_edge) { sL[0] = L.sf[k + 1].x; vL[0] = L.vx[k].x; rdx[0] = r_res; }
            if (xr_edge) { sR[1] = L.sf[k + 1].y; vR[1] = L.vx[k].y; rdx[1] = r_res; }
            if (k == k_top) { sU[0] = L.sf[k + 1].x; sU[1] = L.sf[k + 1].y; vU[0] = L.vy[k + 1].x; vU[1] = L.vy[k + 1].y; }
            if (k == k_bot) { sD[0] = L.sf[k + 1].x; sD[1] = L.sf[k + 1].y; vD[0] = L.vy[k + 1].x; vD[1] = L.vy[k + 1].y; }
            if (k == k_top || k == k_bot) rdy = r_res;
        }
        const double dhm[2] = {L.dh[k].x, L.dh[k].y};              // dhdt - smb (combined by the caller)
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const bool on = (mcbits >> (2 * k + x)) & 1u;
            const double gx = fma(vR[x], sR[x], -(vL[x] * sL[x])), gy = fma(vD[x], sD[x], -(vU[x] * sU[x]));
            Q.K[k][x] = on ? fma(gx, rdx[x], fma(gy, rdy, dhm[x])) : 0.0;
            Q.cR[k][x] = on ? -(rdx[x] * vR[x]) : 0.0;
            Q.cL[k][x] = on ? rdx[x] * vL[x] : 0.0;
            Q.cD[k][x] = on ? -(rdy * vD[x]) : 0.0;
            Q.cU[k][x] = on ? rdy * vU[x] : 0.0;
        }
    }
}

template <bool EDGE>
__device__ __forceinline__ void r2_rows_lin(const R2Lin& Q, unsigned sa, bool xl_edge, bool xr_edge, int k_top, int k_bot,
                                            double& acc, unsigned release_bar, int lane) {
    double2 bd[R2_RW + 2];
    double bl[R2_RW], br[R2_RW];
#pragma unroll
    for (int k = 0; k < R2_RW + 2; ++k)
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(bd[k].x), "=d"(bd[k].y) : "r"(sa + k * (R2_BOXW * 8)));
    // (x-neighbours through warp shuffles instead of these 8-byte loads: 67 % instead of 74 % of the HBM peak at 4096 x 500^2 -
    // profiles/r2/stencil_lin_ab.txt; the kernel sits at the shared-memory / MIO pipe's limit and shuffles queue there too)
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bl[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) - 8));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(br[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) + 16));
    }
    __syncwarp();
    if (lane == 0) r2_mbar_arrive(release_bar);                    // every lane has its operands: the stage may be refilled
    double sq[R2_RW][2];
#ifdef R2_EXPERIMENT_NOCOMPUTE                                     // timing experiments only: the tile stream without the arithmetic
    acc = bd[1].x + bl[0] + br[1] + bd[3].y + bd[0].x + bd[2].y;
    return;
#endif
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) {
        double bL[2] = {bl[k], bd[k + 1].x}, bR[2] = {bd[k + 1].y, br[k]};
        double bU[2] = {bd[k].x, bd[k].y}, bD[2] = {bd[k + 2].x, bd[k + 2].y};
        if (EDGE) {
            if (xl_edge) bL[0] = bd[k + 1].x;
            if (xr_edge) bR[1] = bd[k + 1].y;
            if (k == k_top) { bU[0] = bd[k + 1].x; bU[1] = bd[k + 1].y; }
            if (k == k_bot) { bD[0] = bd[k + 1].x; bD[1] = bd[k + 1].y; }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const double r = fma(Q.cR[k][x], bR[x], fma(Q.cL[k][x], bL[x], fma(Q.cD[k][x], bD[x], fma(Q.cU[k][x], bU[x], Q.K[k][x]))));
            sq[k][x] = mul_rn(r, r);
        }
    }
    static_assert(R2_RW == 2, "the loss tree is written for two rows per warp");
    acc = add_rn(add_rn(sq[0][0], sq[0][1]), add_rn(sq[1][0], sq[1][1]));
    if (acc != acc) {                                              // cold: a nan cell somewhere; nan cells count 0, same tree
#pragma unroll
        for (int k = 0; k < R2_RW; ++k)
#pragma unroll
            for (int x = 0; x < 2; ++x) sq[k][x] = (sq[k][x] == sq[k][x]) ? sq[k][x] : 0.0;
        acc = add_rn(add_rn(sq[0][0], sq[0][1]), add_rn(sq[1][0], sq[1][1]));
    }
}

// ---- the same linear form with a lane owning 4 rows x 1 column (warp: 4 rows x 32 columns, CTA tile 8 x 64 as 2 x 2 warps) ---
// The variant is bound by shared-memory wavefronts, not by arithmetic (ncu: 68 % of the pipe's peak, short_scoreboard the top
// stall).  With 2 x 2 cells per lane the x-neighbours are 8-byte loads at a 16-byte lane stride - 4 wavefronts for 256
// useful bytes - and a warp reads 4 rows to produce 2: 32 wavefronts per 128 cells.  With 4 x 1 cells per lane every access
// is a fully coalesced 8-byte load (2 wavefronts) and a warp reads 6 rows to produce 4: 6 + 8 loads = 28 wavefronts per 128
// cells, same 20 coefficients per lane.  The coefficients are formed from global memory in the CTA prologue.
struct R2Lin4 {
    double K[4], cL[4], cR[4], cU[4], cD[4];
};

template <bool EDGE>
__device__ __forceinline__ void r2_lin4_setup(R2Lin4& Q, const GmcDev& d, int i0, int j, double r_res, double r_two_res) {
    const int H = d.H, W = d.W;
    // 16-byte loads of the packed statics ({surf, velx}, {surf, vely}, {dhdt, smb} per cell, built by gmc_set_static for the
    // step kernel): the lane's column of {surf, vely} for rows i0-1 .. i0+4 once, then per cell its two x-neighbours and its
    // own {dhdt, smb} - 22 loads per lane.  Rows / columns outside the grid are never used: the edge rules replace them.
    const bool vcol = !EDGE || j < W;
    double2 cy[6];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        const int i = i0 - 1 + m;
        cy[m] = (vcol && (!EDGE || (i >= 0 && i < H))) ? __ldg(d.sy + (int64_t)i * W + j) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = i0 + k;
        double K = 0.0, cL = 0.0, cR = 0.0, cU = 0.0, cD = 0.0;
        if (!EDGE || (i < H && vcol)) {
            const int64_t idx = (int64_t)i * W + j;
            if (__ldg(d.flags + idx) & FLAG_MC) {
                const bool xl = EDGE && j == 0, xr = EDGE && j == W - 1, top = EDGE && i == 0, bot = EDGE && i == H - 1;
                const double2 xL = __ldg(d.sv + (xl ? idx : idx - 1)), xR = __ldg(d.sv + (xr ? idx : idx + 1));   // {surf, velx}
                const double2 yU = top ? cy[k + 1] : cy[k], yD = bot ? cy[k + 1] : cy[k + 2];                       // {surf, vely}
                const double2 hs = __ldg(d.ds + idx);                                                             // {dhdt, smb}
                const double rdx = (xl || xr) ? r_res : r_two_res, rdy = (top || bot) ? r_res : r_two_res;
                const double dhm = hs.x - hs.y;
                const double gx = fma(xR.y, xR.x, -(xL.y * xL.x)), gy = fma(yD.y, yD.x, -(yU.y * yU.x));
                K = fma(gx, rdx, fma(gy, rdy, dhm));
                cR = -(rdx * xR.y);
                cL = rdx * xL.y;
                cD = -(rdy * yD.y);
                cU = rdy * yU.y;
            }
        }
        Q.K[k] = K;
        Q.cL[k] = cL;
        Q.cR[k] = cR;
        Q.cU[k] = cU;
        Q.cD[k] = cD;
    }
}

template <bool EDGE>
__device__ __forceinline__ void r2_rows_lin4(const R2Lin4& Q, unsigned sa, bool xl_edge, bool xr_edge, int k_top, int k_bot,
                                             double& acc, unsigned release_bar, int lane) {
    double bd[6], bl[4], br[4];
#pragma unroll
    for (int m = 0; m < 6; ++m) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bd[m]) : "r"(sa + m * (R2_BOXW * 8)));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(bl[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) - 8));
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(br[k]) : "r"(sa + (k + 1) * (R2_BOXW * 8) + 8));
    }
    __syncwarp();
    if (lane == 0) r2_mbar_arrive(release_bar);                    // every lane has its operands: the stage may be refilled
    double sq[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double bL = bl[k], bR = br[k], bU = bd[k], bD = bd[k + 2];
        if (EDGE) {
            if (xl_edge) bL = bd[k + 1];
            if (xr_edge) bR = bd[k + 1];
            if (k == k_top) bU = bd[k + 1];
            if (k == k_bot) bD = bd[k + 1];
        }
        const double r = fma(Q.cR[k], bR, fma(Q.cL[k], bL, fma(Q.cD[k], bD, fma(Q.cU[k], bU, Q.K[k]))));
        sq[k] = mul_rn(r, r);
    }
    acc = add_rn(add_rn(sq[0], sq[1]), add_rn(sq[2], sq[3]));
    if (acc != acc) {                                              // cold: a nan cell somewhere; nan cells count 0, same tree
#pragma unroll
        for (int k = 0; k < 4; ++k) sq[k] = (sq[k] == sq[k]) ? sq[k] : 0.0;
        acc = add_rn(add_rn(sq[0], sq[1]), add_rn(sq[2], sq[3]));
    }
}

template <bool WRITE_RES, bool DO_LOSS, bool EDGE, bool TMA_STORE>
__device__ __forceinline__ void r2_chain_loop(const CUtensorMap* tm_bed, const CUtensorMap* tm_out, const GmcDev& d,
                                              unsigned char* smem, double* __restrict__ res_all, double* __restrict__ partials,
                                              int n_tiles, int C, double r_res, double r_two_res) {
    const int H = d.H, W = d.W;
    const int tid = threadIdx.x, warp = tid >> 5;
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    constexpr int NS = r2_stages(WRITE_RES);        // ring depth of this variant
    const R2Layout lay = r2_layout(WRITE_RES && TMA_STORE, DO_LOSS, NS);
    const unsigned stage0 = r2_smem_u32(smem);
    const unsigned bars = stage0 + lay.bars;      // full[s] at bars + 8 s, empty[s] at bars + 8 (NS + s)
    const int tx0 = blockIdx.x * R2_TW, ty0 = blockIdx.y * R2_TH;
    const int i0 = ty0 + warp * R2_RW;            // first grid row of this warp
    const int c0 = tx0 + 2 * lane;                // first grid column of this lane
    const int G = gridDim.z;
    const int n_iter = (C - (int)blockIdx.z + G - 1) / G;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            r2_mbar_init(bars + 8 * s, 1);
            r2_mbar_init(bars + 8 * (NS + s), R2_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // prologue of the ring: NS - R2_LAG tiles in flight before anything else is fetched
    if (tid == 0) {
        for (int s = 0; s < NS - R2_LAG && s < n_iter; ++s) {
            r2_mbar_expect_tx(bars + 8 * s, R2_BOX_BYTES);
            r2_tma_load3(stage0 + s * R2_STAGE_BYTES, tm_bed, tx0 - 2, ty0 - 1, (int)blockIdx.z + s * G, bars + 8 * s);
        }
    }

    // ---- the lane's chain-independent operands, straight from global memory (one exposed latency, behind the tiles) ----
    R2Lane L;
    const bool vcol = !EDGE || c0 < W;
    constexpr bool LIN4 = !WRITE_RES && (R2_LIN4 != 0);           // loss-only variant with the 4 x 1 lane layout
    if (!LIN4) {
        const double2 z2 = make_double2(0.0, 0.0);
#pragma unroll
        for (int k = 0; k < R2_RW + 2; ++k) {
            const int i = i0 - 1 + k;
            const bool in = !EDGE || (vcol && i >= 0 && i < H);
            const int64_t idx = (int64_t)i * W + c0;
            L.sf[k] = in ? __ldg(reinterpret_cast<const double2*>(d.surf + idx)) : z2;
            L.vy[k] = in ? __ldg(reinterpret_cast<const double2*>(d.vely + idx)) : z2;
            if (k >= 1 && k <= R2_RW) {
                L.vx[k - 1] = in ? __ldg(reinterpret_cast<const double2*>(d.velx + idx)) : z2;
                L.dh[k - 1] = in ? __ldg(reinterpret_cast<const double2*>(d.dhdt + idx)) : z2;
                L.sm[k - 1] = in ? __ldg(reinterpret_cast<const double2*>(d.smb + idx)) : z2;
                if (!WRITE_RES) {                              // loss-only variant: see r2_rows, EXACT = false
                    L.dh[k - 1].x -= L.sm[k - 1].x;
                    L.dh[k - 1].y -= L.sm[k - 1].y;
                }
                const bool inl = !EDGE || (in && c0 > 0), inr = !EDGE || (in && c0 + 2 < W);
                L.sfl[k - 1] = inl ? __ldg(d.surf + idx - 1) : 0.0;
                L.vxl[k - 1] = inl ? __ldg(d.velx + idx - 1) : 0.0;
                L.sfr[k - 1] = inr ? __ldg(d.surf + idx + 2) : 0.0;
                L.vxr[k - 1] = inr ? __ldg(d.velx + idx + 2) : 0.0;
            }
        }
    }
    // loss-mask bits: bit 2k (+1) = row k, first (second) column
    unsigned mcbits = 0;
    if (DO_LOSS) {
#pragma unroll
        for (int k = 0; k < R2_RW; ++k) {
            const int i = i0 + k;
            if (!EDGE || (vcol && i < H)) {
                const uchar2 f = __ldg(reinterpret_cast<const uchar2*>(d.flags + (int64_t)i * W + c0));
                if (f.x & FLAG_MC) mcbits |= 1u << (2 * k);
                if (f.y & FLAG_MC) mcbits |= 1u << (2 * k + 1);
            }
        }
    }
    const bool xl_edge = EDGE && c0 == 0, xr_edge = EDGE && c0 + 1 == W - 1;
    const int k_top = (EDGE && i0 == 0) ? 0 : -1;
    const int k_bot = (EDGE && H - 1 >= i0 && H - 1 < i0 + R2_RW) ? H - 1 - i0 : -1;
    bool vrow[R2_RW];
#pragma unroll
    for (int k = 0; k < R2_RW; ++k) vrow[k] = !EDGE || (vcol && i0 + k < H);
    R2Lin Q;
    if (!WRITE_RES && !LIN4) r2_lin_setup<EDGE>(Q, L, r_res, r_two_res, xl_edge, xr_edge, k_top, k_bot, mcbits);   // L is dead after this
    static_assert(R2_WARPS == 4 && R2_TH == 8 && R2_TW == 64, "the 4 x 1 lane layout tiles 8 x 64 cells as 2 x 2 warps");
    R2Lin4 Q4;
    const int i04 = ty0 + 4 * (warp >> 1), j4 = tx0 + 32 * (warp & 1) + lane;      // first row / the column of this lane
    const bool xl4 = EDGE && j4 == 0, xr4 = EDGE && j4 == W - 1;
    const int kt4 = (EDGE && i04 == 0) ? 0 : -1;
    const int kb4 = (EDGE && H - 1 >= i04 && H - 1 < i04 + 4) ? H - 1 - i04 : -1;
    const unsigned sa4 = stage0 + (unsigned)((4 * (warp >> 1)) * R2_BOXW + 2 + 32 * (warp & 1) + lane) * 8u;
    if (LIN4) r2_lin4_setup<EDGE>(Q4, d, i04, j4, r_res, r_two_res);

    const unsigned lane_sa = stage0 + (unsigned)((warp * R2_RW) * R2_BOXW + 2 + 2 * lane) * 8u;   // stage 0, this lane's halo-row pair
    const int tile_id = (blockIdx.y * R2_WARPS + warp) * gridDim.x + blockIdx.x;
    const int64_t plane = (int64_t)H * W;
    double* po = WRITE_RES ? res_all + (int64_t)blockIdx.z * plane + (int64_t)i0 * W + c0 : nullptr;
    const int64_t cstride = (int64_t)G * plane;
    double* pp = partials + (int64_t)blockIdx.z * n_tiles + tile_id;
    const int64_t pstride = (int64_t)G * n_tiles;
    const unsigned out_base = stage0 + lay.out + warp * (R2_OUT_BUFS * R2_RW * R2_TW * 8);
    double* lsum = reinterpret_cast<double*>(smem + lay.lsum) + warp * (R2_LS_IT * R2_LS_PITCH);
    int zc = (int)blockIdx.z;                      // chain of the current iteration

    // The ring loop stays ROLLED: unrolled by the stage count the kernel was 40 KB of code per variant and ran at an
    // instruction-cache hit rate of 64 % (ncu: no_instruction the top stall); the running stage offsets cost ~5 instructions.
    unsigned phase = 0;                            // parity of this trip around the ring
    int s = 0;                                     // stage of the current iteration
    for (int it = 0; it < n_iter; ++it) {
        // ---- producer (warp 0): refill the stage consumed R2_LAG iterations ago with the tile of chain it+STAGES-LAG ----
        if (warp == 0) {
            if (lane == 0 && it + NS - R2_LAG < n_iter) {
                int ps = s - R2_LAG;
                unsigned pph = phase;
                if (ps < 0) { ps += NS; pph ^= 1u; }
                if (it >= R2_LAG) r2_mbar_wait(bars + 8 * (NS + ps), pph);             // all four warps released it
                r2_mbar_expect_tx(bars + 8 * ps, R2_BOX_BYTES);
                r2_tma_load3(stage0 + ps * R2_STAGE_BYTES, tm_bed, tx0 - 2, ty0 - 1, zc + (NS - R2_LAG) * G, bars + 8 * ps);
            }
            __syncwarp();
        }
        r2_mbar_wait(bars + 8 * s, phase);
        double2 r[R2_RW];
        double acc;
        if (WRITE_RES)
            r2_rows<DO_LOSS, EDGE, true>(L, d, lane_sa + s * R2_STAGE_BYTES, r_res, r_two_res, xl_edge, xr_edge, k_top, k_bot, vrow, mcbits, r, acc,
                                         bars + 8 * (NS + s), lane);
        else
            if (LIN4) r2_rows_lin4<EDGE>(Q4, sa4 + s * R2_STAGE_BYTES, xl4, xr4, kt4, kb4, acc, bars + 8 * (NS + s), lane);
            else r2_rows_lin<EDGE>(Q, lane_sa + s * R2_STAGE_BYTES, xl_edge, xr_edge, k_top, k_bot, acc, bars + 8 * (NS + s), lane);
        if (WRITE_RES) {
            if (TMA_STORE) {
                // the warp's 2 x 64 residual tile -> its staging buffer -> one tensor store (clipped at the grid edge by TMA)
                const unsigned ob = out_base + (unsigned)(it & (R2_OUT_BUFS - 1)) * (R2_RW * R2_TW * 8);
#pragma unroll
                for (int k = 0; k < R2_RW; ++k)
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ob + (unsigned)(k * R2_TW + 2 * lane) * 8u), "d"(r[k].x), "d"(r[k].y) : "memory");
                r2_fence_async();
                __syncwarp();
                if (lane == 0) {
                    if (!EDGE || i0 < H) r2_tma_store3(tm_out, tx0, i0, zc, ob);
                    r2_bulk_commit();
                    r2_bulk_wait_read<R2_OUT_BUFS - 1>();          // the buffer of the next iteration has been read out
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int k = 0; k < R2_RW; ++k)
                    if (!EDGE || vrow[k]) __stcs(reinterpret_cast<double2*>(po + (int64_t)k * W), r[k]);
                po += cstride;
            }
        }
        if (DO_LOSS) {
            // per-lane partial of this chain -> row (it mod 8) of the warp's scratch; after 8 chains the 8 x 32 block is
            // summed by all lanes (4 lanes per chain, fixed order) instead of a 5-step shuffle tree per chain
            const int row = it & (R2_LS_IT - 1);
            lsum[row * R2_LS_PITCH + lane] = acc;
            if (row == R2_LS_IT - 1 || it == n_iter - 1) {
                __syncwarp();
                const int rr = lane >> 2, qd = lane & 3;
                double t = 0.0;
#pragma unroll
                for (int j = 0; j < 8; ++j) t = add_rn(t, lsum[R2_LS_IDX(rr, qd, j)]);
                t = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 1));
                t = add_rn(t, __shfl_xor_sync(0xffffffffu, t, 2));
                if (qd == 0 && rr <= row) pp[(int64_t)(rr - row) * pstride] = t;
                __syncwarp();
            }
            pp += pstride;
        }
        zc += G;
        if (++s == NS) { s = 0; phase ^= 1u; }
    }
    if (WRITE_RES && TMA_STORE) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before the CTA exits
    }
}

template <bool WRITE_RES, bool DO_LOSS, bool TMA_STORE>
__global__ void __launch_bounds__(R2_THREADS, (WRITE_RES ? R2_MIN_CTAS : R2_LOSS_MIN_CTAS))
    residual_tma_kernel(const __grid_constant__ CUtensorMap tm_bed, const __grid_constant__ CUtensorMap tm_out, GmcDev d,
                        double* __restrict__ res_all, double* __restrict__ partials, int n_tiles, int C, double r_res,
                        double r_two_res) {
    extern __shared__ __align__(128) unsigned char r2_raw[];
    const int tx0 = blockIdx.x * R2_TW, ty0 = blockIdx.y * R2_TH;
    const bool interior = tx0 > 0 && tx0 + R2_TW < d.W && ty0 > 0 && ty0 + R2_TH < d.H;     // CTA-uniform
    if (interior) r2_chain_loop<WRITE_RES, DO_LOSS, false, TMA_STORE>(&tm_bed, &tm_out, d, r2_raw, res_all, partials, n_tiles, C, r_res, r_two_res);
    else r2_chain_loop<WRITE_RES, DO_LOSS, true, TMA_STORE>(&tm_bed, &tm_out, d, r2_raw, res_all, partials, n_tiles, C, r_res, r_two_res);
}

// ---- host: tensor maps ------------------------------------------------------------------------------------------
typedef CUresult (*r2_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static r2_encode_fn r2_encoder() {
    static r2_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<r2_encode_fn>(p);
    }
    return fn;
}

// [C][H][W] float64 tensor, box [1][box_h][box_w]; returns false when the driver entry point is unavailable or refuses
static bool r2_encode(CUtensorMap* tm, const double* base, int C, int H, int W, int box_w, int box_h) {
    r2_encode_fn enc = r2_encoder();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 8, (cuuint64_t)W * H * 8};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
