// residual.cu — K2/K3: batched full-grid mass-conservation residual and masked loss (the "stencil" of the metric).
//
// Reference semantics (Topography.py:592-600, MCMC.py:1041):
//   thick = surf - bed;  dx = np.gradient(velx*thick, res, axis=1);  dy = np.gradient(vely*thick, res, axis=0)
//   out = ((dx + dy) + dhdt) - smb;   loss = nansum(out[mask==1]^2) / (2 sigma^2)
// np.gradient (uniform spacing, edge_order=1): interior (f[i+1]-f[i-1])/(2.*h), first/last (f[1]-f[0])/h.
// Every operation is rounded separately (no FMA contraction) and the divisions are correctly rounded, so the
// residual is bit-identical to numpy.  The loss is summed in a fixed order (deterministic); it differs from numpy's
// pairwise order by rounding only (<= 1e-12 relative; the contract is 1e-9).
//
// Design (HBM-bound: 8 B read + 8 B written per cell per chain, everything else must stay on chip):
//   * a CTA owns an (RS_RW*RS_WARPS) x 64 cell tile and loops over a group of chains; the five chain-independent
//     fields of the tile (+halo) are staged ONCE in shared memory and from there into each lane's registers, so per
//     chain only the bed is read from and the residual written to HBM;
//   * a warp owns RS_RW rows x 64 columns, a lane 2 adjacent columns: 16-byte accesses; x-neighbours come from warp
//     shuffles, y-neighbours from the lane's own registers (the warp fetches its own halo rows);
//   * the bed of the next chains streams in through a per-warp cp.async ring in shared memory (RS_STAGES-1 chains in
//     flight per warp): bytes in flight, not registers or warps, are what HBM latency has to be covered with;
//   * x/(2 res) uses the precomputed reciprocal with one FMA residual correction (correctly rounded, see div_const).
#include "common.cuh"
#include <cmath>
#include <cstdlib>

#ifndef RS_RW
#define RS_RW 2                       // rows per warp (even)
#endif
#ifndef RS_WARPS
#define RS_WARPS 4                    // warps per CTA
#endif
#define RS_TH (RS_RW * RS_WARPS)      // tile height 16
#define RS_TW 64                      // tile width (2 columns per lane)
#define RS_PITCH (RS_TW + 4)          // smem row pitch in doubles: [pad, haloL, 64 cells, haloR, pad] -> cells 16 B aligned
#define RS_THREADS (RS_WARPS * 32)
#ifndef RS_STAGES
#define RS_STAGES 4                   // cp.async ring depth per warp
#endif
#ifndef RS_MIN_CTAS
#define RS_MIN_CTAS 3                 // 3 CTAs/SM (<= 168 registers): the lane's statics live in registers, no spills
#endif

struct ResSmem {
    double surf[RS_TH + 2][RS_PITCH];
    double velx[RS_TH][RS_PITCH];
    double vely[RS_TH + 2][RS_PITCH];
    double dhdt[RS_TH][RS_PITCH];
    double smb[RS_TH][RS_PITCH];
    double ring[RS_STAGES][RS_WARPS][RS_RW + 2][RS_PITCH];   // per-warp landing zone of the streamed bed rows
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Loop-invariant geometry of one lane.
struct LaneGeom {
    int W, H, i0, lane, wr0, sc, hsc;
    unsigned rowmask;      // bit k: grid row i0-1+k exists
    bool v0, v1, hval;
    bool own;              // this lane's columns belong to this tile (false for the overlap of the shifted last tile)
    bool xl_edge, xr0_edge, xr1_edge;
    int k_top, k_bot;      // k of grid row 0 / H-1 inside this warp's rows, or -1
    int64_t hoff;          // offset of the halo column relative to the lane's first column
};

// MODE 1 (interior): the warp's rows and all 64 columns (+halo) lie strictly inside the grid: no predicates, no edge
// rules.  MODE 2 (column edge): rows strictly inside, all 64 columns valid, but the tile touches the first or last grid
// column (tiles are 64 wide everywhere: the last one is shifted left to end at W, see the kernel) - only lane 0 / lane
// 31 need np.gradient's one-sided rule.  MODE 0: everything else (first/last rows, odd or narrow grids), fully predicated.
// Every lane copies exactly the cells it will read back itself, so cp.async.wait_group alone orders the ring.
template <bool VEC, int MODE>
__device__ __forceinline__ void fetch_bed(const LaneGeom& g, const double* __restrict__ p, double (*slot)[RS_PITCH]) {
    constexpr bool INTERIOR = MODE == 1, ROWS_IN = MODE != 0;
    // p -> (row i0-1, column c0) of the chain's bed; slot -> this warp's [RS_RW+2][RS_PITCH] ring stage
#pragma unroll
    for (int k = 0; k < RS_RW + 2; ++k) {
        if (ROWS_IN || (((g.rowmask >> k) & 1u) && g.v0)) {
            const double* q = p + (int64_t)k * g.W;
            if (VEC) cp_async16(&slot[k][g.sc], q);
            else {
                cp_async8(&slot[k][g.sc], q);
                if (g.v1) cp_async8(&slot[k][g.sc + 1], q + 1);
            }
        }
    }
    if (INTERIOR ? (g.lane == 0 || g.lane == 31) : g.hval) {
#pragma unroll
        for (int k = 1; k <= RS_RW; ++k)
            if (ROWS_IN || ((g.rowmask >> k) & 1u)) cp_async8(&slot[k][g.hsc], p + (int64_t)k * g.W + g.hoff);
    }
}

// The lane's chain-independent operands, read once from the staged tile and then kept in registers for the whole chain
// loop: with them in shared memory the kernel was shared-memory-bandwidth bound (73 % of the LSU wavefront peak).
struct LaneStatics {
    double2 sf[RS_RW + 2], vy[RS_RW + 2];     // surf, vely of rows i0-1 .. i0+RS_RW, the lane's two columns
    double2 vx[RS_RW], dh[RS_RW], sm[RS_RW];  // velx, dhdt, smb of rows i0 .. i0+RS_RW-1
    double hsf[RS_RW], hvx[RS_RW];            // surf, velx of the lane's halo column (lanes 0 and 31)
    unsigned mcbits;                          // loss-mask bits: bit 2k (+1) = row k, first (second) column
};

__device__ __forceinline__ void load_statics(const ResSmem& S, const LaneGeom& g, LaneStatics& L) {
#pragma unroll
    for (int k = 0; k < RS_RW + 2; ++k) {
        L.sf[k] = *reinterpret_cast<const double2*>(&S.surf[g.wr0 + k][g.sc]);
        L.vy[k] = *reinterpret_cast<const double2*>(&S.vely[g.wr0 + k][g.sc]);
    }
#pragma unroll
    for (int k = 0; k < RS_RW; ++k) {
        L.vx[k] = *reinterpret_cast<const double2*>(&S.velx[g.wr0 + k][g.sc]);
        L.dh[k] = *reinterpret_cast<const double2*>(&S.dhdt[g.wr0 + k][g.sc]);
        L.sm[k] = *reinterpret_cast<const double2*>(&S.smb[g.wr0 + k][g.sc]);
        L.hsf[k] = S.surf[g.wr0 + k + 1][g.hsc];
        L.hvx[k] = S.velx[g.wr0 + k][g.hsc];
    }
}

// loss-mask bits of the lane's cells, straight from global memory (issued before the staging wait, used after it)
__device__ __forceinline__ unsigned load_mask_bits(const GmcDev& d, const LaneGeom& g, int c0) {
    unsigned bits = 0;
#pragma unroll
    for (int k = 0; k < RS_RW; ++k) {
        const int i = g.i0 + k;
        if (i < g.H && g.v0) {
            const uint8_t* f = d.flags + (int64_t)i * g.W + c0;
            if (__ldg(f) & FLAG_MC) bits |= 1u << (2 * k);
            if (g.v1 && (__ldg(f + 1) & FLAG_MC)) bits |= 1u << (2 * k + 1);
        }
    }
    return bits;
}

// true when |x| lies in [2^-930, 2^930]: the FMA-corrected quotient is then free of over/underflow (see div_const)
__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }

template <bool WRITE_RES, bool DO_LOSS, bool VEC, int MODE>
__device__ __forceinline__ void compute_rows(const GmcDev& d, const LaneStatics& L, const LaneGeom& g,
                                             const double (*bed)[RS_PITCH], double* __restrict__ out,
                                             double* __restrict__ partial, double r_res, double r_two_res) {
    constexpr bool INTERIOR = MODE == 1, ROWS_IN = MODE != 0;
    // ---- fluxes: fy for rows -1..RS_RW, fx for rows 0..RS_RW-1 --------------------------------------------------
    double fy0[RS_RW + 2], fy1[RS_RW + 2], fx0[RS_RW], fx1[RS_RW], fxh[RS_RW];
#pragma unroll
    for (int k = 0; k < RS_RW + 2; ++k) {
        const double2 sf = L.sf[k], vy = L.vy[k];
        const double2 bd = *reinterpret_cast<const double2*>(&bed[k][g.sc]);
        const double t0 = sub_rn(sf.x, bd.x), t1 = sub_rn(sf.y, bd.y);
        fy0[k] = mul_rn(vy.x, t0);
        fy1[k] = mul_rn(vy.y, t1);
        if (k >= 1 && k <= RS_RW) {
            const double2 vx = L.vx[k - 1];
            fx0[k - 1] = mul_rn(vx.x, t0);
            fx1[k - 1] = mul_rn(vx.y, t1);
            fxh[k - 1] = mul_rn(L.hvx[k - 1], sub_rn(L.hsf[k - 1], bed[k][g.hsc]));
        }
    }
    // np.gradient's one-sided first/last rows as data: duplicating the edge row into the missing neighbour turns the
    // central difference into (f[1]-f[0]) resp. (f[n-1]-f[n-2]); only the divisor changes (res instead of 2 res).
    if (!ROWS_IN) {
#pragma unroll
        for (int k = 0; k < RS_RW; ++k) {
            if (k == g.k_top) { fy0[k] = fy0[k + 1]; fy1[k] = fy1[k + 1]; }
            if (k == g.k_bot) { fy0[k + 2] = fy0[k + 1]; fy1[k + 2] = fy1[k + 1]; }
        }
    }
    // x-neighbours of every row first (straight-line code: the shuffles of all rows overlap)
    double fl[RS_RW], fr[RS_RW];
#pragma unroll
    for (int k = 0; k < RS_RW; ++k) {
        // left of column c0 is lane-1's second column, right of column c0+1 is lane+1's first
        fl[k] = __shfl_up_sync(0xffffffffu, fx1[k], 1);
        fr[k] = __shfl_down_sync(0xffffffffu, fx0[k], 1);
        if (g.lane == 0) fl[k] = fxh[k];
        if (g.lane == 31) fr[k] = fxh[k];
    }
    double acc = 0.0;
    // two rows at a time: 8 independent quotient chains in flight, one range check and one (cold) branch per pair
#pragma unroll
    for (int kk = 0; kk < RS_RW; kk += 2) {
        double num[2][4], den[2][4], rdn[2][4], quo[2][4];
        int lo = 0x7fffffff, hi = 0;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = kk + u;
            double a0 = fx1[k], l = fl[k], rr = fr[k];
            double deny = d.two_res, rdeny = r_two_res, den0 = d.two_res, rden0 = r_two_res, den1 = d.two_res, rden1 = r_two_res;
            if (!INTERIOR) {
                if (g.xl_edge) l = fx0[k];                        // column 0:   (f[1]   - f[0]  )/res
                if (g.xr1_edge) rr = fx1[k];                      // column W-1: (f[W-1] - f[W-2])/res (second column)
                if (!ROWS_IN && g.xr0_edge) a0 = fx0[k];          // column W-1 as first column (odd W)
                if (!ROWS_IN && ((k == g.k_top) || (k == g.k_bot))) { deny = d.res; rdeny = r_res; }
                if (g.xl_edge || (!ROWS_IN && g.xr0_edge)) { den0 = d.res; rden0 = r_res; }
                if (g.xr1_edge) { den1 = d.res; rden1 = r_res; }
            }
            num[u][0] = sub_rn(a0, l);
            num[u][1] = sub_rn(rr, fx0[k]);
            num[u][2] = sub_rn(fy0[k + 2], fy0[k]);
            num[u][3] = sub_rn(fy1[k + 2], fy1[k]);
            den[u][0] = den0; rdn[u][0] = rden0;
            den[u][1] = den1; rdn[u][1] = rden1;
            den[u][2] = den[u][3] = deny;
            rdn[u][2] = rdn[u][3] = rdeny;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // x / den by reciprocal + one FMA residual correction (see div_const)
                const double q = mul_rn(num[u][j], rdn[u][j]);
                // the correction never changes the sign; copysign also keeps -0 (0/den) exact
                quo[u][j] = copysign(fma(fma(-den[u][j], q, num[u][j]), rdn[u][j], q), q);
                int e = hi_abs(q);
                if ((e | __double2loint(q)) == 0) e = 0x3ff00000;            // exact zero: the fast path is exact
                if (!ROWS_IN && !(((g.rowmask >> (k + 1)) & 1u) && g.v0)) e = 0x3ff00000;   // padding lane/row: ignored
                lo = min(lo, e);
                hi = max(hi, e);
            }
        }
        if (lo < 0x05d00000 || hi > 0x7a100000) {                 // zero, subnormal, huge, inf or nan somewhere (cold)
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    quo[u][j] = div_slow(num[u][j], den[u][j]);
                }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = kk + u;
            const double2 dh = L.dh[k], sm = L.sm[k];
            const double r0 = sub_rn(add_rn(add_rn(quo[u][0], quo[u][2]), dh.x), sm.x);
            const double r1 = sub_rn(add_rn(add_rn(quo[u][1], quo[u][3]), dh.y), sm.y);
            if (ROWS_IN || (((g.rowmask >> (k + 1)) & 1u) && g.v0)) {
                if (WRITE_RES && (INTERIOR || g.own)) {
                    double* q = out + (int64_t)(k + 1) * g.W;
                    if (VEC) __stcs(reinterpret_cast<double2*>(q), make_double2(r0, r1));
                    else {
                        __stcs(q, r0);
                        if (ROWS_IN || g.v1) __stcs(q + 1, r1);
                    }
                }
                if (DO_LOSS) {                                   // mask bits of columns another tile owns are already 0
                    if (((L.mcbits >> (2 * k)) & 1u) && r0 == r0) acc = add_rn(acc, mul_rn(r0, r0));
                    if ((ROWS_IN || g.v1) && ((L.mcbits >> (2 * k + 1)) & 1u) && r1 == r1) acc = add_rn(acc, mul_rn(r1, r1));
                }
            }
        }
    }
    if (DO_LOSS) {
        acc = warp_sum(acc);
        if (g.lane == 0) *partial = acc;
    }
}

template <bool WRITE_RES, bool DO_LOSS, bool VEC, int MODE>
__device__ __forceinline__ void chain_loop(const GmcDev& d, ResSmem& S, const LaneGeom& g, int warp, const double* __restrict__ pb,
                                           double* __restrict__ po, double* __restrict__ pp, int64_t plane, int n_tiles, int C,
                                           double r_res, double r_two_res, unsigned mcbits) {
    const int G = gridDim.z;
    const int64_t cstride = (int64_t)G * plane, pstride = (int64_t)G * n_tiles;
    const int n_iter = (C - (int)blockIdx.z + G - 1) / G;          // chains blockIdx.z, +G, +2G, ...
    // prologue: RS_STAGES-1 chains in flight (empty groups keep the group count uniform), issued behind the statics'
    // copy group so that both latencies overlap
#pragma unroll
    for (int s = 0; s < RS_STAGES - 1; ++s) {
        if (s < n_iter) fetch_bed<VEC, MODE>(g, pb + s * cstride, S.ring[s][warp]);
        cp_async_commit();
    }
    cp_async_wait<RS_STAGES - 1>();                                // this thread's share of the statics has landed
    __syncthreads();                                               // ... and everybody else's
    LaneStatics L;
    load_statics(S, g, L);
    L.mcbits = mcbits;
    int stage = 0;
    for (int it = 0; it < n_iter; ++it) {
        const int nxt = it + RS_STAGES - 1;
        int ns = stage + RS_STAGES - 1;
        if (ns >= RS_STAGES) ns -= RS_STAGES;
        if (nxt < n_iter) fetch_bed<VEC, MODE>(g, pb + (int64_t)nxt * cstride, S.ring[ns][warp]);
        cp_async_commit();
        cp_async_wait<RS_STAGES - 1>();                            // this lane's copies of chain `it` have landed
        compute_rows<WRITE_RES, DO_LOSS, VEC, MODE>(d, L, g, S.ring[stage][warp], po, pp, r_res, r_two_res);
        if (WRITE_RES) po += cstride;
        pp += pstride;
        if (++stage == RS_STAGES) stage = 0;
    }
    cp_async_wait<0>();
}

template <bool WRITE_RES, bool DO_LOSS, bool VEC>
__global__ void __launch_bounds__(RS_THREADS, RS_MIN_CTAS)
    residual_kernel(GmcDev d, const double* __restrict__ bed_all, double* __restrict__ res_all,
                    double* __restrict__ partials, int n_tiles, int C, double r_res, double r_two_res) {
    extern __shared__ __align__(16) unsigned char rs_raw[];
    ResSmem& S = *reinterpret_cast<ResSmem*>(rs_raw);
    const int H = d.H, W = d.W;
    const int tid = threadIdx.x, warp = tid >> 5;
    // read once through a volatile asm: otherwise the compiler re-reads SR_TID.X inside the chain loop to save a register,
    // and the S2R latency (~25 cycles, twice per trip) lands on the critical path
    int lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    // Tiles are RS_TW wide; with 16-byte accesses (even W >= RS_TW) the last tile of a row is shifted left to end exactly
    // at W, so every lane of every tile holds valid columns.  The columns it shares with its left neighbour are computed
    // twice (same bits) but stored and summed once (`own`).
    const bool shift = VEC && W >= RS_TW;
    const int tx_nom = blockIdx.x * RS_TW;
    const int tx0 = shift ? min(tx_nom, W - RS_TW) : tx_nom;
    const int ty0 = blockIdx.y * RS_TH;
    const int64_t plane = (int64_t)H * W;

    // ---- stage the chain-independent fields of this tile (+ one-cell halo): asynchronous copies, all in flight at once
    // (one exposed memory latency per CTA instead of one per loop trip); cells outside the grid read as 0.
    const bool cta_interior = tx0 > 0 && tx0 + RS_TW < W && ty0 > 0 && ty0 + RS_TH < H;
    for (int t = tid; t < (RS_TH + 2) * (RS_TW + 2); t += RS_THREADS) {
        const int rr = t / (RS_TW + 2), cc = t - rr * (RS_TW + 2);      // rr in [0,18): row ty0-1+rr; cc in [0,66): col tx0-1+cc
        const int i = ty0 - 1 + rr, j = tx0 - 1 + cc;
        const bool in = cta_interior || (i >= 0 && i < H && j >= 0 && j < W);
        const int64_t idx = (int64_t)i * W + j;
        const bool mid = rr >= 1 && rr <= RS_TH;
        if (in) {
            cp_async8(&S.surf[rr][cc + 1], d.surf + idx);
            cp_async8(&S.vely[rr][cc + 1], d.vely + idx);
            if (mid) {
                cp_async8(&S.velx[rr - 1][cc + 1], d.velx + idx);
                cp_async8(&S.dhdt[rr - 1][cc + 1], d.dhdt + idx);
                cp_async8(&S.smb[rr - 1][cc + 1], d.smb + idx);
            }
        } else {
            S.surf[rr][cc + 1] = 0.0;
            S.vely[rr][cc + 1] = 0.0;
            if (mid) S.velx[rr - 1][cc + 1] = S.dhdt[rr - 1][cc + 1] = S.smb[rr - 1][cc + 1] = 0.0;
        }
    }
    cp_async_commit();
    if (!cta_interior) {
        // boundary tiles: padding lanes / rows of the ring are read but never copied; give them a defined value.  Must
        // precede the first ring copy.
        for (int t = tid; t < RS_STAGES * RS_WARPS * (RS_RW + 2) * RS_PITCH; t += RS_THREADS) (&S.ring[0][0][0][0])[t] = 0.0;
        __syncthreads();
    }

    LaneGeom g;
    g.W = W;
    g.H = H;
    g.lane = lane;
    g.wr0 = warp * RS_RW;                         // first tile row of this warp
    g.i0 = ty0 + g.wr0;                           // first grid row
    const bool warp_on = g.i0 < H && (int)blockIdx.z < C;      // warp-uniform; idle warps still join the staging barrier
    const int c0 = tx0 + 2 * lane;                // first grid column of this lane
    g.sc = 2 * lane + 2;                          // smem column of c0 (16 B aligned)
    g.v0 = c0 < W;
    g.v1 = c0 + 1 < W;
    g.own = c0 >= tx_nom;
    const int hcol = (lane == 0) ? c0 - 1 : c0 + 2;          // lane 0: left of the tile, lane 31: right of it
    g.hval = (lane == 0 && hcol >= 0) || (lane == 31 && hcol < W);
    g.hsc = (lane == 0) ? g.sc - 1 : g.sc + 2;
    g.hoff = hcol - c0;
    g.xl_edge = (c0 == 0);
    g.xr0_edge = (c0 == W - 1);
    g.xr1_edge = (c0 + 1 == W - 1);
    g.rowmask = 0;
    for (int k = 0; k < RS_RW + 2; ++k)
        if (g.i0 - 1 + k >= 0 && g.i0 - 1 + k < H) g.rowmask |= 1u << k;
    g.k_top = (g.i0 == 0) ? 0 : -1;
    g.k_bot = (H - 1 >= g.i0 && H - 1 < g.i0 + RS_RW) ? H - 1 - g.i0 : -1;
    const int tile_id = (blockIdx.y * RS_WARPS + warp) * gridDim.x + blockIdx.x;

    // chains c = blockIdx.z, +G, +2G, ...: two register sets ping-pong so the next chain's bed is always in flight
    const int64_t lane_off = (int64_t)(g.i0 - 1) * W + c0;
    const double* pb = bed_all + (int64_t)blockIdx.z * plane + lane_off;
    double* po = WRITE_RES ? res_all + (int64_t)blockIdx.z * plane + lane_off : nullptr;
    double* pp = partials + (int64_t)blockIdx.z * n_tiles + tile_id;
    const bool interior = tx0 > 0 && tx0 + RS_TW < W && g.i0 > 0 && g.i0 + RS_RW < H;   // warp-uniform
    const unsigned mcbits = (DO_LOSS && warp_on && g.own) ? load_mask_bits(d, g, c0) : 0u;
    if (!warp_on) {
        cp_async_wait<0>();
        __syncthreads();
        return;
    }
    const bool rows_in = g.i0 > 0 && g.i0 + RS_RW < H;                                  // warp-uniform
    if (interior) chain_loop<WRITE_RES, DO_LOSS, VEC, 1>(d, S, g, warp, pb, po, pp, plane, n_tiles, C, r_res, r_two_res, mcbits);
    else if (shift && rows_in) chain_loop<WRITE_RES, DO_LOSS, VEC, 2>(d, S, g, warp, pb, po, pp, plane, n_tiles, C, r_res, r_two_res, mcbits);
    else chain_loop<WRITE_RES, DO_LOSS, VEC, 0>(d, S, g, warp, pb, po, pp, plane, n_tiles, C, r_res, r_two_res, mcbits);
}

// masked nansum of squares of given residuals: one CTA per (chunk, chain)
#define LOSS_THREADS 256
__global__ void __launch_bounds__(LOSS_THREADS)
    loss_kernel(GmcDev d, const double* __restrict__ res_all, double* __restrict__ partials, int n_tiles) {
    __shared__ double scratch[33];
    const int c = blockIdx.y;
    const int64_t plane = (int64_t)d.H * d.W;
    const double* res = res_all + c * plane;
    const int64_t chunk = (plane + n_tiles - 1) / n_tiles;
    const int64_t lo = blockIdx.x * chunk;
    const int64_t hi = (lo + chunk < plane) ? lo + chunk : plane;
    double acc = 0.0;
    for (int64_t k = lo + threadIdx.x; k < hi; k += LOSS_THREADS) {
        const double v = __ldcs(res + k);
        if ((__ldg(d.flags + k) & FLAG_MC) && v == v) acc = add_rn(acc, mul_rn(v, v));
    }
    const double t = block_sum<LOSS_THREADS>(acc, scratch);
    if (threadIdx.x == 0) partials[(int64_t)c * n_tiles + blockIdx.x] = t;
}

// fixed-order sum of the per-tile partials: one CTA per chain (thread-strided sums, then the block tree).  One warp per chain
// took 45 us for 256 x 2016 partials - a quarter of the fused residual+loss time at 500x500 - because 32 CTAs of dependent
// adds cannot hide the load latency.
#define FIN_THREADS 256
__global__ void __launch_bounds__(FIN_THREADS)
    finalize_loss_kernel(const double* __restrict__ partials, int n_tiles, double two_sigma2, double* __restrict__ loss_out,
                         double* __restrict__ ssq_out, int C) {
    __shared__ double scratch[33];
    const int c = blockIdx.x;
    const double* p = partials + (int64_t)c * n_tiles;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int k = threadIdx.x;
    for (; k + 3 * FIN_THREADS < n_tiles; k += 4 * FIN_THREADS) {      // four loads in flight per thread
        a0 += p[k];
        a1 += p[k + FIN_THREADS];
        a2 += p[k + 2 * FIN_THREADS];
        a3 += p[k + 3 * FIN_THREADS];
    }
    for (; k < n_tiles; k += FIN_THREADS) a0 += p[k];
    const double acc = block_sum<FIN_THREADS>((a0 + a1) + (a2 + a3), scratch);
    if (threadIdx.x == 0) {
        if (ssq_out) ssq_out[c] = acc;
        if (loss_out) loss_out[c] = div_rn(acc, two_sigma2);
    }
}

// debug/test hook: div_const vs the true division on caller-provided dividends
__global__ void div_check_kernel(const double* __restrict__ x, int64_t n, double dd, double r, unsigned long long* mismatches) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double a = div_const(x[k], dd, r), b = div_rn(x[k], dd);
    if (__double_as_longlong(a) != __double_as_longlong(b) && !(a != a && b != b)) atomicAdd(mismatches, 1ull);
}

#include "stencil_tma.cuh"
static_assert(R2_TH == RS_TH && R2_WARPS == RS_WARPS && R2_TW == RS_TW, "both stencil kernels share the partial-sum tiling");

static int tiles_x(const gmc_ctx* c) { return (c->W + RS_TW - 1) / RS_TW; }
static int tiles_y(const gmc_ctx* c) { return (c->H + RS_TH - 1) / RS_TH; }

static int ensure_partials(gmc_ctx* c) {
    const int tiles = tiles_x(c) * tiles_y(c) * RS_WARPS;
    if (!c->d_partials || c->n_tiles != tiles) {
        cudaFree(c->d_partials);
        c->d_partials = nullptr;
        GMC_CUDA(cudaMalloc(&c->d_partials, (size_t)c->max_chains * tiles * sizeof(double)));
        GMC_CUDA(cudaMemset(c->d_partials, 0, (size_t)c->max_chains * tiles * sizeof(double)));
        c->n_tiles = tiles;
    }
    return GMC_OK;
}

static int check_common(gmc_ctx* c, const void* p, int C, const char* who) {
    if (!c) GMC_FAIL(GMC_EINVAL, "%s: ctx is NULL", who);
    if (!c->have_static) GMC_FAIL(GMC_ESTATE, "%s: call gmc_set_static first", who);
    if (!p) GMC_FAIL(GMC_EINVAL, "%s: NULL array pointer", who);
    if (C < 1 || C > c->max_chains) GMC_FAIL(GMC_ESHAPE, "%s: C=%d outside [1,%d]", who, C, c->max_chains);
    GMC_CUDA(cudaSetDevice(c->device));
    return GMC_OK;
}

// TMA path (even W, 16-byte aligned bases): tensor maps over this call's bed / residual arrays, chain groups sized for the
// kernel's own occupancy.  Returns false when it does not apply (the caller then launches residual_kernel).
template <bool WR, bool LS, bool TST>
static bool launch_tma_variant(gmc_ctx* c, cudaStream_t st, const double* bed, double* res, int C, double* partials) {
    CUtensorMap tm_bed, tm_out;
    if (!r2_encode(&tm_bed, bed, C, c->H, c->W, R2_BOXW, R2_FETCH_ROWS)) return false;
    if (!r2_encode(&tm_out, WR ? res : bed, C, c->H, c->W, R2_TW, R2_RW)) return false;
    const int smem = r2_layout(WR && TST, LS, r2_stages(WR)).total;
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) {
        if (cudaFuncSetAttribute(residual_tma_kernel<WR, LS, TST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return false;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, residual_tma_kernel<WR, LS, TST>, R2_THREADS, smem);
        ctas_per_sm = nb > 0 ? nb : 1;
    }
    const int tx = (c->W + R2_TW - 1) / R2_TW, ty = (c->H + R2_TH - 1) / R2_TH;
    const int slots = ctas_per_sm * c->sm_count;
    int groups = 1;
    double best = -1.0;
    // at most ~48 chains per CTA: CTAs of vertically adjacent tiles share two halo rows through L2 only while they work on
    // the same chains at about the same time; over hundreds of chains they drift apart and the halo rows come from DRAM
    // again (measured: 4096 chains in 7 groups 72 % of the HBM peak vs 82 % for 256 chains in 7 groups)
    // The loss-only variant (WR = false) forms its 20 coefficients per lane from ~44 scalar loads: its prologue costs about four
    // chain-iterations, and longer CTAs pay (sweeps in profiles/r2/stencil_lin_ab.txt: 4096 chains 81 % of the HBM peak with 43
    // groups of 95 chains, 78 % with 86 groups of 48), so its cap is 96 chains per CTA.
    const int cap = WR ? 48 : 96;
    const double prologue = WR ? 1.0 : 4.0;
    const int g_min = std::max(1, (C + cap - 1) / cap);
    for (int G = g_min; G <= std::max(g_min, std::min(C / 4, 4096)); ++G) {
        // cost model fitted to the sweeps in profiles/r2/stencil_groups.txt: a CTA's prologue costs about one chain-iteration
        // (four for the loss-only variant), and the tail of the grid about one CTA duration (CTAs differ: boundary tiles,
        // the producer warp), i.e. 1 / waves
        const double n = (double)C / G, waves = (double)tx * ty * G / slots;
        const double score = n / (n + prologue) * waves / (waves + 1.0);
        if (score > best + 1e-9) {
            best = score;
            groups = G;
        }
    }
    if (const char* e = getenv("GMC_RS_GROUPS")) groups = std::max(1, std::min(atoi(e), C));
    groups = std::min(groups, 65535);
    residual_tma_kernel<WR, LS, TST><<<dim3(tx, ty, groups), R2_THREADS, smem, st>>>(tm_bed, tm_out, c->dev, res, partials, c->n_tiles, C,
                                                                                c->dev.r_res, c->dev.r_two_res);
    return true;
}

template <bool WR, bool LS>
static void launch_variant(gmc_ctx* c, dim3 grid, cudaStream_t st, const double* bed, double* res, int C, bool vec, double* partials) {
    const double r_res = c->dev.r_res, r_two = c->dev.r_two_res;
    const size_t smem = sizeof(ResSmem);
    if (vec) {
        cudaFuncSetAttribute(residual_kernel<WR, LS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        residual_kernel<WR, LS, true><<<grid, RS_THREADS, smem, st>>>(c->dev, bed, res, partials, c->n_tiles, C, r_res, r_two);
    } else {
        cudaFuncSetAttribute(residual_kernel<WR, LS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        residual_kernel<WR, LS, false><<<grid, RS_THREADS, smem, st>>>(c->dev, bed, res, partials, c->n_tiles, C, r_res, r_two);
    }
}

static int launch_residual(gmc_ctx* c, const double* bed, double* res_out, double* loss_out, double* ssq_out, int C,
                           bool do_loss, cudaStream_t st, int chain0 = 0) {
    int rc = ensure_partials(c);
    if (rc) return rc;
    // chain0 selects this call's rows of the partial-sum workspace, so calls for disjoint chain ranges may overlap on
    // different streams
    double* partials = c->d_partials + (size_t)chain0 * c->n_tiles;
    const int tx = tiles_x(c), ty = tiles_y(c);
    // chain groups G (gridDim.z): a CTA keeps its tile's statics in registers and walks C/G chains.  Few groups amortise
    // the tile prologue (~4 chain-iterations' worth) over many chains; the CTA count tiles*G should fill whole waves of
    // RS_MIN_CTAS CTAs per SM.  Pick the G that maximises (useful fraction of a CTA) x (wave efficiency).
    const int slots = RS_MIN_CTAS * c->sm_count;
    int groups = 1;
    double best = -1.0;
    for (int G = 1; G <= std::max(1, std::min(C / 4, 256)); ++G) {
        const double n = (double)C / G, waves = (double)tx * ty * G / slots;
        const double score = n / (n + 4.0) * (waves / std::ceil(waves)) * (waves < 1.0 ? waves : 1.0);
        if (score > best + 1e-9) {
            best = score;
            groups = G;
        }
    }
    if (const char* e = getenv("GMC_RS_GROUPS")) groups = std::max(1, std::min(atoi(e), C));   // tuning override
    groups = std::min(groups, 65535);
    const dim3 grid(tx, ty, groups);
    // 16-byte accesses need even W and 16 B aligned bases
    const bool vec = (c->W % 2 == 0) && ((uintptr_t)bed % 16 == 0) && (!res_out || (uintptr_t)res_out % 16 == 0);
    static const bool legacy = getenv("GMC_RS_LEGACY") != nullptr;      // A/B switch: the cp.async kernel of round 1
    if (vec && !legacy) {
        // residual tile out through TMA tensor stores (GMC_RS_TMA_STORE=1) or 16-byte streaming stores (default: measured
        // 81.7 % vs 75.1 % of the HBM peak at 256 x 500^2, 90.6 % vs 91.3 % at 128 x 2000^2 - profiles/README.md)
        static const bool tma_store = getenv("GMC_RS_TMA_STORE") && getenv("GMC_RS_TMA_STORE")[0] == '1';
        bool ok;
        if (res_out && do_loss) ok = tma_store ? launch_tma_variant<true, true, true>(c, st, bed, res_out, C, partials)
                                               : launch_tma_variant<true, true, false>(c, st, bed, res_out, C, partials);
        else if (res_out) ok = tma_store ? launch_tma_variant<true, false, true>(c, st, bed, res_out, C, partials)
                                         : launch_tma_variant<true, false, false>(c, st, bed, res_out, C, partials);
        else ok = launch_tma_variant<false, true, false>(c, st, bed, res_out, C, partials);
        if (ok) {
            c->launches++;
            if (do_loss) {
                finalize_loss_kernel<<<C, FIN_THREADS, 0, st>>>(partials, c->n_tiles, c->dev.two_sigma2, loss_out, ssq_out, C);
                c->launches++;
            }
            GMC_CUDA(cudaGetLastError());
            return GMC_OK;
        }
    }
    if (res_out && do_loss) launch_variant<true, true>(c, grid, st, bed, res_out, C, vec, partials);
    else if (res_out) launch_variant<true, false>(c, grid, st, bed, res_out, C, vec, partials);
    else launch_variant<false, true>(c, grid, st, bed, res_out, C, vec, partials);
    c->launches++;
    if (do_loss) {
        finalize_loss_kernel<<<C, FIN_THREADS, 0, st>>>(partials, c->n_tiles, c->dev.two_sigma2, loss_out, ssq_out, C);
        c->launches++;
    }
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_residual(gmc_ctx* c, const double* bed, double* res_out, int C, void* stream) {
    int rc = check_common(c, bed, C, "gmc_residual");
    if (rc) return rc;
    if (!res_out) GMC_FAIL(GMC_EINVAL, "gmc_residual: res_out is NULL");
    return launch_residual(c, bed, res_out, nullptr, nullptr, C, false, (cudaStream_t)stream);
}

extern "C" int gmc_residual_loss(gmc_ctx* c, const double* bed, double* res_out, double* loss_out, double* ssq_out,
                                 int C, void* stream) {
    int rc = check_common(c, bed, C, "gmc_residual_loss");
    if (rc) return rc;
    if (!loss_out && !ssq_out) GMC_FAIL(GMC_EINVAL, "gmc_residual_loss: loss_out and ssq_out are both NULL");
    return launch_residual(c, bed, res_out, loss_out, ssq_out, C, true, (cudaStream_t)stream);
}

extern "C" int gmc_residual_loss_range(gmc_ctx* c, const double* bed, double* res_out, double* loss_out, double* ssq_out,
                                       int C, int chain0, void* stream) {
    int rc = check_common(c, bed, C, "gmc_residual_loss_range");
    if (rc) return rc;
    if (!loss_out && !ssq_out) GMC_FAIL(GMC_EINVAL, "gmc_residual_loss_range: loss_out and ssq_out are both NULL");
    if (chain0 < 0 || chain0 + C > c->max_chains)
        GMC_FAIL(GMC_ESHAPE, "gmc_residual_loss_range: chains [%d,%d) outside the context's capacity %d", chain0, chain0 + C, c->max_chains);
    return launch_residual(c, bed, res_out, loss_out, ssq_out, C, true, (cudaStream_t)stream, chain0);
}

extern "C" int gmc_loss(gmc_ctx* c, const double* res, double* loss_out, double* ssq_out, int C, void* stream) {
    int rc = check_common(c, res, C, "gmc_loss");
    if (rc) return rc;
    if (!loss_out && !ssq_out) GMC_FAIL(GMC_EINVAL, "gmc_loss: loss_out and ssq_out are both NULL");
    rc = ensure_partials(c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // chunks of >= 16 K cells per CTA (each thread streams >= 64 cells), at most as many as there are partial slots
    const int64_t plane = (int64_t)c->H * c->W;
    const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(c->n_tiles, plane / 16384));
    for (int c0 = 0; c0 < C; c0 += 65535) {
        const int cn = (C - c0 < 65535) ? C - c0 : 65535;
        loss_kernel<<<dim3(chunks, cn), LOSS_THREADS, 0, st>>>(c->dev, res + (size_t)c0 * c->H * c->W,
                                                              c->d_partials + (size_t)c0 * chunks, chunks);
        c->launches++;
    }
    finalize_loss_kernel<<<C, FIN_THREADS, 0, st>>>(c->d_partials, chunks, c->dev.two_sigma2, loss_out, ssq_out, C);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_debug_div_check(gmc_ctx* c, const double* x, int64_t n, double divisor, int64_t* mismatches_out) {
    if (!c || !x || !mismatches_out || n < 1) GMC_FAIL(GMC_EINVAL, "gmc_debug_div_check: bad argument");
    GMC_CUDA(cudaSetDevice(c->device));
    unsigned long long* dm = nullptr;
    GMC_CUDA(cudaMalloc(&dm, sizeof(unsigned long long)));
    GMC_CUDA(cudaMemset(dm, 0, sizeof(unsigned long long)));
    div_check_kernel<<<(unsigned)((n + 255) / 256), 256>>>(x, n, divisor, 1.0 / divisor, dm);
    unsigned long long h = 0;
    GMC_CUDA(cudaMemcpy(&h, dm, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(dm);
    *mismatches_out = (int64_t)h;
    return GMC_OK;
}
