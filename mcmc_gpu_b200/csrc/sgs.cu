// sgs.cu — K6: the small-scale chain step (block re-simulation by Sequential Gaussian Simulation).
//
// Reference: chain_sgs.run loop body (MCMC.py:1747-1829), MCMC.sgs (MCMC.py:91-173), octant neighbour search
// (gstatsim_custom/neighbors.py:4-64), ordinary kriging (gstatsim_custom/_krige.py:5-44), covariance models
// (gstatsim_custom/covariance.py), sklearn QuantileTransformer column transform (MCMC.py:1767,1777).
//
// One CTA (256 threads = 8 warps = the 8 octants) owns one chain.  Per iteration:
//   1. the block's cells are reset to the normal-scored radar values (NaN elsewhere) in shared memory;
//   2. along the (injected or Philox-shuffled) path every unconditioned node is kriged: warp o walks octant o's offset
//      list — offsets pre-sorted on the host by (distance, row-major position), so the first num_points/8 conditioned
//      hits ARE the reference's argsort selection — the (n+1)x(n+1) ordinary-kriging system is assembled from a host-built
//      covariance table over integer offsets (this is where scipy's Bessel-K Matern lives) and solved in shared memory
//      (elimination without pivoting on the SPD block, two right-hand sides, Lagrange multiplier by Schur complement);
//   3. the block is mapped back through the normal-score tables, the residual is recomputed on the block plus its
//      one-cell ring (the reference recomputes the full grid: the result is identical, nothing else changed), the loss
//      changes by that region's delta, the full-grid thickness guard is a running violation count, Metropolis decides.
// Where the reference calls np.linalg.lstsq (SVD, minimum norm) we solve exactly; the systems are well conditioned
// (cond ~1e4, |d est| <= 6e-13), parity is <= 1e-9 as for every floating-point result of this chain.
#include "common.cuh"
#include <cstdlib>

#define SGS_THREADS 256
#define SGS_MAX_BLOCK 32          // largest block edge (cells)
#define SGS_MAX_NEIGH 64          // largest num_points
#define SGS_SIG_PITCH (SGS_MAX_NEIGH + 3)   // augmented matrix [Sigma | rho | 1], odd pitch
#define SGS_NEAR 64               // offsets per octant staged in shared memory

#define SGS_MAX_LEVELS 4           // search radii radius, radius + 100 km, ... (MCMC.py:149-155, interpolate.py:149-155)

struct SgsDev {
    const double* trend;          // [H][W] or nullptr
    const double* zcond;          // [H][W] normal-scored radar values, NaN where none
    const uint8_t* grounded;      // [H][W]
    const double* quant;          // [nq] quantiles (nullptr: no transform)
    const double* refs;           // [nq] references
    int nq;
    const int16_t* oct_off;       // [8][lmax][2] (di, dj), sorted by (distance, di, dj); octant b-(-4) of neighbors.py:54
    const int32_t* oct_cnt;       // [n_levels][8]: prefix lengths of the lists for radius, radius + 100 km, ... (MCMC.py:149-155)
    int n_levels;
    int lmax, hw, per_oct;        // per_oct = num_points // 8
    const double* lut;            // [(4hw+1)][(4hw+1)] covariance of the offset (di, dj)
    int lut_w;                    // 4hw+1
    double sill;
    int bmin_x, bmax_x, bmin_y, bmax_y;
    long long* phase;             // optional cycle counters (debug), see gmc_debug_phase_timing
};

struct gmc_sgs_state {
    SgsDev dev;
    bool ready;
    bool warp_solver;
    void* owned[8];
};

// ---- normal-score transform (QuantileTransformer._transform_col, output_distribution='normal') ---------------------
// np.interp(x, xp, fp): binary search + slope*(x - xp[j]) + fp[j]; clamped outside [xp[0], xp[n-1]]
__device__ __forceinline__ double interp_tab(double x, const double* __restrict__ xp, const double* __restrict__ fp, int n,
                                             bool negrev) {
    // negrev: use the tables -xp[::-1], -fp[::-1] (second interpolation of the forward transform)
    auto XP = [&](int k) { return negrev ? -xp[n - 1 - k] : xp[k]; };
    auto FP = [&](int k) { return negrev ? -fp[n - 1 - k] : fp[k]; };
    if (x <= XP(0)) return FP(0);
    if (x >= XP(n - 1)) return FP(n - 1);
    int lo = 0, hi = n - 1;                       // XP(lo) <= x < XP(hi)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (XP(mid) <= x) lo = mid;
        else hi = mid;
    }
    const double slope = div_rn(sub_rn(FP(lo + 1), FP(lo)), sub_rn(XP(lo + 1), XP(lo)));
    return add_rn(mul_rn(slope, sub_rn(x, XP(lo))), FP(lo));
}

__device__ __forceinline__ double nst_forward(const SgsDev& s, double x) {
    if (!s.quant) return x;
    if (x != x) return x;
    const double lo_q = s.quant[0], hi_q = s.quant[s.nq - 1];
    double y = 0.5 * (interp_tab(x, s.quant, s.refs, s.nq, false) - interp_tab(-x, s.quant, s.refs, s.nq, true));
    if (x + 1e-7 > hi_q) y = 1.0;
    if (x - 1e-7 < lo_q) y = 0.0;
    double z = normcdfinv(y);
    // clip_min = norm.ppf(1e-7 - spacing(1)), clip_max = norm.ppf(1 - (1e-7 - spacing(1)))  (not symmetric in floating point)
    return fmin(fmax(z, -5.199337582605575), 5.19933758270342);
}

__device__ __forceinline__ double nst_inverse(const SgsDev& s, double z) {
    if (!s.quant) return z;
    if (z != z) return z;
    const double p = normcdf(z);
    double x = interp_tab(p, s.refs, s.quant, s.nq, false);
    if (p + 1e-7 > 1.0) x = s.quant[s.nq - 1];
    if (p - 1e-7 < 0.0) x = s.quant[0];
    return x;
}

__global__ void sgs_transform_kernel(SgsDev s, const double* __restrict__ in, double* __restrict__ out, int64_t n, int inverse) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = inverse ? nst_inverse(s, in[k]) : nst_forward(s, in[k]);
}

// bedc = bed - trend;  z = forward(bedc);  nviol = #{grounded & surf - bed <= 0}
__global__ void sgs_init_kernel(GmcDev d, SgsDev s, const double* __restrict__ bed, double* __restrict__ bedc,
                                double* __restrict__ z, int32_t* __restrict__ nviol, int C) {
    const int c = blockIdx.y;
    const int64_t plane = (int64_t)d.H * d.W;
    int local = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < plane; k += (int64_t)gridDim.x * blockDim.x) {
        const double b = bed[c * plane + k];
        const double bc = s.trend ? sub_rn(b, s.trend[k]) : b;
        bedc[c * plane + k] = bc;
        z[c * plane + k] = nst_forward(s, bc);
        const double full = s.trend ? add_rn(bc, s.trend[k]) : bc;
        if (s.grounded[k] == 1 && sub_rn(d.surf[k], full) <= 0.0) ++local;
    }
    if (local) atomicAdd(nviol + c, local);
}

// full bed = bedc + trend  (for the residual of the initial state)
__global__ void sgs_fullbed_kernel(GmcDev d, SgsDev s, const double* __restrict__ bedc, double* __restrict__ full, int64_t n) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) full[k] = s.trend ? add_rn(bedc[k], s.trend[k % ((int64_t)d.H * d.W)]) : bedc[k];
}

// ---- the step -------------------------------------------------------------------------------------------------------
struct SgsShared {
    double blk_z[SGS_MAX_BLOCK * SGS_MAX_BLOCK];       // simulated normal scores of the block (NaN: not yet)
    double cand[SGS_MAX_BLOCK * SGS_MAX_BLOCK];        // candidate detrended bed of the block
    double newres[(SGS_MAX_BLOCK + 2) * (SGS_MAX_BLOCK + 2)];
    double sig[(SGS_MAX_NEIGH) * SGS_SIG_PITCH];       // Sigma (row-major, pitch n+... fixed)
    double rhs_a[SGS_MAX_NEIGH], rhs_b[SGS_MAX_NEIGH]; // right-hand sides rho and 1
    double nval[SGS_MAX_NEIGH];
    int16_t ndi[SGS_MAX_NEIGH], ndj[SGS_MAX_NEIGH];
    int oct_n[8];
    int path[SGS_MAX_BLOCK * SGS_MAX_BLOCK];
    short2 near_off[8][SGS_NEAR];                      // the nearest offsets of every octant (almost every search ends here)
    int16_t ord[SGS_MAX_BLOCK * SGS_MAX_BLOCK];        // warp solver: -1 = radar-conditioned, else position among the nodes to simulate
    int16_t todo[SGS_MAX_BLOCK * SGS_MAX_BLOCK];       // warp solver: the nodes to simulate, in path order
    int16_t todo_k[SGS_MAX_BLOCK * SGS_MAX_BLOCK];     // ... and their positions in the full path (index of the injected normals)
    int n_todo;
    double scratch[40];
    int ix, iy, bsx, bsy, x0, x1, y0, y1, accept, n_nb, err;
    double u;
};

__device__ __forceinline__ double lut_cov(const SgsDev& s, int di, int dj) {
    return __ldg(s.lut + (di + 2 * s.hw) * s.lut_w + (dj + 2 * s.hw));
}

// ---- warp-per-node kriging ------------------------------------------------------------------------------------------
// Which cells are conditioned when a node's turn comes is known from the path alone (radar cells, cells outside the block,
// and block cells earlier in the path); the simulated VALUES enter only through est = mean + sum_i w_i (v_i - mean).  So the
// neighbour search and the kriging solve of every node are independent of each other: eight warps work on eight
// consecutive path nodes at once (phase 1), then one warp turns the eight (weights, variance) into values in path order
// (phase 2, a 48-term dot product each).  A warp holds its augmented matrix [Sigma | rho | 1] (48 x 50) in REGISTERS,
// lane (lr, lc) of a 4 x 8 layout owning rows lr+4a (a<12) and columns lc+8b (b<7) - cyclic, so the shrinking active part
// stays balanced; per pivot only the pivot row and column pass through a per-warp shared-memory buffer (__syncwarp, no
// CTA barrier).  Gauss-Jordan without pivoting on the SPD block, as in the CTA-wide solver it replaces for num_points <= 48.
#define SGS_WN 48
struct SgsWarpRec {
    double nval[SGS_WN], w[SGS_WN], rho[SGS_WN];
    double xrow[2][56], xcol[2][SGS_WN];
    double dg[SGS_WN], ra[SGS_WN], rb[SGS_WN];
    int16_t ndi[SGS_WN], ndj[SGS_WN], nsrc[SGS_WN];
    int n, node, kpath;
    double var, zn;
};
static_assert(8 * sizeof(SgsWarpRec) <= SGS_MAX_NEIGH * SGS_SIG_PITCH * sizeof(double), "warp records must fit in SgsShared::sig");

// Kriging weights and variance of one node from its n <= 48 neighbour offsets R.ndi/ndj (n > 0): assemble, eliminate,
// Lagrange multiplier.  Results: R.w[0..n), R.var.  One warp; see the layout notes above.
__device__ __forceinline__ void sgs_warp_solve(const SgsDev& s, SgsWarpRec& R, int n) {
    const int lane = threadIdx.x & 31;
    // (b) assemble [Sigma | rho | 1] straight into registers                                       _krige.py:20-33
    const int lr = lane & 3, lc = lane >> 2;
    double A[12][7];
    {
        int rdi[12], rdj[12], cdi[6], cdj[6];
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            const int r = lr + 4 * a;
            rdi[a] = (r < n) ? R.ndi[r] : 0;
            rdj[a] = (r < n) ? R.ndj[r] : 0;
        }
#pragma unroll
        for (int b = 0; b < 6; ++b) {
            const int c = lc + 8 * b;
            cdi[b] = (c < n) ? R.ndi[c] : 0;
            cdj[b] = (c < n) ? R.ndj[c] : 0;
        }
#pragma unroll
        for (int a = 0; a < 12; ++a) {
            const int r = lr + 4 * a;
#pragma unroll
            for (int b = 0; b < 6; ++b) {
                const int c = lc + 8 * b;
                A[a][b] = (r < n && c < n) ? lut_cov(s, rdi[a] - cdi[b], rdj[a] - cdj[b]) : 0.0;
            }
            double aug = 0.0;                                 // columns 48 (rho) and 49 (ones) live in lanes lc = 0, 1
            if (r < n) {
                if (lc == 0) {
                    aug = lut_cov(s, -rdi[a], -rdj[a]);
                    R.rho[r] = aug;
                } else if (lc == 1) aug = 1.0;
            }
            A[a][6] = aug;
        }
    }
    // (c) Gauss-Jordan; pivot p = 4 ap + lrp lives in register row ap of the lanes lr == lrp and in register column
    // bp = ap / 2 of the lanes lc == p % 8
#pragma unroll
    for (int ap = 0; ap < 12; ++ap) {
        constexpr int dummy = 0;
        (void)dummy;
        const int bp = ap / 2;
        for (int lrp = 0; lrp < 4; ++lrp) {
            const int p = 4 * ap + lrp;
            if (p >= n) break;                                // warp-uniform
            const int buf = lrp & 1;
            if (lr == lrp) {
#pragma unroll
                for (int b = 0; b < 7; ++b)
                    if (b >= bp) R.xrow[buf][lc + 8 * b] = A[ap][b];
            }
            if (lc == (p & 7)) {
#pragma unroll
                for (int a = 0; a < 12; ++a) R.xcol[buf][lr + 4 * a] = A[a][bp];
            }
            __syncwarp();
            const double inv = __drcp_rn(R.xrow[buf][p]);
            double pr[7];
#pragma unroll
            for (int b = 0; b < 7; ++b)
                if (b >= bp) pr[b] = R.xrow[buf][lc + 8 * b];
#pragma unroll
            for (int a = 0; a < 12; ++a) {
                const int r = lr + 4 * a;
                const double f = (r != p) ? R.xcol[buf][r] * inv : 0.0;
#pragma unroll
                for (int b = 0; b < 7; ++b)
                    if (b >= bp) A[a][b] = fma(-f, pr[b], A[a][b]);
            }
        }
    }
    // (d) gather the diagonal and the two solved right-hand sides, then weights and variance     _krige.py:36-43
#pragma unroll
    for (int a = 0; a < 12; ++a) {
        const int r = lr + 4 * a;
#pragma unroll
        for (int b = 0; b < 6; ++b)
            if (r == lc + 8 * b && r < n) R.dg[r] = A[a][b];
        if (r < n) {
            if (lc == 0) R.ra[r] = A[a][6];
            if (lc == 1) R.rb[r] = A[a][6];
        }
    }
    __syncwarp();
    double xa[2] = {0.0, 0.0}, xb[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int r = lane + 32 * q;
        if (r < n) {
            const double dg = R.dg[r];
            xa[q] = R.ra[r] / dg;
            xb[q] = R.rb[r] / dg;
        }
    }
    double s1a = warp_sum(xa[0] + xa[1]), s1b = warp_sum(xb[0] + xb[1]);
    s1a = __shfl_sync(0xffffffffu, s1a, 0);
    s1b = __shfl_sync(0xffffffffu, s1b, 0);
    const double mu = (s1a - 1.0) / s1b;
    double pv = 0.0;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int r = lane + 32 * q;
        if (r < n) {
            const double w = xa[q] - mu * xb[q];
            R.w[r] = w;
            pv += w * R.rho[r];
        }
    }
    pv = warp_sum(pv);
    if (lane == 0) R.var = fabs(s.sill - pv);
    __syncwarp();
}

template <bool INJECT>
__device__ __forceinline__ void sgs_warp_node(const GmcDev& d, const SgsDev& s, const SgsShared& S, SgsWarpRec& R,
                                              const double* __restrict__ z, int t_ord, int bw, const double* zn_in,
                                              const Philox& rng, uint32_t it_lo, uint32_t it_hi) {
    const int H = d.H, W = d.W, lane = threadIdx.x & 31;
    const int x0 = S.x0, x1 = S.x1, y0 = S.y0, y1 = S.y1;
    const int node = S.todo[t_ord];
    const int bi = node / bw, bj = node - bi * bw;
    const int i = x0 + bi, j = y0 + bj;
    // (a) octant search, octants in the reference's order                                       neighbors.py:52-60
    // The lists are sorted by distance, so the offsets closer than a level's radius are a prefix: a node that finds nothing
    // within `radius` searches again with radius + 100 km, as the reference does (MCMC.py:149-155).
    int n = 0;
    for (int lev = 0; lev < s.n_levels && n == 0; ++lev)
    for (int o = 0; o < 8; ++o) {
        const int16_t* off = s.oct_off + (int64_t)o * s.lmax * 2;
        const int cnt = __ldg(s.oct_cnt + lev * 8 + o);
        int found = 0;
        for (int base = 0; base < cnt && found < s.per_oct; base += 32) {
            const int t = base + lane;
            bool ok = false;
            int di = 0, dj = 0, src = -1;
            double v = 0.0;
            if (t < cnt) {
                if (t < SGS_NEAR) {
                    const short2 o2 = S.near_off[o][t];
                    di = o2.x;
                    dj = o2.y;
                } else {
                    di = off[2 * t];
                    dj = off[2 * t + 1];
                }
                const int ci = i + di, cj = j + dj;
                if (ci >= 0 && ci < H && cj >= 0 && cj < W) {
                    if (ci >= x0 && ci < x1 && cj >= y0 && cj < y1) {
                        src = (ci - x0) * bw + (cj - y0);
                        ok = S.ord[src] < t_ord;                 // radar cell (-1) or simulated earlier in the path
                    } else {
                        v = __ldcg(z + (int64_t)ci * W + cj);
                        ok = (v == v);
                    }
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            const int rank = found + __popc(m & ((1u << lane) - 1u));
            if (ok && rank < s.per_oct) {
                const int slot = n + rank;
                R.ndi[slot] = (int16_t)di;
                R.ndj[slot] = (int16_t)dj;
                R.nsrc[slot] = (int16_t)src;
                R.nval[slot] = v;
            }
            found += __popc(m);
        }
        n += min(found, s.per_oct);
    }
    if (lane == 0) {
        R.n = n;
        R.node = node;
        R.kpath = S.todo_k[t_ord];
        double zn, z1;
        if (INJECT) zn = zn_in[S.todo_k[t_ord]];
        else box_muller(rng((uint32_t)node, it_lo, it_hi, 6u), zn, z1);               // stream 6: node normals
        R.zn = zn;
    }
    __syncwarp();
    if (n == 0) return;                                      // reported by phase 2 (the reference would widen the radius)
    sgs_warp_solve(s, R, n);
}

template <bool INJECT, bool WS>
__device__ void sgs_one_step(const GmcDev& d, const SgsDev& s, SgsShared& S, double* bedc, double* z, double* mcres, double& ssq,
                             int& nviol, const int32_t* path_in, const double* zn_in, const Philox& rng, uint32_t it_lo,
                             uint32_t it_hi, int32_t* resampled, double* loss_next_out) {
    const int H = d.H, W = d.W, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int x0 = S.x0, x1 = S.x1, y0 = S.y0, y1 = S.y1;
    const int bh = x1 - x0, bw = y1 - y0, nblk = bh * bw;
    long long tclk = (s.phase && tid == 0) ? clock64() : 0;
    auto mark = [&](int ph) {
        if (s.phase && tid == 0) {
            const long long now = clock64();
            atomicAdd(reinterpret_cast<unsigned long long*>(s.phase + ph), (unsigned long long)(now - tclk));
            tclk = now;
        }
    };
    // (1) block <- normal-scored conditioning data (NaN where there is none)              MCMC.py:1771
    for (int e = tid; e < nblk; e += SGS_THREADS) {
        const int bi = e / bw, bj = e - bi * bw;
        S.blk_z[e] = __ldg(s.zcond + (int64_t)(x0 + bi) * W + (y0 + bj));
    }
    // path: injected order, or a uniformly random permutation by sorting Philox keys         MCMC.py:125
    if (INJECT) {
        for (int e = tid; e < nblk; e += SGS_THREADS) S.path[e] = path_in[e];
    } else {
        unsigned long long* keys = reinterpret_cast<unsigned long long*>(S.sig);      // the kriging matrix is idle here
        int npad = 1;
        while (npad < nblk) npad <<= 1;
        for (int e = tid; e < npad; e += SGS_THREADS) {
            unsigned long long k = ~0ull;
            if (e < nblk) {
                const uint4 r = rng((uint32_t)e, it_lo, it_hi, 5u);          // stream 5: path keys
                k = ((((unsigned long long)r.x << 32) | r.y) & ~0x3ffull) | (unsigned long long)e;   // unique: index in low bits
            }
            keys[e] = k;
        }
        __syncthreads();
        for (int k2 = 2; k2 <= npad; k2 <<= 1)                                 // bitonic sort, ascending
            for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                for (int e = tid; e < npad; e += SGS_THREADS) {
                    const int p = e ^ j2;
                    if (p > e) {
                        const unsigned long long a = keys[e], b = keys[p];
                        const bool up = (e & k2) == 0;
                        if ((a > b) == up) { keys[e] = b; keys[p] = a; }
                    }
                }
                __syncthreads();
            }
        for (int e = tid; e < nblk; e += SGS_THREADS) S.path[e] = (int)(keys[e] & 0x3ffull);
        __syncthreads();
    }
    __syncthreads();

    mark(0);
    if constexpr (WS) {
        // (2w) warp-per-node simulation: list the nodes to simulate in path order ...
        if (wid == 0) {
            int count = 0;
            for (int base = 0; base < nblk; base += 32) {
                const int k = base + lane;
                int node = 0;
                bool unc = false;
                if (k < nblk) {
                    node = S.path[k];
                    const double cur = S.blk_z[node];
                    unc = !(cur == cur);
                }
                const unsigned m = __ballot_sync(0xffffffffu, unc);
                if (k < nblk) {
                    if (unc) {
                        const int t = count + __popc(m & ((1u << lane) - 1u));
                        S.todo[t] = (int16_t)node;
                        S.todo_k[t] = (int16_t)k;
                        S.ord[node] = (int16_t)t;
                    } else S.ord[node] = -1;
                }
                count += __popc(m);
            }
            if (lane == 0) S.n_todo = count;
        }
        __syncthreads();
        const int n_todo = S.n_todo;
        SgsWarpRec* recs = reinterpret_cast<SgsWarpRec*>(S.sig);
        // ... every warp walks its own nodes t = wid, wid + 8, ... of the path: search + solve (no simulated value is read),
        // then the value, for which the in-block neighbours - all EARLIER in the path - must have been simulated: a cell
        // still to be simulated holds NaN in S.blk_z, so a lane re-reads its neighbour until it is a number.  The warp on
        // the earliest unfinished node never waits (no deadlock), and there is no CTA barrier and no serial values pass
        // between batches of eight nodes any more (they cost 28 % of the node time: profiles/README.md).
        volatile double* blk = S.blk_z;
        for (int t = wid; t < n_todo; t += 8) {
            SgsWarpRec& R = recs[wid];
            sgs_warp_node<INJECT>(d, s, S, R, z, t, bw, zn_in, rng, it_lo, it_hi);
            const int n = R.n;
            double val = 0.0;
            if (n == 0) {
                if (lane == 0) S.err = 1;
            } else {
                double v[2] = {0.0, 0.0}, wv[2] = {0.0, 0.0};
                int src[2] = {-1, -1};
                bool need[2] = {false, false};
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r = lane + 32 * h2;
                    if (r < n) {
                        src[h2] = R.nsrc[r];
                        wv[h2] = R.w[r];
                        if (src[h2] >= 0) need[h2] = true;
                        else v[h2] = R.nval[r];
                    }
                }
                unsigned spins = 0;
                for (;;) {                                    // MCMC.py:163-169: values in path order
                    bool pending = false;
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2)
                        if (need[h2]) {
                            v[h2] = blk[src[h2]];
                            if (v[h2] == v[h2]) need[h2] = false;
                            else pending = true;
                        }
                    if (!__any_sync(0xffffffffu, pending)) break;
                    if (++spins > (1u << 22)) break;          // a genuinely NaN neighbour: carry it, never hang
                }
                double sv = warp_sum(v[0] + v[1]);
                sv = __shfl_sync(0xffffffffu, sv, 0);
                const double mean = sv / (double)n;
                double pe = 0.0;
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2)
                    if (lane + 32 * h2 < n) pe += wv[h2] * (v[h2] - mean);
                pe = warp_sum(pe);
                val = (mean + pe) + sqrt(R.var) * R.zn;                                  // MCMC.py:168
            }
            if (lane == 0) {
                blk[R.node] = val;
                __threadfence_block();
            }
            __syncwarp();
        }
        __syncthreads();
        mark(4);
        mark(5);
    } else {
    // (2) sequential simulation along the path                                              MCMC.py:130-169
    for (int k = 0; k < nblk; ++k) {
        const int node = S.path[k];
        const double cur = S.blk_z[node];
        if (cur == cur) continue;                       // conditioned (radar value): uniform across the CTA
        const int bi = node / bw, bj = node - bi * bw;
        const int i = x0 + bi, j = y0 + bj;
        // (a) octant search: warp `wid` owns octant wid (b = wid - 4)                           neighbors.py:52-60
        int lev = 0;
    search_again:
        {
            const int16_t* off = s.oct_off + (int64_t)wid * s.lmax * 2;
            const int cnt = __ldg(s.oct_cnt + lev * 8 + wid);
            int found = 0;
            for (int base = 0; base < cnt && found < s.per_oct; base += 32) {
                const int t = base + lane;
                bool ok = false;
                int di = 0, dj = 0;
                double v = 0.0;
                if (t < cnt) {
                    if (t < SGS_NEAR) {
                        const short2 o2 = S.near_off[wid][t];
                        di = o2.x;
                        dj = o2.y;
                    } else {
                        di = off[2 * t];
                        dj = off[2 * t + 1];
                    }
                    const int ci = i + di, cj = j + dj;
                    if (ci >= 0 && ci < H && cj >= 0 && cj < W) {
                        if (ci >= x0 && ci < x1 && cj >= y0 && cj < y1) v = S.blk_z[(ci - x0) * bw + (cj - y0)];
                        else v = __ldcg(z + (int64_t)ci * W + cj);
                        ok = (v == v);
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int rank = found + __popc(m & ((1u << lane) - 1u));
                if (ok && rank < s.per_oct) {
                    const int slot = wid * s.per_oct + rank;
                    S.ndi[slot] = (int16_t)di;
                    S.ndj[slot] = (int16_t)dj;
                    S.nval[slot] = v;
                }
                found += __popc(m);
            }
            if (lane == 0) S.oct_n[wid] = min(found, s.per_oct);
        }
        __syncthreads();
        mark(1);
        // (b) compact the octant lists (octant order -4..3, each sorted by distance) -> n neighbours
        int n = 0, start[8];
#pragma unroll
        for (int o = 0; o < 8; ++o) {
            start[o] = n;
            n += S.oct_n[o];
        }
        if (n == 0 && lev + 1 < s.n_levels) {           // nothing within this radius: widen by 100 km (MCMC.py:149-155)
            ++lev;
            __syncthreads();
            goto search_again;
        }
        if (n == 0) {                                   // no conditioned cell within the widest table either
            if (tid == 0) S.err = 1;
            if (tid == 0) S.blk_z[node] = 0.0;
            __syncthreads();
            continue;
        }
        // read my neighbour (threads < 64 hold one slot each), then write it to its compacted position
        int mdi = 0, mdj = 0, dst = -1;
        double mval = 0.0;
        if (tid < 8 * s.per_oct) {
            const int o = tid / s.per_oct, r = tid - o * s.per_oct;
            if (r < S.oct_n[o]) {
                mdi = S.ndi[tid];
                mdj = S.ndj[tid];
                mval = S.nval[tid];
                dst = start[o] + r;
            }
        }
        __syncthreads();
        if (dst >= 0) {
            S.ndi[dst] = (int16_t)mdi;
            S.ndj[dst] = (int16_t)mdj;
            S.nval[dst] = mval;
        }
        __syncthreads();
        mark(2);
        // (c) assemble Sigma (n x n), rho and the ones vector                                   _krige.py:20-33
        for (int t0 = tid; t0 < n * n; t0 += 4 * SGS_THREADS) {          // 4 independent table loads in flight per thread
            double cv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q * SGS_THREADS;
                const int a = (t < n * n) ? t / n : 0, b = (t < n * n) ? t - a * n : 0;
                cv[q] = lut_cov(s, S.ndi[a] - S.ndi[b], S.ndj[a] - S.ndj[b]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int t = t0 + q * SGS_THREADS;
                if (t < n * n) S.sig[(t / n) * SGS_SIG_PITCH + (t % n)] = cv[q];
            }
        }
        if (tid < n) {
            S.rhs_a[tid] = lut_cov(s, -S.ndi[tid], -S.ndj[tid]);
            S.rhs_b[tid] = 1.0;
        }
        __syncthreads();
        // rho is needed again for the variance: keep a register copy in the first n threads
        const double rho_keep = (tid < n) ? S.rhs_a[tid] : 0.0;
        mark(3);
        // (d) Gauss-Jordan elimination without pivoting (SPD block) on the augmented matrix [Sigma | rho | 1], held in
        // REGISTERS: thread (ty, tx) of a 16 x 16 layout owns rows ty+16a (a<4) and columns tx+16b (b<5) — a cyclic
        // distribution, so the shrinking active region stays balanced.  Per pivot only the pivot row, the pivot column
        // and 1/pivot travel through shared memory (double-buffered: one barrier per pivot); the n^3/2 multiply-adds run
        // on registers with every lane busy.
        if (tid < n) {
            S.sig[tid * SGS_SIG_PITCH + n] = S.rhs_a[tid];
            S.sig[tid * SGS_SIG_PITCH + n + 1] = S.rhs_b[tid];
        }
        __syncthreads();
        {
            const int ty = tid >> 4, tx = tid & 15;
            double A[4][5];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 5; ++b2) {
                    const int r = ty + 16 * a, c = tx + 16 * b2;
                    A[a][b2] = (r < n && c < n + 2) ? S.sig[r * SGS_SIG_PITCH + c] : 0.0;
                }
            __syncthreads();                                 // S.sig is reused below as the exchange buffers
            double* xrow = S.sig;                            // [2][80] pivot row, [2][64] pivot column, [2] 1/pivot
            double* xcol = S.sig + 2 * 80;
            double* xinv = S.sig + 2 * 80 + 2 * 64;
            for (int p = 0; p < n; ++p) {
                const int buf = p & 1, pa = p >> 4, pb = p >> 4, pty = p & 15, ptx = p & 15;
                if (ty == pty) {                             // owners of the pivot row publish their entries
#pragma unroll
                    for (int a = 0; a < 4; ++a)
                        if (a == pa) {
#pragma unroll
                            for (int b2 = 0; b2 < 5; ++b2) xrow[buf * 80 + tx + 16 * b2] = A[a][b2];
                            if (tx == ptx) {
#pragma unroll
                                for (int b2 = 0; b2 < 5; ++b2)
                                    if (b2 == pb) xinv[buf] = 1.0 / A[a][b2];
                            }
                        }
                }
                if (tx == ptx) {                             // owners of the pivot column publish theirs
#pragma unroll
                    for (int b2 = 0; b2 < 5; ++b2)
                        if (b2 == pb) {
#pragma unroll
                            for (int a = 0; a < 4; ++a) xcol[buf * 64 + ty + 16 * a] = A[a][b2];
                        }
                }
                __syncthreads();
                const double inv = xinv[buf];
                double pr[5];
#pragma unroll
                for (int b2 = 0; b2 < 5; ++b2) pr[b2] = xrow[buf * 80 + tx + 16 * b2];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = ty + 16 * a;
                    const double f = (r != p) ? xcol[buf * 64 + r] * inv : 0.0;
#pragma unroll
                    for (int b2 = 0; b2 < 5; ++b2)
                        if (tx + 16 * b2 > p) A[a][b2] -= f * pr[b2];
                }
            }
            __syncthreads();
            // publish the diagonal and the two solved right-hand sides
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 5; ++b2) {
                    const int r = ty + 16 * a, c = tx + 16 * b2;
                    if (r < n && (c == r || c == n || c == n + 1)) S.sig[r * SGS_SIG_PITCH + c] = A[a][b2];
                }
            __syncthreads();
        }
        if (tid < n) {
            const double dg = S.sig[tid * SGS_SIG_PITCH + tid];
            S.rhs_a[tid] = S.sig[tid * SGS_SIG_PITCH + n] / dg;
            S.rhs_b[tid] = S.sig[tid * SGS_SIG_PITCH + n + 1] / dg;
        }
        __syncthreads();
        mark(4);
        // (e) Lagrange multiplier, weights, estimate and variance (warp 0)                        _krige.py:36-43
        if (wid == 0) {
            double s1a = 0.0, s1b = 0.0, sv = 0.0;
            for (int t = lane; t < n; t += 32) {
                s1a += S.rhs_a[t];
                s1b += S.rhs_b[t];
                sv += S.nval[t];
            }
            s1a = warp_sum(s1a);
            s1b = warp_sum(s1b);
            sv = warp_sum(sv);
            s1a = __shfl_sync(0xffffffffu, s1a, 0);
            s1b = __shfl_sync(0xffffffffu, s1b, 0);
            sv = __shfl_sync(0xffffffffu, sv, 0);
            const double mu = (s1a - 1.0) / s1b;
            for (int t = lane; t < n; t += 32) S.rhs_a[t] = S.rhs_a[t] - mu * S.rhs_b[t];      // kriging weights
            if (lane == 0) S.scratch[36] = sv / (double)n;                                     // local mean of the neighbours
        }
        __syncthreads();
        {
            // sum w_i rho_i and sum w_i (v_i - mean) over the first n threads (rho_keep lives in their registers)
            const double mean = S.scratch[36];
            double pv = 0.0, pe = 0.0;
            if (tid < n) {
                const double w = S.rhs_a[tid];
                pv = w * rho_keep;
                pe = w * (S.nval[tid] - mean);
            }
            const double sum_v = block_sum<SGS_THREADS>(pv, S.scratch);
            const double sum_e = block_sum<SGS_THREADS>(pe, S.scratch);
            if (tid == 0) {
                const double var = fabs(s.sill - sum_v);
                const double est = mean + sum_e;
                double zn;
                if (INJECT) zn = zn_in[k];
                else {
                    double z1;
                    box_muller(rng((uint32_t)node, it_lo, it_hi, 6u), zn, z1);       // stream 6: node normals
                }
                S.blk_z[node] = est + sqrt(var) * zn;                                  // MCMC.py:168
            }
        }
        __syncthreads();
    }
    }

    mark(5);
    // (3) candidate bed of the block = inverse normal score of the simulated values            MCMC.py:1776-1777
    for (int e = tid; e < nblk; e += SGS_THREADS) S.cand[e] = nst_inverse(s, S.blk_z[e]);
    __syncthreads();
    auto full_bed = [&](int i, int j) -> double {
        const int64_t idx = (int64_t)i * W + j;
        const double bc = (i >= x0 && i < x1 && j >= y0 && j < y1) ? S.cand[(i - x0) * bw + (j - y0)] : __ldcg(bedc + idx);
        return s.trend ? add_rn(bc, __ldg(s.trend + idx)) : bc;
    };
    // residual on the block plus its one-cell ring (everything a changed bed cell can influence)
    const int rx0 = max(x0 - 1, 0), rx1 = min(x1 + 1, H), ry0 = max(y0 - 1, 0), ry1 = min(y1 + 1, W);
    const int rh = rx1 - rx0, rw = ry1 - ry0;
    double delta = 0.0;
    int dviol = 0;
    for (int e = tid; e < rh * rw; e += SGS_THREADS) {
        const int ri = e / rw, rj = e - ri * rw;
        const int i = rx0 + ri, j = ry0 + rj;
        const int64_t idx = (int64_t)i * W + j;
        const int jl = max(j - 1, 0), jr = min(j + 1, W - 1), iu = max(i - 1, 0), id = min(i + 1, H - 1);
        const bool ex = (j == 0) || (j == W - 1), ey = (i == 0) || (i == H - 1);
        const double2 xr = __ldg(d.sv + (int64_t)i * W + jr), xl = __ldg(d.sv + (int64_t)i * W + jl);
        const double2 yd = __ldg(d.sy + (int64_t)id * W + j), yu = __ldg(d.sy + (int64_t)iu * W + j);
        const double2 hs = __ldg(d.ds + idx);
        const double fr = mul_rn(xr.y, sub_rn(xr.x, full_bed(i, jr))), fl = mul_rn(xl.y, sub_rn(xl.x, full_bed(i, jl)));
        const double fd = mul_rn(yd.y, sub_rn(yd.x, full_bed(id, j))), fu = mul_rn(yu.y, sub_rn(yu.x, full_bed(iu, j)));
        const double dx = div_const(sub_rn(fr, fl), ex ? d.res : d.two_res, ex ? d.r_res : d.r_two_res);
        const double dy = div_const(sub_rn(fd, fu), ey ? d.res : d.two_res, ey ? d.r_res : d.r_two_res);
        const double rnew = sub_rn(add_rn(add_rn(dx, dy), hs.x), hs.y);
        S.newres[e] = rnew;
        if (__ldg(d.flags + idx) & FLAG_MC) {
            const double rold = __ldcg(mcres + idx);
            delta += ((rnew == rnew) ? rnew * rnew : 0.0) - ((rold == rold) ? rold * rold : 0.0);
        }
        if (i >= x0 && i < x1 && j >= y0 && j < y1 && __ldg(s.grounded + idx) == 1) {       // guard delta on the block
            const double sf = __ldg(d.surf + idx);
            const double oldfull = s.trend ? add_rn(__ldcg(bedc + idx), __ldg(s.trend + idx)) : __ldcg(bedc + idx);
            dviol += (sub_rn(sf, full_bed(i, j)) <= 0.0) - (sub_rn(sf, oldfull) <= 0.0);
        }
    }
    const double dsum = block_sum<SGS_THREADS>(delta, S.scratch);
    const double dv = block_sum<SGS_THREADS>((double)dviol, S.scratch);
    if (tid == 0) {
        const int nviol_next = nviol + (int)dv;
        const double ssq_next = ssq + dsum;
        const double loss_prev = div_rn(ssq, d.two_sigma2);
        double loss_next = div_rn(ssq_next, d.two_sigma2);
        if (nviol_next > 0) loss_next = __longlong_as_double(0x7ff0000000000000LL);        // MCMC.py:1794-1795
        double acc;
        if (loss_prev > loss_next) acc = 1.0;
        else {
            const double ex = exp(loss_prev - loss_next);
            acc = (ex < 1.0) ? ex : 1.0;
        }
        S.accept = (S.u <= acc) ? 1 : 0;
        S.scratch[34] = ssq_next;
        S.scratch[35] = loss_next;
        S.scratch[37] = (double)nviol_next;
    }
    __syncthreads();
    if (loss_next_out && tid == 0) *loss_next_out = S.scratch[35];
    if (S.accept) {
        ssq = S.scratch[34];
        nviol = (int)S.scratch[37];
        for (int e = tid; e < nblk; e += SGS_THREADS) {
            const int bi = e / bw, bj = e - bi * bw;
            const int64_t idx = (int64_t)(x0 + bi) * W + (y0 + bj);
            const double c = S.cand[e];
            __stcg(bedc + idx, c);
            __stcg(z + idx, nst_forward(s, c));          // the reference re-transforms the accepted bed every step (MCMC.py:1767)
            if (resampled) __stcg(resampled + idx, __ldcg(resampled + idx) + 1);          // MCMC.py:1806
        }
        for (int e = tid; e < rh * rw; e += SGS_THREADS) {
            const int ri = e / rw, rj = e - ri * rw;
            __stcg(mcres + (int64_t)(rx0 + ri) * W + (ry0 + rj), S.newres[e]);
        }
    }
    __syncthreads();
    mark(6);
}

__device__ __forceinline__ void sgs_stage_offsets(const SgsDev& s, SgsShared& S) {
    for (int t = threadIdx.x; t < 8 * SGS_NEAR; t += SGS_THREADS) {
        const int o = t / SGS_NEAR, k = t - o * SGS_NEAR;
        short2 v = make_short2(0, 0);
        if (k < s.lmax) v = make_short2(s.oct_off[((int64_t)o * s.lmax + k) * 2], s.oct_off[((int64_t)o * s.lmax + k) * 2 + 1]);
        S.near_off[o][k] = v;
    }
}

__device__ __forceinline__ void sgs_window(SgsShared& S, int H, int W) {
    // MCMC.py:1759-1762: int(index - size/2) truncates toward zero, then clamps
    S.x0 = max(0, (int)((double)S.ix - (double)S.bsx / 2.0));
    S.x1 = min(H, (int)((double)S.ix + (double)S.bsx / 2.0));
    S.y0 = max(0, (int)((double)S.iy - (double)S.bsy / 2.0));
    S.y1 = min(W, (int)((double)S.iy + (double)S.bsy / 2.0));
}

template <bool WS>
__global__ void __launch_bounds__(SGS_THREADS)
    sgs_step_injected_kernel(GmcDev d, SgsDev s, double* bedc_all, double* z_all, double* mcres_all, double* ssq_all,
                             int32_t* nviol_all, const int32_t* __restrict__ centre, const int32_t* __restrict__ bs,
                             const int32_t* __restrict__ path, const double* __restrict__ zn, int64_t path_stride,
                             const double* __restrict__ u, uint8_t* accepted_out, double* loss_out, double* loss_next_out,
                             int32_t* resampled_all, int32_t* err_out) {
    extern __shared__ __align__(16) unsigned char sgs_raw[];
    SgsShared& S = *reinterpret_cast<SgsShared*>(sgs_raw);
    const int c = blockIdx.x;
    const int64_t plane = (int64_t)d.H * d.W;
    sgs_stage_offsets(s, S);
    if (threadIdx.x == 0) {
        S.ix = centre[2 * c];
        S.iy = centre[2 * c + 1];
        S.bsx = bs[2 * c];
        S.bsy = bs[2 * c + 1];
        S.u = u[c];
        S.err = 0;
        sgs_window(S, d.H, d.W);
    }
    __syncthreads();
    double ssq = ssq_all[c];
    int nviol = nviol_all[c];
    const Philox rng(0ull);
    sgs_one_step<true, WS>(d, s, S, bedc_all + c * plane, z_all + c * plane, mcres_all + c * plane, ssq, nviol, path + c * path_stride,
                       zn + c * path_stride, rng, 0u, 0u, resampled_all ? resampled_all + c * plane : nullptr,
                       loss_next_out ? loss_next_out + c : nullptr);
    if (threadIdx.x == 0) {
        ssq_all[c] = ssq;
        nviol_all[c] = nviol;
        if (accepted_out) accepted_out[c] = (uint8_t)S.accept;
        if (loss_out) loss_out[c] = div_rn(ssq, d.two_sigma2);
        if (err_out && S.err) atomicOr(err_out, 1);
    }
}

// sched (may be NULL): as in run_kernel (step.cu) - with more chains than resident CTAs the iterations are cut into chunks
// and the (chunk, chain) items are drawn from sched[0]; sched[1 + chain] counts the chain's completed chunks.  The chain
// state already travels through L2 only (ld.cg / st.cg), so a chain may continue on any CTA.
template <bool WS>
__global__ void __launch_bounds__(SGS_THREADS)
    sgs_run_kernel(GmcDev d, SgsDev s, double* bedc_all, double* z_all, double* mcres_all, double* ssq_all, int32_t* nviol_all,
                   const uint64_t* __restrict__ seeds, uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache,
                   int32_t* blocks_cache, int64_t cache_stride, int64_t cache_offset, int32_t* resampled_all, int32_t* err_out,
                   int C, int* sched, int chunk, int* dev_err, unsigned spin_limit) {
    extern __shared__ __align__(16) unsigned char sgs_raw[];
    SgsShared& S = *reinterpret_cast<SgsShared*>(sgs_raw);
    __shared__ long long s_item;
    const int64_t plane = (int64_t)d.H * d.W;
    if (threadIdx.x == 0) S.err = 0;
    sgs_stage_offsets(s, S);
    const int n_chunks = sched ? (n_steps + chunk - 1) / chunk : 1;
    const long long n_items = (long long)n_chunks * C;
    for (long long item = blockIdx.x;; item += gridDim.x) {
        if (sched) {
            if (threadIdx.x == 0) {
                const long long it2 = atomicAdd(reinterpret_cast<unsigned int*>(sched), 1u);
                if (it2 < n_items) {
                    volatile int* done = sched + 1 + (int)(it2 % C);
                    const int jj = (int)(it2 / C);
                    unsigned spins = 0;
                    while (*done < jj) {
                        if (*(volatile int*)dev_err) break;   // a wait already gave up: the launch is void, do not wait again
                        __nanosleep(200);
                        if (++spins > spin_limit) {           // an item only waits for items drawn earlier; never hang, and
                            *(volatile int*)dev_err = GMC_DEVERR_WAIT_TIMEOUT;   // never carry on silently (host: GMC_ECUDA)
                            __threadfence_system();
                            break;
                        }
                    }
                    __threadfence();
                }
                s_item = it2;
            }
            __syncthreads();
            item = s_item;
        }
        if (item >= n_items) break;
        const int c = (int)(item % C), j = (int)(item / C);
        const int k0 = sched ? j * chunk : 0, k1 = sched ? min(n_steps, k0 + chunk) : n_steps;
        const Philox rng(seeds[c]);
        double ssq = __ldcg(ssq_all + c);
        int nviol = __ldcg(nviol_all + c);
        for (int k = k0; k < k1; ++k) {
            const uint64_t it = iter0 + (uint64_t)k;
            const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
            if (threadIdx.x == 0) {
                // chain stream: centre (uniform over region cells), block sizes (upper bound exclusive), acceptance uniform
                const uint4 c0 = rng(0u, it_lo, it_hi, GMC_STREAM_CHAIN);
                const uint4 c1 = rng(1u, it_lo, it_hi, GMC_STREAM_CHAIN);
                const uint4 c2 = rng(2u, it_lo, it_hi, GMC_STREAM_CHAIN);
                if (d.n_centre_cells > 0) {
                    const int32_t cell = d.centre_cells[bounded_u64(c0.x, c0.y, (uint64_t)d.n_centre_cells)];
                    S.ix = cell / d.W;
                    S.iy = cell - S.ix * d.W;
                } else {
                    S.ix = (int)bounded_u64(c0.x, c0.y, (uint64_t)d.H);
                    S.iy = (int)bounded_u64(c0.z, c0.w, (uint64_t)d.W);
                }
                S.u = u01_halfopen(c1.x, c1.y);
                S.bsx = s.bmin_x + (int)bounded_u64(c2.x, c2.y, (uint64_t)(s.bmax_x - s.bmin_x));
                S.bsy = s.bmin_y + (int)bounded_u64(c2.z, c2.w, (uint64_t)(s.bmax_y - s.bmin_y));
                sgs_window(S, d.H, d.W);
            }
            __syncthreads();
            sgs_one_step<false, WS>(d, s, S, bedc_all + c * plane, z_all + c * plane, mcres_all + c * plane, ssq, nviol, nullptr,
                                    nullptr, rng, it_lo, it_hi, resampled_all ? resampled_all + c * plane : nullptr, nullptr);
            if (threadIdx.x == 0) {
                const int64_t slot = (int64_t)c * cache_stride + cache_offset + k;
                if (loss_cache) loss_cache[slot] = div_rn(ssq, d.two_sigma2);
                if (step_cache) step_cache[slot] = (uint8_t)S.accept;
                if (blocks_cache) reinterpret_cast<int4*>(blocks_cache)[slot] = make_int4(S.ix, S.iy, S.bsx, S.bsy);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            __stcg(ssq_all + c, ssq);
            __stcg(nviol_all + c, nviol);
        }
        if (!sched) break;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(sched + 1 + c, j + 1);
    }
    if (threadIdx.x == 0 && err_out && S.err) atomicOr(err_out, 1);
}

// ---------------------------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
static int upload(const T* src, size_t n, T** dst, void** slot) {
    GMC_CUDA(cudaMalloc(dst, n * sizeof(T)));
    GMC_CUDA(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyDefault));
    *slot = *dst;
    return GMC_OK;
}

static void sgs_free(gmc_sgs_state* st) {
    if (!st) return;
    for (void*& p : st->owned) {
        cudaFree(p);
        p = nullptr;
    }
}

extern "C" int gmc_sgs_setup(gmc_ctx* c, const double* trend, const double* zcond, const uint8_t* grounded,
                             const double* quantiles, const double* references, int n_quantiles, const int16_t* oct_off,
                             const int32_t* oct_cnt, int n_levels, int lmax, int hw, int num_points, const double* lut, double sill,
                             int block_min_x, int block_max_x, int block_min_y, int block_max_y) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: ctx is NULL");
    if (n_levels < 1 || n_levels > SGS_MAX_LEVELS)
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: n_levels=%d outside [1,%d]", n_levels, SGS_MAX_LEVELS);
    if (!c->have_static) GMC_FAIL(GMC_ESTATE, "gmc_sgs_setup: call gmc_set_static first");
    if (!zcond || !grounded || !oct_off || !oct_cnt || !lut) GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: NULL argument");
    if ((quantiles == nullptr) != (references == nullptr) || (quantiles && n_quantiles < 2))
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: quantiles/references must both be given (>= 2 entries) or both be NULL");
    if (num_points < 8 || num_points > SGS_MAX_NEIGH)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_sgs_setup: num_points=%d outside [8,%d] (the reference takes num_points//8 per octant)", num_points, SGS_MAX_NEIGH);
    if (hw < 1 || lmax < 1) GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: empty search stencil");
    if (block_min_x < 1 || block_min_y < 1 || block_max_x <= block_min_x || block_max_y <= block_min_y)
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_setup: block sizes are drawn from [min, max) and need max > min >= 1 (MCMC.py:1755-1756)");
    if (block_max_x - 1 > SGS_MAX_BLOCK || block_max_y - 1 > SGS_MAX_BLOCK)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_sgs_setup: blocks larger than %d cells per edge are not supported", SGS_MAX_BLOCK);
    GMC_CUDA(cudaSetDevice(c->device));
    if (!c->sgs) {
        c->sgs = new gmc_sgs_state();
        memset(c->sgs, 0, sizeof(gmc_sgs_state));
    }
    gmc_sgs_state* st = c->sgs;
    sgs_free(st);
    st->ready = false;
    const size_t n = (size_t)c->H * c->W;
    SgsDev& s = st->dev;
    memset(&s, 0, sizeof(s));
    double* dp = nullptr;
    uint8_t* up = nullptr;
    int16_t* sp = nullptr;
    int32_t* ip = nullptr;
    int rc;
    if (trend) {
        if ((rc = upload(trend, n, &dp, &st->owned[0]))) return rc;
        s.trend = dp;
    }
    if ((rc = upload(zcond, n, &dp, &st->owned[1]))) return rc;
    s.zcond = dp;
    if ((rc = upload(grounded, n, &up, &st->owned[2]))) return rc;
    s.grounded = up;
    if (quantiles) {
        if ((rc = upload(quantiles, (size_t)n_quantiles, &dp, &st->owned[3]))) return rc;
        s.quant = dp;
        if ((rc = upload(references, (size_t)n_quantiles, &dp, &st->owned[4]))) return rc;
        s.refs = dp;
        s.nq = n_quantiles;
    }
    if ((rc = upload(oct_off, (size_t)8 * lmax * 2, &sp, &st->owned[5]))) return rc;
    s.oct_off = sp;
    if ((rc = upload(oct_cnt, (size_t)8 * n_levels, &ip, &st->owned[6]))) return rc;
    s.oct_cnt = ip;
    s.n_levels = n_levels;
    s.lut_w = 4 * hw + 1;
    if ((rc = upload(lut, (size_t)s.lut_w * s.lut_w, &dp, &st->owned[7]))) return rc;
    s.lut = dp;
    s.lmax = lmax;
    s.hw = hw;
    s.per_oct = num_points / 8;
    s.sill = sill;
    s.bmin_x = block_min_x;
    s.bmax_x = block_max_x;
    s.bmin_y = block_min_y;
    s.bmax_y = block_max_y;
    GMC_CUDA(cudaFuncSetAttribute(sgs_step_injected_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SgsShared)));
    GMC_CUDA(cudaFuncSetAttribute(sgs_run_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SgsShared)));
    GMC_CUDA(cudaFuncSetAttribute(sgs_step_injected_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SgsShared)));
    GMC_CUDA(cudaFuncSetAttribute(sgs_run_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SgsShared)));
    // warp-per-node solver for up to 48 neighbours; GMC_SGS_SOLVER=cta forces the CTA-wide solver (A/B timing, tests)
    const char* force = getenv("GMC_SGS_SOLVER");
    st->warp_solver = (s.per_oct * 8 <= SGS_WN) && !(force && force[0] == 'c');
    st->ready = true;
    return GMC_OK;
}

void gmc_sgs_destroy(gmc_ctx* c) {
    if (c && c->sgs) {
        sgs_free(c->sgs);
        delete c->sgs;
        c->sgs = nullptr;
    }
}

static int sgs_check(gmc_ctx* c, int C, const char* who) {
    if (!c) GMC_FAIL(GMC_EINVAL, "%s: ctx is NULL", who);
    if (!c->sgs || !c->sgs->ready) GMC_FAIL(GMC_ESTATE, "%s: call gmc_sgs_setup first", who);
    if (C < 1 || C > c->max_chains) GMC_FAIL(GMC_ESHAPE, "%s: C=%d outside [1,%d]", who, C, c->max_chains);
    GMC_CUDA(cudaSetDevice(c->device));
    return GMC_OK;
}

extern "C" int gmc_sgs_transform(gmc_ctx* c, const double* in, double* out, int64_t n, int inverse, void* stream) {
    int rc = sgs_check(c, 1, "gmc_sgs_transform");
    if (rc) return rc;
    if (!in || !out || n < 1) GMC_FAIL(GMC_EINVAL, "gmc_sgs_transform: bad argument");
    sgs_transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c->sgs->dev, in, out, n, inverse);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_sgs_init(gmc_ctx* c, const double* bed, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol,
                            double* scratch_full, int C, void* stream) {
    int rc = sgs_check(c, C, "gmc_sgs_init");
    if (rc) return rc;
    if (!bed || !bedc || !z || !mcres || !ssq || !nviol || !scratch_full) GMC_FAIL(GMC_EINVAL, "gmc_sgs_init: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)C * c->H * c->W;
    GMC_CUDA(cudaMemsetAsync(nviol, 0, (size_t)C * sizeof(int32_t), st));
    sgs_init_kernel<<<dim3(64, C), 256, 0, st>>>(c->dev, c->sgs->dev, bed, bedc, z, nviol, C);
    sgs_fullbed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c->dev, c->sgs->dev, bedc, scratch_full, n);
    c->launches += 2;
    GMC_CUDA(cudaGetLastError());
    // residual + loss of the (re-assembled) full bed, like MCMC.py:1663-1669
    return gmc_residual_loss(c, scratch_full, mcres, nullptr, ssq, C, stream);
}

extern "C" int gmc_sgs_step_injected(gmc_ctx* c, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol,
                                     const int32_t* centre, const int32_t* block_size, const int32_t* path, const double* znorm,
                                     int64_t path_stride, const double* u, uint8_t* accepted_out, double* loss_out,
                                     double* loss_next_out, int32_t* resampled, int32_t* err_flag, int C, void* stream) {
    int rc = sgs_check(c, C, "gmc_sgs_step_injected");
    if (rc) return rc;
    if (!bedc || !z || !mcres || !ssq || !nviol || !centre || !block_size || !path || !znorm || !u)
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_step_injected: NULL argument");
    if (c->sgs->warp_solver)
        sgs_step_injected_kernel<true><<<C, SGS_THREADS, sizeof(SgsShared), (cudaStream_t)stream>>>(
            c->dev, c->sgs->dev, bedc, z, mcres, ssq, nviol, centre, block_size, path, znorm, path_stride, u, accepted_out, loss_out,
            loss_next_out, resampled, err_flag);
    else
        sgs_step_injected_kernel<false><<<C, SGS_THREADS, sizeof(SgsShared), (cudaStream_t)stream>>>(
            c->dev, c->sgs->dev, bedc, z, mcres, ssq, nviol, centre, block_size, path, znorm, path_stride, u, accepted_out, loss_out,
            loss_next_out, resampled, err_flag);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_sgs_run(gmc_ctx* c, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol, const uint64_t* seeds,
                           uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache,
                           int64_t cache_stride, int64_t cache_offset, int32_t* resampled, int32_t* err_flag, int C, void* stream) {
    int rc = sgs_check(c, C, "gmc_sgs_run");
    if (rc) return rc;
    if (!bedc || !z || !mcres || !ssq || !nviol || !seeds) GMC_FAIL(GMC_EINVAL, "gmc_sgs_run: NULL argument");
    if (n_steps < 0) GMC_FAIL(GMC_EINVAL, "gmc_sgs_run: negative n_steps");
    if ((loss_cache || step_cache || blocks_cache) && (cache_offset < 0 || cache_offset + n_steps > cache_stride))
        GMC_FAIL(GMC_ESHAPE, "gmc_sgs_run: cache window exceeds stride");
    if (n_steps == 0) return GMC_OK;
    c->sgs->dev.phase = c->d_phase;
    {
        int rc1 = gmc_check_device_error(c, "gmc_sgs_run");
        if (rc1) return rc1;
    }
    // more chains than resident CTAs: dynamic (chunk, chain) items, see sgs_run_kernel
    int per_sm = 0;
    if (c->sgs->warp_solver)
        GMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgs_run_kernel<true>, SGS_THREADS, sizeof(SgsShared)));
    else
        GMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sgs_run_kernel<false>, SGS_THREADS, sizeof(SgsShared)));
    const int slots = std::max(1, per_sm) * c->sm_count;
    int grid = C, chunk = n_steps;
    int* sched = nullptr;
    if (C > slots && n_steps > 1 && !getenv("GMC_STATIC_SCHED")) {
        // launches for disjoint chain ranges may run concurrently on different streams (run_pipelined): each takes the
        // next of GMC_SCHED_SLOTS scheduler areas
        int rc2 = gmc_sched_acquire(c, C, (cudaStream_t)stream, &sched);
        if (rc2) return rc2;
        grid = slots;
        chunk = std::min(256, std::max(2, (n_steps + 31) / 32));   // see gmc_run (step.cu)
    }
    if (c->sgs->warp_solver)
        sgs_run_kernel<true><<<grid, SGS_THREADS, sizeof(SgsShared), (cudaStream_t)stream>>>(
            c->dev, c->sgs->dev, bedc, z, mcres, ssq, nviol, seeds, iter0, n_steps, loss_cache, step_cache, blocks_cache, cache_stride,
            cache_offset, resampled, err_flag, C, sched, chunk, c->d_err, c->spin_limit);
    else
        sgs_run_kernel<false><<<grid, SGS_THREADS, sizeof(SgsShared), (cudaStream_t)stream>>>(
            c->dev, c->sgs->dev, bedc, z, mcres, ssq, nviol, seeds, iter0, n_steps, loss_cache, step_cache, blocks_cache, cache_stride,
            cache_offset, resampled, err_flag, C, sched, chunk, c->d_err, c->spin_limit);
    if (sched) gmc_sched_release(c, (cudaStream_t)stream);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

// =====================================================================================================================
// Whole-grid Sequential Gaussian Simulation: gstatsim_custom/interpolate.py:92-191, the generator of
// the large-scale chains' initial beds (one realisation per chain, ~16 min each in the reference).
//
// A node's neighbour set and kriging weights depend on WHICH cells are conditioned when its turn comes - known from the
// path alone - not on the simulated values.  So the simulation splits into
//   1. sgs_grid_solve_kernel: all nodes of all realisations in parallel, one warp per node: octant search in the ORDER
//      grid (ord[cell] = -1 for conditioning data, else the cell's position in the path; a cell is available to node t
//      iff ord < t), kriging solve (sgs_warp_solve), record (neighbour cells, weights, sd) to global memory;
//   2. sgs_grid_values_kernel: thirty-two warps per realisation walk its path (see the kernel): value = est + sd * noise with
//      est = mean + sum w_i (v_i - mean) over the recorded neighbours (a sparse triangular solve), the truncated-normal
//      draw when bounds are given (interpolate.py:166-181), writing the normal-score grid in place.
// =====================================================================================================================
struct SgsGridShared {
    SgsWarpRec rec[8];
    int cell[8][SGS_WN];
    short2 near_off[8][SGS_NEAR];
    int oct_cnt[SGS_MAX_LEVELS][8];
};

__global__ void __launch_bounds__(SGS_THREADS, 1)
    sgs_grid_solve_kernel(SgsDev s, int H, int W, const int32_t* __restrict__ ord_all, const int32_t* __restrict__ path_all,
                          int64_t n_path, int n_real, int n_levels, int32_t* __restrict__ rec_n, int32_t* __restrict__ rec_idx,
                          double* __restrict__ rec_w, double* __restrict__ rec_sd, int32_t* err_out) {
    extern __shared__ __align__(16) unsigned char sgs_raw[];
    SgsGridShared& S = *reinterpret_cast<SgsGridShared*>(sgs_raw);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int t = tid; t < 8 * SGS_NEAR; t += SGS_THREADS) {
        const int o = t / SGS_NEAR, k = t - o * SGS_NEAR;
        short2 v = make_short2(0, 0);
        if (k < s.lmax) v = make_short2(s.oct_off[((int64_t)o * s.lmax + k) * 2], s.oct_off[((int64_t)o * s.lmax + k) * 2 + 1]);
        S.near_off[o][k] = v;
    }
    if (tid < 8 * n_levels) S.oct_cnt[tid >> 3][tid & 7] = s.oct_cnt[tid];      // prefix lengths of the lists per radius level
    __syncthreads();
    SgsWarpRec& R = S.rec[wid];
    int* cellv = S.cell[wid];
    const int64_t plane = (int64_t)H * W, total = (int64_t)n_real * n_path;
    for (int64_t item = (int64_t)blockIdx.x * 8 + wid; item < total; item += (int64_t)gridDim.x * 8) {
        const int r = (int)(item / n_path);
        const int64_t t = item - (int64_t)r * n_path;
        const int32_t* ord = ord_all + r * plane;
        const int cell0 = path_all[r * n_path + t];
        if (__ldg(ord + cell0) != (int32_t)t) {             // conditioning cell: nothing to simulate
            if (lane == 0) rec_n[item] = -1;
            continue;
        }
        const int i = cell0 / W, j = cell0 - i * W;
        int n = 0;
        // the lists are sorted by distance, so "all offsets closer than the level's radius" is a prefix; a node that finds
        // nothing within `radius` searches again with radius + 100 km, as the reference does (interpolate.py:149-155)
        for (int lev = 0; lev < n_levels && n == 0; ++lev)
        for (int o = 0; o < 8; ++o) {                        // neighbors.py:52-60
            const int16_t* off = s.oct_off + (int64_t)o * s.lmax * 2;
            const int cnt = S.oct_cnt[lev][o];
            int found = 0;
            for (int base = 0; base < cnt && found < s.per_oct; base += 32) {
                const int q = base + lane;
                bool ok = false;
                int di = 0, dj = 0, cell = 0;
                if (q < cnt) {
                    if (q < SGS_NEAR) {
                        const short2 o2 = S.near_off[o][q];
                        di = o2.x;
                        dj = o2.y;
                    } else {
                        di = off[2 * q];
                        dj = off[2 * q + 1];
                    }
                    const int ci = i + di, cj = j + dj;
                    if (ci >= 0 && ci < H && cj >= 0 && cj < W) {
                        cell = ci * W + cj;
                        ok = __ldg(ord + cell) < (int32_t)t;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                const int rank = found + __popc(m & ((1u << lane) - 1u));
                if (ok && rank < s.per_oct) {
                    const int slot = n + rank;
                    R.ndi[slot] = (int16_t)di;
                    R.ndj[slot] = (int16_t)dj;
                    cellv[slot] = cell;
                }
                found += __popc(m);
            }
            n += min(found, s.per_oct);
        }
        __syncwarp();
        if (n == 0) {                                        // nothing even within the widest radius the caller provided
            if (lane == 0) {
                rec_n[item] = 0;
                rec_sd[item] = 0.0;
                atomicOr(err_out, 1);
            }
            continue;
        }
        sgs_warp_solve(s, R, n);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int q = lane + 32 * h2;
            if (q < n) {
                rec_idx[item * SGS_WN + q] = cellv[q];
                rec_w[item * SGS_WN + q] = R.w[q];
            }
        }
        if (lane == 0) {
            rec_n[item] = n;
            rec_sd[item] = sqrt(R.var);                      // interpolate.py:163 var = abs(var)
        }
        __syncwarp();
    }
}

// inverse cdf of the standard normal restricted to [a, b] at probability u (what scipy's truncnorm.rvs applies to its one
// uniform draw), through the survival function when both bounds lie in the upper tail
__device__ __forceinline__ double truncnorm_ppf(double u, double a, double b) {
    if (a > 0.0) {
        const double sa = normcdf(-a), sb = normcdf(-b);
        return -normcdfinv(sa - u * (sa - sb));
    }
    const double pa = normcdf(a), pb = normcdf(b);
    return normcdfinv(pa + u * (pb - pa));
}

// One CTA of SGV_WARPS warps per realisation; warp w takes the path nodes w, w + SGV_WARPS, ... in order.  A node's value
// needs the values of its recorded neighbours, all EARLIER in the path: cells still to be simulated hold NaN in z, so a
// lane simply re-reads its neighbour (through L2) until it is a number.  The warp working on the earliest unfinished node
// never waits, so the walk cannot deadlock; neighbours are rarely among the last SGV_WARPS path positions, so the warps
// mostly run independently and the dependent chain of the one-warp-per-realisation walk (2.2 us per node: record, gather,
// reduce, store) is overlapped SGV_WARPS-fold.  Same arithmetic in the same order per node: results are bit-identical to
// the sequential walk.  (A neighbour that is genuinely NaN - a NaN weight or bound upstream - stops the waiting after a
// bound and propagates, as it would sequentially.)
#define SGV_WARPS 32
__device__ __forceinline__ double ld_cg_f64(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(SGV_WARPS * 32)
    sgs_grid_values_kernel(int H, int W, double* __restrict__ z_all, const int32_t* __restrict__ path_all, int64_t n_path,
                           const int32_t* __restrict__ rec_n, const int32_t* __restrict__ rec_idx,
                           const double* __restrict__ rec_w, const double* __restrict__ rec_sd,
                           const double* __restrict__ noise_all, const double* __restrict__ blo, const double* __restrict__ bhi) {
    __shared__ int poisoned;                                 // some warp met a neighbour that never became a number
    const int r = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t plane = (int64_t)H * W;
    double* z = z_all + r * plane;
    const int32_t* path = path_all + r * n_path;
    const int64_t base = (int64_t)r * n_path;
    if (threadIdx.x == 0) poisoned = 0;
    __syncthreads();
    for (int64_t t = wid; t < n_path; t += SGV_WARPS) {
        const int n = rec_n[base + t];
        if (n < 0) continue;                                 // conditioning cell (warp-uniform)
        const int cell = path[t];
        int idx[2] = {0, 0};
        double w[2] = {0.0, 0.0}, v[2] = {0.0, 0.0};
        bool need[2];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int q = lane + 32 * h2;
            need[h2] = q < n;
            if (need[h2]) {
                idx[h2] = __ldg(rec_idx + (base + t) * SGS_WN + q);
                w[h2] = __ldg(rec_w + (base + t) * SGS_WN + q);
            }
        }
        const double sd = rec_sd[base + t], nz = noise_all[base + t];      // independent of the values: in flight early
        unsigned spins = 0;
        for (;;) {
            bool pending = false;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2)
                if (need[h2]) {
                    v[h2] = ld_cg_f64(z + idx[h2]);
                    if (v[h2] == v[h2]) need[h2] = false;
                    else pending = true;
                }
            if (!__any_sync(0xffffffffu, pending)) break;
            if (*(volatile int*)&poisoned || ++spins > (1u << 18)) {      // never hang: carry the NaN like the sequential walk
                poisoned = 1;
                break;
            }
        }
        double sv = 0.0, swv = 0.0, sw = 0.0;
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int q = lane + 32 * h2;
            if (q < n) {
                sv += v[h2];
                swv += w[h2] * v[h2];
                sw += w[h2];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sv += __shfl_down_sync(0xffffffffu, sv, o);
            swv += __shfl_down_sync(0xffffffffu, swv, o);
            sw += __shfl_down_sync(0xffffffffu, sw, o);
        }
        if (lane == 0) {
            double val = 0.0;
            if (n > 0) {
                const double mean = sv / (double)n;
                const double est = mean + (swv - mean * sw);             // = mean + sum w (v - mean)     _krige.py:42
                if (!blo) val = est + sd * nz;                           // rng.normal(est, sd)           :173
                else {
                    const double lo = blo[cell], hi = bhi[cell];
                    if (lo == hi) val = lo;                              // :177-178
                    else val = est + sd * truncnorm_ppf(nz, (lo - est) / sd, (hi - est) / sd);     // :180-181
                }
            }
            __stcg(z + cell, val);
            __threadfence();                                 // visible to the warps (of this CTA, on this SM or through L2) that wait for it
        }
        __syncwarp();
    }
}

// standalone normal-score transform (QuantileTransformer tables on the device): interpolate.py:185, utilities.py:21-24
extern "C" int gmc_nst_transform(int device, const double* quantiles, const double* references, int n_quantiles,
                                 const double* in, double* out, int64_t n, int inverse, void* stream) {
    if (!quantiles || !references || !in || !out) GMC_FAIL(GMC_EINVAL, "gmc_nst_transform: NULL argument");
    if (n_quantiles < 2 || n < 1) GMC_FAIL(GMC_EINVAL, "gmc_nst_transform: need >= 2 quantiles and >= 1 value");
    GMC_CUDA(cudaSetDevice(device));
    SgsDev s;
    memset(&s, 0, sizeof(s));
    s.quant = quantiles;
    s.refs = references;
    s.nq = n_quantiles;
    sgs_transform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(s, in, out, n, inverse);
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_sgs_grid_solve(int device, int H, int W, const int32_t* ord, const int32_t* path, int64_t n_path, int n_real,
                                  const int16_t* oct_off, const int32_t* oct_cnt, int n_levels, int lmax, int hw, int num_points,
                                  const double* lut, double sill, int32_t* rec_n, int32_t* rec_idx, double* rec_w,
                                  double* rec_sd, int32_t* err_flag, void* stream) {
    if (!ord || !path || !oct_off || !oct_cnt || !lut || !rec_n || !rec_idx || !rec_w || !rec_sd || !err_flag)
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_grid_solve: NULL argument");
    if (H < 1 || W < 1 || n_path < 1 || n_path > (int64_t)H * W || n_real < 1) GMC_FAIL(GMC_ESHAPE, "gmc_sgs_grid_solve: bad sizes");
    if (num_points < 8 || num_points > SGS_WN)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_sgs_grid_solve: num_points=%d outside [8,%d]", num_points, SGS_WN);
    if (hw < 1 || lmax < 1) GMC_FAIL(GMC_EINVAL, "gmc_sgs_grid_solve: empty search stencil");
    if (n_levels < 1 || n_levels > SGS_MAX_LEVELS)
        GMC_FAIL(GMC_EINVAL, "gmc_sgs_grid_solve: n_levels=%d outside [1,%d]", n_levels, SGS_MAX_LEVELS);
    GMC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GMC_CUDA(cudaGetDeviceProperties(&prop, device));
    SgsDev s;
    memset(&s, 0, sizeof(s));
    s.oct_off = oct_off;
    s.oct_cnt = oct_cnt;
    s.lmax = lmax;
    s.hw = hw;
    s.per_oct = num_points / 8;
    s.lut = lut;
    s.lut_w = 4 * hw + 1;
    s.sill = sill;
    GMC_CUDA(cudaFuncSetAttribute(sgs_grid_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SgsGridShared)));
    const int64_t items = (int64_t)n_real * n_path;
    const int ctas = (int)std::min<int64_t>((items + 7) / 8, (int64_t)prop.multiProcessorCount * 8);
    sgs_grid_solve_kernel<<<ctas, SGS_THREADS, sizeof(SgsGridShared), (cudaStream_t)stream>>>(
        s, H, W, ord, path, n_path, n_real, n_levels, rec_n, rec_idx, rec_w, rec_sd, err_flag);
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

// cells to simulate get the "not yet simulated" marker (NaN) whatever the caller left there
__global__ void sgs_grid_mark_kernel(int H, int W, double* __restrict__ z_all, const int32_t* __restrict__ path_all, int64_t n_path,
                                     const int32_t* __restrict__ rec_n, int64_t total) {
    const int64_t item = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (item >= total) return;
    if (rec_n[item] >= 0) {
        const int64_t r = item / n_path;
        z_all[r * (int64_t)H * W + path_all[item]] = __longlong_as_double(0x7ff8000000000000LL);
    }
}

extern "C" int gmc_sgs_grid_values(int device, int H, int W, double* z, const int32_t* path, int64_t n_path, int n_real,
                                   const int32_t* rec_n, const int32_t* rec_idx, const double* rec_w, const double* rec_sd,
                                   const double* noise, const double* bound_lo, const double* bound_hi, void* stream) {
    if (!z || !path || !rec_n || !rec_idx || !rec_w || !rec_sd || !noise) GMC_FAIL(GMC_EINVAL, "gmc_sgs_grid_values: NULL argument");
    if ((bound_lo == nullptr) != (bound_hi == nullptr)) GMC_FAIL(GMC_EINVAL, "gmc_sgs_grid_values: give both bounds or neither");
    if (H < 1 || W < 1 || n_path < 1 || n_real < 1) GMC_FAIL(GMC_ESHAPE, "gmc_sgs_grid_values: bad sizes");
    GMC_CUDA(cudaSetDevice(device));
    const int64_t total = (int64_t)n_real * n_path;
    sgs_grid_mark_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(H, W, z, path, n_path, rec_n, total);
    sgs_grid_values_kernel<<<n_real, SGV_WARPS * 32, 0, (cudaStream_t)stream>>>(H, W, z, path, n_path, rec_n, rec_idx, rec_w, rec_sd, noise,
                                                                    bound_lo, bound_hi);
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}
