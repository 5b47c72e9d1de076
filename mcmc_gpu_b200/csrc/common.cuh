// common.cuh — context layout, error plumbing and small device helpers shared by the libgmc translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gmc.h"

#define GMC_MAX_RADIX 64       // largest prime factor of a block edge handled by the smem FFT (generic stage)
#define GMC_MAX_FACTORS 12
#ifndef GMC_STEP_THREADS
#define GMC_STEP_THREADS 256
#endif
#ifndef GMC_STEP_MIN_CTAS
#define GMC_STEP_MIN_CTAS 2
#endif
#define GMC_N_PHASES 8

// ---------------------------------------------------------------------------------------------------------------
// host: errors
// ---------------------------------------------------------------------------------------------------------------
void gmc_set_error(const char* fmt, ...);

#define GMC_FAIL(code, ...)         \
    do {                            \
        gmc_set_error(__VA_ARGS__); \
        return (code);              \
    } while (0)

#define GMC_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            gmc_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return GMC_ECUDA;                                                                       \
        }                                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------
// FFT plan of one edge length n (mixed radix, in-place decimation in time, digit-reversed load)
// ---------------------------------------------------------------------------------------------------------------
struct GmcFftPlan {
    int n;
    int n_factors;
    int radix[GMC_MAX_FACTORS];   // stage order: radix[0] is the innermost (first executed) stage
    int tw_off;                   // offset (in double2) of exp(+2 pi i k / n), k in [0,n), in ctx->d_twiddle
    int htw_off;                  // offset (in double2) of exp(+2 pi i k / (2n)), k in [0,n): real-row recombination
    int perm_off;                 // offset (in int16) into ctx->d_perm: position p holds logical index perm[p]
    int pos_off;                  // offset into ctx->d_pos: logical index k is stored at position pos[k]
};

// per block-size pair
struct GmcPair {
    int h, w;            // field shape [h][w] (rows, cols)
    int pitchc;          // row pitch of the half plane in double2 units (>= w/2+1, odd)
    int ksq_off_h, ksq_off_w;  // offsets (in double) of (2 pi fftfreq(n, d))^2, k in [0, n/2], for n = h and n = w
    int64_t mask_off;    // offset (in double) of the taper in ctx->d_edge_masks
    GmcFftPlan ph, pw;   // column transform of length h, row transform of length w/2 (copied in: one load gets all)
};

#define GMC_MAX_EDGE 128   // largest block edge (bounded anyway by one SM's shared memory)

struct GmcFieldModel {
    int model;
    int isotropic;
    double smoothness;
    double range_min_x, range_max_x, range_min_y, range_max_y;
    double scale_min, scale_max, nugget_max;
    double matern_num;    // 4 pi Gamma(nu+1) (2 nu)^nu                     MCMC.py:236
    double matern_gamma;  // Gamma(nu)
};

// everything a kernel needs, passed by value
struct GmcDev {
    int H, W;
    const double* surf;
    const double* velx;
    const double* vely;
    const double* dhdt;
    const double* smb;
    const double* crf_weight;   // NULL for block_type 'RF'
    const double2* sv;          // packed {surf, velx} per cell: one 16 B load feeds an x-flux
    const double2* sy;          // packed {surf, vely} per cell: one 16 B load feeds a y-flux
    const double2* ds;          // packed {dhdt, smb} per cell
    const uint8_t* flags;       // bit0 gate, bit1 mc
    const int32_t* centre_cells;
    int64_t n_centre_cells;
    double res;                 // chain.resolution
    double two_res;             // 2.*res
    double two_sigma2;          // 2*sigma_mc**2
    double r_res, r_two_res;    // RN(1/res), RN(1/(2 res)) for div_const
    // block table
    int n_pairs;
    const GmcPair* pairs;
    const double2* twiddle;
    const int16_t* perm;
    const int16_t* pos;
    const double* ksq;
    const double* edge_masks;
    GmcFieldModel fm;
};

struct gmc_sgs_state;
void gmc_sgs_destroy(struct gmc_ctx* c);

struct gmc_ctx {
    int device;
    int H, W, max_chains;
    int sm_count;
    bool have_static, have_model, have_blocks;
    GmcDev dev;
    // owned device memory
    double* d_static;      // 5 or 6 planes of H*W
    uint8_t* d_flags;
    int32_t* d_centre;
    double* d_partials;    // [max_chains][n_tiles] loss partials
    int n_tiles;
    GmcPair* d_pairs;
    double2* d_twiddle;
    int16_t* d_perm;
    int16_t* d_pos;
    double* d_ksq;
    double* d_edge_masks;
    std::vector<GmcPair> h_pairs;
    std::vector<GmcFftPlan> h_plans;
    int max_h, max_w;
    int step_smem_bytes;
    int step_tile_off;     // offset (in doubles) of the candidate tile inside the step kernel's dynamic shared memory
    int step_ctas_per_sm;
    int spectral;          // RandField.set_generation_method: 1 = FFT synthesis (A3), 0 = randomization method (A5)
    int n_modes;           // wave vectors per randomization-method field (gstools mode_no, default 1000)
    int rm_smem_bytes;     // dynamic shared memory / tile offset of the randomization-method kernels
    int rm_tile_off;
    double field_res;      // grid spacing of the proposal fields (gmc_set_blocks)
    int64_t launches;
    long long* d_phase;    // optional per-phase cycle counters of run_kernel (debug)
    int* d_sched;          // GMC_SCHED_SLOTS areas of [work counter, completed chunks per chain] (launches with C > resident CTAs)
    unsigned sched_next;
    cudaEvent_t sched_ev[16];   // recorded behind the launch that used the area: the next user waits for it on ITS stream
    bool sched_used[16];
    int sched_cur;         // area handed out by the last gmc_sched_acquire
    int* h_err;            // pinned + mapped: a kernel that gives up a bounded wait stores a GMC_DEVERR_* code here
    int* d_err;            // device alias of h_err
    unsigned spin_limit;   // bound of the in-kernel waits (GMC_DEBUG_SPIN_LIMIT overrides the default 2^26)
    int step_wide_ctas;    // occupancy of the 512-thread step kernel (0 = unavailable)
    double* d_ring;        // split mode: ring of proposal fields [C][depth][max_h * max_w] written by the producer CTAs
    size_t ring_bytes;
    int* d_pflags;         // split mode: [max_chains][2] fields produced / consumed in the current launch
    cudaStream_t aux_stream;   // split mode: the producer kernel runs here, fenced to the caller's stream by two events
    cudaEvent_t ev_fork, ev_join;
    int step_cta_mode;     // gmc_set_step_cta: 0 auto (512 threads when C <= SMs), 1 always 256 threads, 2 512 threads when it fits
    gmc_sgs_state* sgs;    // small-scale (SGS) chain tables, see sgs.cu
};

#define GMC_SCHED_SLOTS 16
#define GMC_DEVERR_WAIT_TIMEOUT 1   // a chunk gave up waiting for its chain's previous chunk (or a tile copy never completed)

// scheduler areas for launches with more chains than resident CTAs (ctx.cu): acquire zeroes an area on `st` after making the
// stream wait for the area's previous user; release records that user's completion.  check_device_error reads the mapped
// flag without synchronising (it reports what earlier, already finished launches stored).
int gmc_sched_acquire(struct gmc_ctx* c, int C, cudaStream_t st, int** sched_out);
void gmc_sched_release(struct gmc_ctx* c, cudaStream_t st);
int gmc_check_device_error(struct gmc_ctx* c, const char* who);
#define FLAG_GATE 1
#define FLAG_MC 2

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Rounded (never FMA-contracted) arithmetic for the parity-critical expressions.
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

// x / d for a loop-invariant d with r = RN(1/d): q = RN(x r), rem = x - d q (exact in one FMA), q' = RN(q + rem r).
// This is the final correction step of the IEEE division routine; with a correctly rounded reciprocal it returns
// RN(x/d) whenever no intermediate over/underflows, which the exponent guard ensures (else the true division runs).
// tests/test_gpu_residual.py checks it against __ddiv_rn on random and adversarial operands.
static __device__ __noinline__ double div_slow(double x, double d) { return __ddiv_rn(x, d); }
__device__ __forceinline__ double div_const(double x, double d, double r) {
    const double q = __dmul_rn(x, r);
    const int e = __double2hiint(q) & 0x7fffffff;
    if (e >= 0x05d00000 && e <= 0x7a100000) return fma(fma(-d, q, x), r, q);   // |q| in [2^-930, 2^930]
    if (x == 0.0) return q;                     // +-0 / d: q already carries the right sign
    return div_slow(x, d);                      // subnormal, huge, inf, nan: rare, kept out of line
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: any draw is addressable by (key, counter) -------------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ explicit Philox(uint64_t key) : k0((uint32_t)key), k1((uint32_t)(key >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            c0 = hi1 ^ c1 ^ a;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ b;
            c3 = lo0;
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Stream ids (counter word 3).  Counter = (draw index, iteration lo, iteration hi, stream).
enum { GMC_STREAM_RF_SCALARS = 0, GMC_STREAM_NOISE = 1, GMC_STREAM_NUGGET = 2, GMC_STREAM_CHAIN = 3,
       GMC_STREAM_RM_MODE = 8, GMC_STREAM_RM_AMP = 9 };   // 5, 6: small-scale chain (sgs.cu)

// uniform in (0,1), exactly representable: (k + 0.5) * 2^-52 with a 52-bit k
__device__ __forceinline__ double u01_open(uint32_t hi, uint32_t lo) {
    const uint64_t k = ((uint64_t)(hi >> 6) << 26) | (uint64_t)(lo >> 6);
    return ((double)k + 0.5) * 2.220446049250313e-16;
}
// 53-bit uniform in [0,1): k * 2^-53 (numpy Generator.random convention, MCMC.py:1336)
__device__ __forceinline__ double u01_halfopen(uint32_t hi, uint32_t lo) {
    const uint64_t k = ((uint64_t)(hi >> 5) << 26) | (uint64_t)(lo >> 6);
    return (double)k * 1.1102230246251565e-16;
}
// unbiased-to-2^-64 bounded integer in [0, n)
__device__ __forceinline__ uint64_t bounded_u64(uint32_t hi, uint32_t lo, uint64_t n) {
    return __umul64hi(((uint64_t)hi << 32) | lo, n);
}
// two independent N(0,1) from one Philox block
static __device__ __noinline__ void box_muller(const uint4 r, double& z0, double& z1) {
    const double u1 = u01_open(r.x, r.y);
    const double u2 = u01_open(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

// ---- block reductions (fixed order => deterministic) ---------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA; result valid in all threads.  `scratch` holds >= 33 doubles.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();                       // protect scratch from a previous use
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double t = (lane < THREADS / 32) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

#endif  // __CUDACC__
