// residual.cu — K2/K3: batched full-grid mass-conservation residual and masked loss.
//
// Reference semantics (Topography.py:592-600, MCMC.py:1041):
//   thick = surf - bed;  dx = np.gradient(velx*thick, res, axis=1);  dy = np.gradient(vely*thick, res, axis=0)
//   out = ((dx + dy) + dhdt) - smb;   loss = nansum(out[mask==1]^2) / (2 sigma^2)
// np.gradient (uniform spacing, edge_order=1): interior (f[i+1]-f[i-1])/(2.*h), first/last (f[1]-f[0])/h.
// All arithmetic is rounded per operation (no FMA contraction, true division) so the result is bit-identical
// to numpy.  The loss is summed in a fixed order (deterministic), which differs from numpy's pairwise order only
// by rounding (<= 1e-12 relative; the contract is 1e-9).
#include "common.cuh"

#define RES_TX 32
#define RES_TY 8
#define RES_ROWS_PER_THREAD 4   // each CTA covers RES_TY*RES_ROWS_PER_THREAD rows x RES_TX*2 columns
#define RES_TILE_H (RES_TY * RES_ROWS_PER_THREAD)
#define RES_TILE_W (RES_TX * 2)

__device__ __forceinline__ double cell_fx(const GmcDev& d, const double* __restrict__ bed, int64_t idx) {
    return mul_rn(__ldg(d.velx + idx), sub_rn(__ldg(d.surf + idx), bed[idx]));
}
__device__ __forceinline__ double cell_fy(const GmcDev& d, const double* __restrict__ bed, int64_t idx) {
    return mul_rn(__ldg(d.vely + idx), sub_rn(__ldg(d.surf + idx), bed[idx]));
}

// residual of one cell from global memory (used by the v1 kernel; the step kernel has its own smem version)
__device__ __forceinline__ double cell_residual(const GmcDev& d, const double* __restrict__ bed, int i, int j) {
    const int H = d.H, W = d.W;
    const int64_t row = (int64_t)i * W;
    double dx, dy;
    if (j == 0)
        dx = div_rn(sub_rn(cell_fx(d, bed, row + 1), cell_fx(d, bed, row)), d.res);
    else if (j == W - 1)
        dx = div_rn(sub_rn(cell_fx(d, bed, row + W - 1), cell_fx(d, bed, row + W - 2)), d.res);
    else
        dx = div_rn(sub_rn(cell_fx(d, bed, row + j + 1), cell_fx(d, bed, row + j - 1)), d.two_res);
    if (i == 0)
        dy = div_rn(sub_rn(cell_fy(d, bed, (int64_t)W + j), cell_fy(d, bed, j)), d.res);
    else if (i == H - 1)
        dy = div_rn(sub_rn(cell_fy(d, bed, row + j), cell_fy(d, bed, row - W + j)), d.res);
    else
        dy = div_rn(sub_rn(cell_fy(d, bed, row + W + j), cell_fy(d, bed, row - W + j)), d.two_res);
    return sub_rn(add_rn(add_rn(dx, dy), __ldg(d.dhdt + row + j)), __ldg(d.smb + row + j));
}

template <bool WRITE_RES, bool DO_LOSS>
__global__ void __launch_bounds__(RES_TX* RES_TY)
    residual_kernel(GmcDev d, const double* __restrict__ bed_all, double* __restrict__ res_all,
                    double* __restrict__ partials, int n_tiles) {
    __shared__ double scratch[33];
    const int c = blockIdx.z;
    const int64_t plane = (int64_t)d.H * d.W;
    const double* bed = bed_all + c * plane;
    const int j0 = (blockIdx.x * RES_TX + threadIdx.x) * 2;
    const int i0 = blockIdx.y * RES_TILE_H + threadIdx.y * RES_ROWS_PER_THREAD;
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < RES_ROWS_PER_THREAD; ++r) {
        const int i = i0 + r;
        if (i >= d.H) break;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int j = j0 + q;
            if (j >= d.W) break;
            const double v = cell_residual(d, bed, i, j);
            const int64_t idx = (int64_t)i * d.W + j;
            if (WRITE_RES) res_all[c * plane + idx] = v;
            if (DO_LOSS) {
                if ((__ldg(d.flags + idx) & FLAG_MC) && v == v) acc = add_rn(acc, mul_rn(v, v));
            }
        }
    }
    if (DO_LOSS) {
        const int tid = threadIdx.y * RES_TX + threadIdx.x;
        // block_sum uses threadIdx.x only; flatten
        double v = warp_sum(acc);
        if ((tid & 31) == 0) scratch[tid >> 5] = v;
        __syncthreads();
        if (tid < 32) {
            double t = (tid < (RES_TX * RES_TY) / 32) ? scratch[tid] : 0.0;
            t = warp_sum(t);
            if (tid == 0) partials[(int64_t)c * n_tiles + blockIdx.y * gridDim.x + blockIdx.x] = t;
        }
    }
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
        const double v = res[k];
        if ((__ldg(d.flags + k) & FLAG_MC) && v == v) acc = add_rn(acc, mul_rn(v, v));
    }
    const double t = block_sum<LOSS_THREADS>(acc, scratch);
    if (threadIdx.x == 0) partials[(int64_t)c * n_tiles + blockIdx.x] = t;
}

// fixed-order sum of the per-tile partials: one warp per chain
__global__ void finalize_loss_kernel(const double* __restrict__ partials, int n_tiles, double two_sigma2,
                                     double* __restrict__ loss_out, double* __restrict__ ssq_out, int C) {
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int k = lane; k < n_tiles; k += 32) acc += partials[(int64_t)c * n_tiles + k];
    acc = warp_sum(acc);
    if (lane == 0) {
        if (ssq_out) ssq_out[c] = acc;
        if (loss_out) loss_out[c] = div_rn(acc, two_sigma2);
    }
}

static int ensure_partials(gmc_ctx* c) {
    const int tiles = ((c->W + RES_TILE_W - 1) / RES_TILE_W) * ((c->H + RES_TILE_H - 1) / RES_TILE_H);
    if (!c->d_partials || c->n_tiles != tiles) {
        cudaFree(c->d_partials);
        c->d_partials = nullptr;
        GMC_CUDA(cudaMalloc(&c->d_partials, (size_t)c->max_chains * tiles * sizeof(double)));
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

static int launch_residual(gmc_ctx* c, const double* bed, double* res_out, double* loss_out, double* ssq_out, int C,
                           bool do_loss, cudaStream_t st) {
    int rc = ensure_partials(c);
    if (rc) return rc;
    const dim3 block(RES_TX, RES_TY);
    const int tx = (c->W + RES_TILE_W - 1) / RES_TILE_W, ty = (c->H + RES_TILE_H - 1) / RES_TILE_H;
    for (int c0 = 0; c0 < C; c0 += 65535) {
        const int cn = (C - c0 < 65535) ? C - c0 : 65535;
        const dim3 grid(tx, ty, cn);
        const double* b = bed + (size_t)c0 * c->H * c->W;
        double* r = res_out ? res_out + (size_t)c0 * c->H * c->W : nullptr;
        double* p = c->d_partials + (size_t)c0 * c->n_tiles;
        if (res_out && do_loss)
            residual_kernel<true, true><<<grid, block, 0, st>>>(c->dev, b, r, p, c->n_tiles);
        else if (res_out)
            residual_kernel<true, false><<<grid, block, 0, st>>>(c->dev, b, r, p, c->n_tiles);
        else
            residual_kernel<false, true><<<grid, block, 0, st>>>(c->dev, b, r, p, c->n_tiles);
        c->launches++;
    }
    if (do_loss) {
        finalize_loss_kernel<<<(C + 7) / 8, 256, 0, st>>>(c->d_partials, c->n_tiles, c->dev.two_sigma2, loss_out, ssq_out, C);
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

extern "C" int gmc_loss(gmc_ctx* c, const double* res, double* loss_out, double* ssq_out, int C, void* stream) {
    int rc = check_common(c, res, C, "gmc_loss");
    if (rc) return rc;
    if (!loss_out && !ssq_out) GMC_FAIL(GMC_EINVAL, "gmc_loss: loss_out and ssq_out are both NULL");
    rc = ensure_partials(c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    for (int c0 = 0; c0 < C; c0 += 65535) {
        const int cn = (C - c0 < 65535) ? C - c0 : 65535;
        loss_kernel<<<dim3(c->n_tiles, cn), LOSS_THREADS, 0, st>>>(c->dev, res + (size_t)c0 * c->H * c->W,
                                                                  c->d_partials + (size_t)c0 * c->n_tiles, c->n_tiles);
        c->launches++;
    }
    finalize_loss_kernel<<<(C + 7) / 8, 256, 0, st>>>(c->d_partials, c->n_tiles, c->dev.two_sigma2, loss_out, ssq_out, C);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}
