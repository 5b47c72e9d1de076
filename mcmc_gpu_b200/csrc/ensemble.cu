// ensemble.cu — K5: per-GPU partial moments of the chain ensemble and the one collective on the path.
//
// The reference never forms cross-chain statistics on the fly (visualization.ipynb loads per-chain files); here each
// GPU reduces its own chains to sum (bed - ref) and sum (bed - ref)^2 per cell (HBM-bound: reads C*H*W*8 bytes once)
// and a single fp64 SUM all-reduce of 2*H*W+1 values combines the GPUs.
#include <dlfcn.h>

#include "common.cuh"

// each thread owns two adjacent cells and walks the chain axis: coalesced 16 B loads, fixed summation order
__global__ void __launch_bounds__(256)
    moments_kernel(const double* __restrict__ bed, const double* __restrict__ ref, double* __restrict__ sum_out,
                   double* __restrict__ sumsq_out, int64_t plane, int C) {
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
    if (k >= plane) return;
    if (k + 1 < plane && (plane & 1) == 0) {
        const double2 r = *reinterpret_cast<const double2*>(ref + k);
        double s0 = 0, s1 = 0, q0 = 0, q1 = 0;
#pragma unroll 4
        for (int c = 0; c < C; ++c) {
            const double2 b = __ldcs(reinterpret_cast<const double2*>(bed + (int64_t)c * plane + k));
            const double d0 = b.x - r.x, d1 = b.y - r.y;
            s0 += d0;
            s1 += d1;
            q0 += d0 * d0;
            q1 += d1 * d1;
        }
        *reinterpret_cast<double2*>(sum_out + k) = make_double2(s0, s1);
        *reinterpret_cast<double2*>(sumsq_out + k) = make_double2(q0, q1);
    } else {
        for (int64_t kk = k; kk < plane && kk < k + 2; ++kk) {
            const double r = ref[kk];
            double s = 0, q = 0;
            for (int c = 0; c < C; ++c) {
                const double dd = bed[(int64_t)c * plane + kk] - r;
                s += dd;
                q += dd * dd;
            }
            sum_out[kk] = s;
            sumsq_out[kk] = q;
        }
    }
}

extern "C" int gmc_ensemble_moments(gmc_ctx* c, const double* bed, const double* ref_bed, double* sum_out,
                                    double* sumsq_out, int C, void* stream) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_ensemble_moments: ctx is NULL");
    if (!bed || !ref_bed || !sum_out || !sumsq_out) GMC_FAIL(GMC_EINVAL, "gmc_ensemble_moments: NULL argument");
    if (C < 1 || C > c->max_chains) GMC_FAIL(GMC_ESHAPE, "gmc_ensemble_moments: C=%d outside [1,%d]", C, c->max_chains);
    GMC_CUDA(cudaSetDevice(c->device));
    const int64_t plane = (int64_t)c->H * c->W;
    const int64_t threads = (plane + 1) / 2;
    moments_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(bed, ref_bed, sum_out, sumsq_out, plane, C);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

// ---- NCCL, resolved at run time so libgmc.so carries no link-time dependency on a particular libnccl -----------
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
typedef int (*nccl_group_fn)(void);

static void* nccl_sym(const char* name) {
    void* s = dlsym(RTLD_DEFAULT, name);
    if (s) return s;
    static void* handle = nullptr;
    if (!handle) handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    return handle ? dlsym(handle, name) : nullptr;
}

extern "C" int gmc_allreduce_moments(gmc_ctx* c, void* comm, double* sum, double* sumsq, double* count, void* stream) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_allreduce_moments: ctx is NULL");
    if (!comm || !sum || !sumsq || !count) GMC_FAIL(GMC_EINVAL, "gmc_allreduce_moments: NULL argument");
    nccl_allreduce_fn ar = (nccl_allreduce_fn)nccl_sym("ncclAllReduce");
    nccl_errstr_fn es = (nccl_errstr_fn)nccl_sym("ncclGetErrorString");
    nccl_group_fn gs = (nccl_group_fn)nccl_sym("ncclGroupStart");
    nccl_group_fn ge = (nccl_group_fn)nccl_sym("ncclGroupEnd");
    if (!ar || !es || !gs || !ge) GMC_FAIL(GMC_ENCCL, "gmc_allreduce_moments: libnccl.so.2 not loadable (%s)", dlerror());
    GMC_CUDA(cudaSetDevice(c->device));
    const size_t plane = (size_t)c->H * c->W;
    const int kFloat64 = 8, kSum = 0;   // ncclFloat64, ncclSum (nccl.h)
    int rc = gs();
    if (!rc) rc = ar(sum, sum, plane, kFloat64, kSum, comm, (cudaStream_t)stream);
    if (!rc) rc = ar(sumsq, sumsq, plane, kFloat64, kSum, comm, (cudaStream_t)stream);
    if (!rc) rc = ar(count, count, 1, kFloat64, kSum, comm, (cudaStream_t)stream);
    const int rc2 = ge();
    if (rc || rc2) GMC_FAIL(GMC_ENCCL, "gmc_allreduce_moments: NCCL error: %s", es(rc ? rc : rc2));
    return GMC_OK;
}

// ---- setup helper: exact distance to the nearest masked point ----------------------------------------
// Utilities.min_dist_from_mask (Utilities.py:21-24) queries a KD-tree; the answer is min_p sqrt((x-px)^2 + (y-py)^2) with
// each operation rounded separately (no FMA), which a brute-force scan reproduces bit-for-bit.  Points are tiled through
// shared memory; one thread per query.  O(N*M): meant for the block tapers and conditioning weights (M = radar cells).
__global__ void __launch_bounds__(256)
    min_dist_kernel(const double* __restrict__ px, const double* __restrict__ py, int64_t M, const double* __restrict__ qx,
                    const double* __restrict__ qy, int64_t N, double* __restrict__ out) {
    __shared__ double sx[1024], sy[1024];
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const double x = (q < N) ? qx[q] : 0.0, y = (q < N) ? qy[q] : 0.0;
    double best = __longlong_as_double(0x7ff0000000000000LL);
    for (int64_t base = 0; base < M; base += 1024) {
        const int n = (int)((M - base < 1024) ? M - base : 1024);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            sx[t] = px[base + t];
            sy[t] = py[base + t];
        }
        __syncthreads();
#pragma unroll 4
        for (int t = 0; t < n; ++t) {
            const double dx = __dsub_rn(x, sx[t]), dy = __dsub_rn(y, sy[t]);
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            best = fmin(best, d2);
        }
    }
    if (q < N) out[q] = sqrt(best);
}

extern "C" int gmc_min_dist(int device, const double* px, const double* py, int64_t M, const double* qx, const double* qy,
                            int64_t N, double* out, void* stream) {
    if (!px || !py || !qx || !qy || !out) GMC_FAIL(GMC_EINVAL, "gmc_min_dist: NULL argument");
    if (M < 1 || N < 1) GMC_FAIL(GMC_EINVAL, "gmc_min_dist: empty point or query set");
    GMC_CUDA(cudaSetDevice(device));
    min_dist_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(px, py, M, qx, qy, N, out);
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

// ---- setup helper: the mode filter of Topography.get_highvel_boundary --------------------------------
// The reference smooths the binary region mask with PIL's ImageFilter.ModeFilter(size) (Topography.py:551-553).  PIL's
// rule (ModeFilter.c): histogram of the (2*(size/2)+1)^2 window clipped to the image, most frequent value wins, the lower
// value on ties, and the pixel is kept when no value occurs more than twice.  For the binary {0, 255} image the reference
// builds, the histogram is one count.  One thread per pixel; the image is a few hundred KB and stays in L2.
__global__ void __launch_bounds__(256)
    mode_filter_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int H, int W, int r) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= W || y >= H) return;
    const int y0 = max(0, y - r), y1 = min(H - 1, y + r), x0 = max(0, x - r), x1 = min(W - 1, x + r);
    int c1 = 0;
    for (int yy = y0; yy <= y1; ++yy) {
        const uint8_t* row = in + (int64_t)yy * W;
        for (int xx = x0; xx <= x1; ++xx) c1 += (row[xx] != 0);
    }
    const int n = (y1 - y0 + 1) * (x1 - x0 + 1), c0 = n - c1;
    uint8_t v = in[(int64_t)y * W + x];
    if (max(c0, c1) > 2) v = (c1 > c0) ? 255 : 0;
    out[(int64_t)y * W + x] = v;
}

extern "C" int gmc_mode_filter_binary(int device, const uint8_t* in, uint8_t* out, int H, int W, int size, void* stream) {
    if (!in || !out) GMC_FAIL(GMC_EINVAL, "gmc_mode_filter_binary: NULL argument");
    if (H < 1 || W < 1 || size < 1) GMC_FAIL(GMC_EINVAL, "gmc_mode_filter_binary: H, W and size must be >= 1");
    if (in == out) GMC_FAIL(GMC_EINVAL, "gmc_mode_filter_binary: in-place filtering is not supported");
    GMC_CUDA(cudaSetDevice(device));
    const dim3 grid((W + 31) / 32, (H + 7) / 8);
    mode_filter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, H, W, size / 2);
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}
