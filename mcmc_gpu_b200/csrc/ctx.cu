// ctx.cu — context lifetime and the setup calls of libgmc (host code only; no kernels here).
#include <math.h>
#include <stdarg.h>

#include <algorithm>

#include "common.cuh"
#include <cstdlib>

static thread_local char g_err[512] = "";

void gmc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* gmc_last_error(void) { return g_err; }
extern "C" int gmc_version(void) { return 100; }

int gmc_step_configure(gmc_ctx* ctx);   // step.cu: sizes dynamic smem for the current block table

extern "C" int gmc_create(gmc_ctx** out, int device, int H, int W, int max_chains) {
    if (!out) GMC_FAIL(GMC_EINVAL, "gmc_create: out is NULL");
    *out = nullptr;
    if (H < 2 || W < 2) GMC_FAIL(GMC_ESHAPE, "gmc_create: grid %dx%d too small (np.gradient needs >= 2 per axis)", H, W);
    if (max_chains < 1) GMC_FAIL(GMC_EINVAL, "gmc_create: max_chains must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        GMC_FAIL(GMC_ECUDA, "gmc_create: no CUDA device available (%s); libgmc has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) GMC_FAIL(GMC_EINVAL, "gmc_create: device %d out of range [0,%d)", device, ndev);
    GMC_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GMC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        GMC_FAIL(GMC_ECUDA, "gmc_create: device %d is sm_%d%d; libgmc is built for sm_100a (B200) only", device, prop.major,
                 prop.minor);
    gmc_ctx* c = new gmc_ctx();
    memset(&c->dev, 0, sizeof(c->dev));
    c->device = device;
    c->H = H;
    c->W = W;
    c->max_chains = max_chains;
    c->sm_count = prop.multiProcessorCount;
    c->have_static = c->have_model = c->have_blocks = false;
    c->d_static = nullptr;
    c->d_flags = nullptr;
    c->d_centre = nullptr;
    c->d_partials = nullptr;
    c->d_pairs = nullptr;
    c->d_twiddle = nullptr;
    c->d_perm = c->d_pos = nullptr;
    c->d_ksq = nullptr;
    c->d_edge_masks = nullptr;
    c->max_h = c->max_w = 0;
    c->step_smem_bytes = 0;
    c->step_tile_off = 0;
    c->step_ctas_per_sm = 0;
    c->spectral = 1;
    c->n_modes = 1000;
    c->rm_smem_bytes = c->rm_tile_off = 0;
    c->field_res = 0.0;
    c->launches = 0;
    c->d_phase = nullptr;
    c->d_sched = nullptr;
    c->sched_next = 0;
    c->sched_cur = -1;
    for (int k = 0; k < GMC_SCHED_SLOTS; ++k) {
        c->sched_ev[k] = nullptr;
        c->sched_used[k] = false;
    }
    c->h_err = c->d_err = nullptr;
    c->step_wide_ctas = 0;
    c->step_cta_mode = 0;
    c->d_ring = nullptr;
    c->ring_bytes = 0;
    c->d_pflags = nullptr;
    c->aux_stream = nullptr;
    c->ev_fork = c->ev_join = nullptr;
    c->spin_limit = 1u << 26;
    if (const char* e = getenv("GMC_DEBUG_SPIN_LIMIT")) c->spin_limit = (unsigned)strtoul(e, nullptr, 10);
    if (cudaHostAlloc((void**)&c->h_err, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&c->d_err, c->h_err, 0) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        GMC_FAIL(GMC_ECUDA, "gmc_create: cannot allocate the mapped device-error flag");
    }
    *c->h_err = 0;
    c->sgs = nullptr;
    c->dev.H = H;
    c->dev.W = W;
    // loss partials: one per (chain, row-tile); sized for the residual kernel's tiling (residual.cu)
    c->n_tiles = 0;
    *out = c;
    return GMC_OK;
}

extern "C" int gmc_destroy(gmc_ctx* c) {
    if (!c) return GMC_OK;
    cudaSetDevice(c->device);
    cudaFree(c->d_static);
    cudaFree(c->d_flags);
    cudaFree(c->d_centre);
    cudaFree(c->d_partials);
    cudaFree(c->d_pairs);
    cudaFree(c->d_twiddle);
    cudaFree(c->d_perm);
    cudaFree(c->d_pos);
    cudaFree(c->d_ksq);
    cudaFree(c->d_edge_masks);
    cudaFree(c->d_phase);
    cudaFree(c->d_sched);
    for (int k = 0; k < GMC_SCHED_SLOTS; ++k)
        if (c->sched_ev[k]) cudaEventDestroy(c->sched_ev[k]);
    if (c->h_err) cudaFreeHost(c->h_err);
    cudaFree(c->d_ring);
    cudaFree(c->d_pflags);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    gmc_sgs_destroy(c);
    delete c;
    return GMC_OK;
}

extern "C" int gmc_set_static(gmc_ctx* c, const double* surf, const double* velx, const double* vely, const double* dhdt,
                              const double* smb, const uint8_t* gate_mask, const uint8_t* mc_mask,
                              const int32_t* centre_cells, int64_t n_centre_cells, const double* crf_weight,
                              double resolution, double sigma_mc) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_set_static: ctx is NULL");
    if (!surf || !velx || !vely || !dhdt || !smb || !gate_mask || !mc_mask)
        GMC_FAIL(GMC_EINVAL, "gmc_set_static: NULL field or mask pointer");
    if (!(resolution > 0.0)) GMC_FAIL(GMC_EINVAL, "gmc_set_static: resolution must be > 0");
    if (n_centre_cells < 0 || (n_centre_cells > 0 && !centre_cells))
        GMC_FAIL(GMC_EINVAL, "gmc_set_static: centre_cells/n_centre_cells inconsistent");
    GMC_CUDA(cudaSetDevice(c->device));
    const size_t n = (size_t)c->H * c->W;
    if (!c->d_static) GMC_CUDA(cudaMalloc(&c->d_static, 12 * n * sizeof(double)));   // 6 planes + 3 packed pair arrays
    if (!c->d_flags) GMC_CUDA(cudaMalloc(&c->d_flags, n));
    const double* src[6] = {surf, velx, vely, dhdt, smb, crf_weight};
    for (int k = 0; k < 6; ++k)
        if (src[k]) GMC_CUDA(cudaMemcpy(c->d_static + k * n, src[k], n * sizeof(double), cudaMemcpyDefault));
    {   // packed pairs for the step kernel's block stencil (one 16 B load per flux operand)
        std::vector<double> h(5 * n), pk(6 * n);
        GMC_CUDA(cudaMemcpy(h.data(), c->d_static, 5 * n * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            pk[2 * i] = h[i];                    // sv = {surf, velx}
            pk[2 * i + 1] = h[n + i];
            pk[2 * n + 2 * i] = h[i];            // sy = {surf, vely}
            pk[2 * n + 2 * i + 1] = h[2 * n + i];
            pk[4 * n + 2 * i] = h[3 * n + i];    // ds = {dhdt, smb}
            pk[4 * n + 2 * i + 1] = h[4 * n + i];
        }
        GMC_CUDA(cudaMemcpy(c->d_static + 6 * n, pk.data(), 6 * n * sizeof(double), cudaMemcpyHostToDevice));
    }
    // pack the two masks into one flag byte per cell (host side; setup is not on the hot path)
    std::vector<uint8_t> g(n), m(n), fl(n);
    GMC_CUDA(cudaMemcpy(g.data(), gate_mask, n, cudaMemcpyDefault));
    GMC_CUDA(cudaMemcpy(m.data(), mc_mask, n, cudaMemcpyDefault));
    for (size_t i = 0; i < n; ++i) {
        if (g[i] > 1 || m[i] > 1) GMC_FAIL(GMC_EINVAL, "gmc_set_static: masks must hold 0/1 (cell %zu has %d/%d)", i, g[i], m[i]);
        fl[i] = (uint8_t)((g[i] ? FLAG_GATE : 0) | (m[i] ? FLAG_MC : 0));
    }
    GMC_CUDA(cudaMemcpy(c->d_flags, fl.data(), n, cudaMemcpyHostToDevice));
    cudaFree(c->d_centre);
    c->d_centre = nullptr;
    if (n_centre_cells > 0) {
        std::vector<int32_t> cc((size_t)n_centre_cells);
        GMC_CUDA(cudaMemcpy(cc.data(), centre_cells, cc.size() * sizeof(int32_t), cudaMemcpyDefault));
        for (int32_t v : cc)
            if (v < 0 || (size_t)v >= n) GMC_FAIL(GMC_EINVAL, "gmc_set_static: centre cell index %d outside the grid", v);
        GMC_CUDA(cudaMalloc(&c->d_centre, cc.size() * sizeof(int32_t)));
        GMC_CUDA(cudaMemcpy(c->d_centre, cc.data(), cc.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    GmcDev& d = c->dev;
    d.surf = c->d_static;
    d.velx = c->d_static + n;
    d.vely = c->d_static + 2 * n;
    d.dhdt = c->d_static + 3 * n;
    d.smb = c->d_static + 4 * n;
    d.crf_weight = crf_weight ? c->d_static + 5 * n : nullptr;
    d.sv = reinterpret_cast<const double2*>(c->d_static + 6 * n);
    d.sy = reinterpret_cast<const double2*>(c->d_static + 8 * n);
    d.ds = reinterpret_cast<const double2*>(c->d_static + 10 * n);
    d.r_res = 1.0 / resolution;
    d.r_two_res = 1.0 / (2.0 * resolution);
    d.flags = c->d_flags;
    d.centre_cells = c->d_centre;
    d.n_centre_cells = n_centre_cells;
    d.res = resolution;
    d.two_res = 2.0 * resolution;              // np.gradient: 2. * ax_dx
    d.two_sigma2 = 2 * (sigma_mc * sigma_mc);  // MCMC.py:1041: 2*self.sigma_mc**2
    c->have_static = true;
    return GMC_OK;
}

extern "C" int gmc_set_field_model(gmc_ctx* c, int model, double smoothness, int isotropic, double range_min_x,
                                   double range_max_x, double range_min_y, double range_max_y, double scale_min,
                                   double scale_max, double nugget_max) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_set_field_model: ctx is NULL");
    if (model != GMC_GAUSSIAN && model != GMC_EXPONENTIAL && model != GMC_MATERN)
        GMC_FAIL(GMC_EINVAL, "gmc_set_field_model: model must be Gaussian(0), Exponential(1) or Matern(2)");
    if (nugget_max < 0.0) GMC_FAIL(GMC_EINVAL, "gmc_set_field_model: nugget_max must be >= 0");
    GmcFieldModel& f = c->dev.fm;
    f.model = model;
    f.isotropic = isotropic ? 1 : 0;
    // MCMC.py:232: nu = RF.smoothness or 1.0
    f.smoothness = (model == GMC_MATERN) ? ((smoothness == 0.0 || smoothness != smoothness) ? 1.0 : smoothness) : 1.0;
    f.range_min_x = range_min_x;
    f.range_max_x = range_max_x;
    f.range_min_y = range_min_y;
    f.range_max_y = range_max_y;
    f.scale_min = scale_min;
    f.scale_max = scale_max;
    f.nugget_max = nugget_max;
    const double nu = f.smoothness;
    f.matern_num = 4 * M_PI * tgamma(nu + 1) * pow(2 * nu, nu);   // MCMC.py:236 numerator
    f.matern_gamma = tgamma(nu);
    c->have_model = true;
    return GMC_OK;
}

// RandField.set_generation_method (MCMC.py:514-522)
extern "C" int gmc_set_generation_method(gmc_ctx* c, int spectral, int n_modes) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_set_generation_method: ctx is NULL");
    if (!spectral && (n_modes < 1 || n_modes > (1 << 20)))
        GMC_FAIL(GMC_EINVAL, "gmc_set_generation_method: n_modes=%d outside [1, 2^20]", n_modes);
    c->spectral = spectral ? 1 : 0;
    if (!spectral) c->n_modes = n_modes;
    return GMC_OK;
}

// ---- FFT plans -----------------------------------------------------------------------------------------------
static bool factorize(int n, GmcFftPlan& p) {
    p.n = n;
    p.n_factors = 0;
    int m = n, twos = 0;
    while (m % 2 == 0) {
        m /= 2;
        ++twos;
    }
    auto push = [&](int r) {
        if (p.n_factors >= GMC_MAX_FACTORS) return false;
        p.radix[p.n_factors++] = r;
        return true;
    };
    for (; twos >= 3; twos -= 3)
        if (!push(8)) return false;
    if (twos == 2 && !push(4)) return false;
    if (twos == 1 && !push(2)) return false;
    for (int r = 3; r < GMC_MAX_RADIX; r += 2) {   // 3, 5, 7, 11, 13 have unrolled stages; larger primes a generic one
        while (m % r == 0) {
            if (!push(r)) return false;
            m /= r;
        }
    }
    return m == 1;
}

extern "C" int gmc_set_blocks(gmc_ctx* c, int n_pairs, const int32_t* pair_w, const int32_t* pair_h,
                              const double* edge_masks, const int64_t* offsets, double field_resolution) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_set_blocks: ctx is NULL");
    if (n_pairs < 1 || !pair_w || !pair_h || !edge_masks || !offsets)
        GMC_FAIL(GMC_EINVAL, "gmc_set_blocks: empty block table or NULL pointer");
    if (!(field_resolution > 0.0)) GMC_FAIL(GMC_EINVAL, "gmc_set_blocks: field_resolution must be > 0");
    c->field_res = field_resolution;
    GMC_CUDA(cudaSetDevice(c->device));
    std::vector<GmcPair> pairs(n_pairs);
    std::vector<GmcFftPlan> plans;
    std::vector<double2> tw;
    std::vector<int16_t> perm, pos;
    std::vector<double> ksq;
    std::vector<std::pair<int, int>> ksq_index;   // (edge length, offset)
    const long double two_pi = 2.0L * 3.14159265358979323846264338327950288L;
    auto plan_for = [&](int n) -> int {
        for (size_t i = 0; i < plans.size(); ++i)
            if (plans[i].n == n) return (int)i;
        GmcFftPlan p;
        if (!factorize(n, p)) return -1;
        p.tw_off = (int)tw.size();
        for (int k = 0; k < n; ++k) {
            const long double a = two_pi * (long double)k / (long double)n;
            tw.push_back(make_double2((double)cosl(a), (double)sinl(a)));
        }
        p.htw_off = (int)tw.size();
        for (int k = 0; k < n; ++k) {
            const long double a = two_pi * (long double)k / (long double)(2 * n);
            tw.push_back(make_double2((double)cosl(a), (double)sinl(a)));
        }
        std::vector<int> cur(1, 0);
        for (int s = 0; s < p.n_factors; ++s) {
            const int r = p.radix[s];
            std::vector<int> nxt;
            nxt.reserve(cur.size() * r);
            for (int q = 0; q < r; ++q)
                for (int v : cur) nxt.push_back(q + r * v);
            cur.swap(nxt);
        }
        p.perm_off = (int)perm.size();
        p.pos_off = (int)pos.size();
        std::vector<int16_t> inv(n);
        for (int i = 0; i < n; ++i) {
            perm.push_back((int16_t)cur[i]);
            inv[cur[i]] = (int16_t)i;
        }
        pos.insert(pos.end(), inv.begin(), inv.end());
        plans.push_back(p);
        return (int)plans.size() - 1;
    };
    auto ksq_for = [&](int n) -> int {
        for (auto& kv : ksq_index)
            if (kv.first == n) return kv.second;
        const int off = (int)ksq.size();
        const double val = 1.0 / (n * field_resolution);   // np.fft.fftfreq: val = 1.0/(n*d); results * val
        for (int i = 0; i <= n / 2; ++i) {
            const double kk = (double)i * val * 2 * M_PI;  // MCMC.py:221: fftfreq(...) * 2 * np.pi
            ksq.push_back(kk * kk);
        }
        ksq_index.push_back({n, off});
        return off;
    };
    int64_t total = 0;
    int mh = 0, mw = 0;
    for (int i = 0; i < n_pairs; ++i) {
        const int w = pair_w[i], h = pair_h[i];
        if (w < 2 || h < 2 || (w & 1) || (h & 1))
            GMC_FAIL(GMC_EINVAL, "gmc_set_blocks: pair %d is %dx%d; block edges must be even and >= 2 (MCMC.py:579)", i, h, w);
        if (h > c->H || w > c->W)
            GMC_FAIL(GMC_ESHAPE, "gmc_set_blocks: pair %d (%dx%d) is larger than the %dx%d grid", i, h, w, c->H, c->W);
        if (w > 32767 || h > 32767) GMC_FAIL(GMC_EUNSUPPORTED, "gmc_set_blocks: block edge > 32767");
        pairs[i].h = h;
        pairs[i].w = w;
        if (h > GMC_MAX_EDGE || w > GMC_MAX_EDGE)
            GMC_FAIL(GMC_EUNSUPPORTED, "gmc_set_blocks: pair %d (%dx%d) exceeds the largest supported block edge %d", i, h, w, GMC_MAX_EDGE);
        const int ip = plan_for(h), iw = plan_for(w / 2);
        if (ip < 0 || iw < 0)
            GMC_FAIL(GMC_EUNSUPPORTED, "gmc_set_blocks: pair %d (%dx%d) has a prime factor >= %d", i, h, w, GMC_MAX_RADIX);
        pairs[i].ph = plans[ip];
        pairs[i].pw = plans[iw];
        pairs[i].pitchc = (w / 2 + 1) | 1;                 // odd pitch (in 16 B units): bank-conflict-free row pass
        pairs[i].ksq_off_h = ksq_for(h);
        pairs[i].ksq_off_w = ksq_for(w);
        pairs[i].mask_off = offsets[i];
        total = std::max<int64_t>(total, offsets[i] + (int64_t)h * w);
        mh = std::max(mh, h);
        mw = std::max(mw, w);
    }
    cudaFree(c->d_pairs);
    cudaFree(c->d_twiddle);
    cudaFree(c->d_perm);
    cudaFree(c->d_pos);
    cudaFree(c->d_ksq);
    cudaFree(c->d_edge_masks);
    c->d_pairs = nullptr;
    c->d_twiddle = nullptr;
    c->d_perm = c->d_pos = nullptr;
    c->d_ksq = nullptr;
    c->d_edge_masks = nullptr;
    c->have_blocks = false;
#define UPLOAD(dst, vec)                                                                          \
    GMC_CUDA(cudaMalloc(&(dst), (vec).size() * sizeof((vec)[0])));                                 \
    GMC_CUDA(cudaMemcpy((dst), (vec).data(), (vec).size() * sizeof((vec)[0]), cudaMemcpyHostToDevice))
    UPLOAD(c->d_pairs, pairs);
    UPLOAD(c->d_twiddle, tw);
    UPLOAD(c->d_perm, perm);
    UPLOAD(c->d_pos, pos);
    UPLOAD(c->d_ksq, ksq);
#undef UPLOAD
    GMC_CUDA(cudaMalloc(&c->d_edge_masks, (size_t)total * sizeof(double)));
    GMC_CUDA(cudaMemcpy(c->d_edge_masks, edge_masks, (size_t)total * sizeof(double), cudaMemcpyDefault));
    c->h_pairs = pairs;
    c->h_plans = plans;
    c->max_h = mh;
    c->max_w = mw;
    GmcDev& d = c->dev;
    d.n_pairs = n_pairs;
    d.pairs = c->d_pairs;
    d.twiddle = c->d_twiddle;
    d.perm = c->d_perm;
    d.pos = c->d_pos;
    d.ksq = c->d_ksq;
    d.edge_masks = c->d_edge_masks;
    const int rc = gmc_step_configure(c);
    if (rc != GMC_OK) return rc;
    c->have_blocks = true;
    return GMC_OK;
}

extern "C" int64_t gmc_launch_count(const gmc_ctx* c) { return c ? c->launches : 0; }

int gmc_sched_acquire(gmc_ctx* c, int C, cudaStream_t st, int** sched_out) {
    const size_t area = (size_t)c->max_chains + 1;
    if (!c->d_sched) GMC_CUDA(cudaMalloc(&c->d_sched, GMC_SCHED_SLOTS * area * sizeof(int)));
    const int k = (int)(c->sched_next++ % GMC_SCHED_SLOTS);
    if (!c->sched_ev[k]) GMC_CUDA(cudaEventCreateWithFlags(&c->sched_ev[k], cudaEventDisableTiming));
    // more than GMC_SCHED_SLOTS scheduled launches in flight: this one queues behind the area's previous user
    if (c->sched_used[k]) GMC_CUDA(cudaStreamWaitEvent(st, c->sched_ev[k], 0));
    int* sched = c->d_sched + (size_t)k * area;
    GMC_CUDA(cudaMemsetAsync(sched, 0, (size_t)(C + 1) * sizeof(int), st));
    c->sched_cur = k;
    *sched_out = sched;
    return GMC_OK;
}

void gmc_sched_release(gmc_ctx* c, cudaStream_t st) {
    if (c->sched_cur < 0) return;
    cudaEventRecord(c->sched_ev[c->sched_cur], st);
    c->sched_used[c->sched_cur] = true;
    c->sched_cur = -1;
}

int gmc_check_device_error(gmc_ctx* c, const char* who) {
    const int code = *(volatile int*)c->h_err;
    if (code == 0) return GMC_OK;
    *c->h_err = 0;
    GMC_FAIL(GMC_ECUDA, "%s: a kernel gave up a bounded in-kernel wait (device error %d: a chunk's predecessor or a tile copy did not "
             "complete within the spin limit, e.g. under a debugger or on a shared GPU); the chain state of that launch is invalid", who, code);
}

// Register-resident DFMA loop: the FP64 FMA rate this GPU sustains (8 independent chains per thread, 32 warps per SM), the
// denominator for the FP64-pipe fractions quoted for the step and kriging kernels.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double b, double c0) {
    double a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 1.0 + 1e-3 * (threadIdx.x + 32 * k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c0);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int gmc_debug_fp64_peak(gmc_ctx* c, double* tflops_out) {
    if (!c || !tflops_out) GMC_FAIL(GMC_EINVAL, "gmc_debug_fp64_peak: NULL argument");
    GMC_CUDA(cudaSetDevice(c->device));
    const int ctas = c->sm_count * 4, iters = 4096;
    double* d = nullptr;
    GMC_CUDA(cudaMalloc(&d, (size_t)ctas * 256 * sizeof(double)));
    cudaEvent_t e0, e1;
    GMC_CUDA(cudaEventCreate(&e0));
    GMC_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {          // first pass warms up; best of the rest
        cudaEventRecord(e0);
        fp64_peak_kernel<<<ctas, 256>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        GMC_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    GMC_CUDA(cudaGetLastError());
    *tflops_out = 2.0 * 8.0 * 4.0 * iters * (double)ctas * 256.0 / (best * 1e-3) / 1e12;
    return GMC_OK;
}

// Reports (then clears) the device-error flag; synchronize != 0 waits for the device first, 0 reads what finished launches
// stored (for callers that have synchronised the streams they care about themselves).
extern "C" int gmc_check(gmc_ctx* c, int synchronize) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_check: ctx is NULL");
    if (synchronize) {
        GMC_CUDA(cudaSetDevice(c->device));
        GMC_CUDA(cudaDeviceSynchronize());
    }
    return gmc_check_device_error(c, "gmc_check");
}

// CTA size of the fused step kernel.  Auto gives a launch with no more chains than SMs 512-thread CTAs (one per SM); that
// is right when the launch has the GPU to itself and wrong when several launches share it (chain ranges on different
// streams, several steps in flight): their 512-thread CTAs cannot co-reside and the launches serialise.  Such callers
// select mode 1.
extern "C" int gmc_set_step_cta(gmc_ctx* c, int mode) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_set_step_cta: ctx is NULL");
    if (mode < 0 || mode > 3) GMC_FAIL(GMC_EINVAL, "gmc_set_step_cta: mode %d outside {0 auto, 1 narrow, 2 wide, 3 split}", mode);
    c->step_cta_mode = mode;
    return GMC_OK;
}

extern "C" int gmc_step_kernel_info(const gmc_ctx* c, int* smem_bytes, int* threads, int* ctas_per_sm) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_step_kernel_info: ctx is NULL");
    if (!c->have_blocks) GMC_FAIL(GMC_ESTATE, "gmc_step_kernel_info: call gmc_set_blocks first");
    if (smem_bytes) *smem_bytes = c->step_smem_bytes;
    if (threads) *threads = GMC_STEP_THREADS;
    if (ctas_per_sm) *ctas_per_sm = c->step_ctas_per_sm;
    return GMC_OK;
}
