// step.cu — K1 (proposal field synthesis) and K4 (Metropolis step) of the large-scale chain, fused.
//
// One CTA owns one chain.  Per iteration everything lives in shared memory:
//   K1  the proposal f = taper * standardised Re(ifft2(noise * sqrt(S)))  (MCMC.py:176-254, 742-778).  Only the real part
//       of the inverse transform is kept by the reference, so the spectrum is folded to its Hermitian part
//       X_h(k) = (X(k) + conj(X(-k)))/2 and a complex-to-real inverse transform of the half plane [h][w/2+1] is run in
//       place: column DFTs (mixed radix, decimation in time, digit-reversed load), a pair-recombination pass, and row DFTs
//       of length w/2 whose complex output IS the real row (x[2m], x[2m+1]).  Mean and variance follow from the spectrum
//       (DC term and Parseval), so standardisation is a scale factor applied before the row pass — no spatial reduction.
//       With device RNG the Hermitian half plane is drawn directly (same distribution, half the normals).
//   K4  the candidate bed of the clipped block plus a one-cell halo is staged as a tile, the residual is recomputed on the
//       block only, the loss changes by the block's delta, and accept/reject plus in-place write-back happen without leaving
//       the kernel (MCMC.py:1263-1360).  HBM sees the bed halo tile, the old block residual and, on accept, the write-back.
#include "common.cuh"
#include <cstdlib>

#undef GMC_STEP_THREADS
#undef GMC_STEP_MIN_CTAS
#define GMC_STEP_THREADS 256
#define GMC_STEP_MIN_CTAS 2
namespace t256 {
#include "step_kernels.cuh"
}
#undef GMC_STEP_THREADS
#undef GMC_STEP_MIN_CTAS
#define GMC_STEP_THREADS 512
#define GMC_STEP_MIN_CTAS 1
#define GMC_STEP_RUN_ONLY
namespace t512 {
#include "step_kernels.cuh"
}
#undef GMC_STEP_RUN_ONLY
#undef GMC_STEP_THREADS
#undef GMC_STEP_MIN_CTAS
#define GMC_STEP_THREADS 256
#define GMC_STEP_MIN_CTAS 2
using namespace t256;   // host code below launches the default-size kernels by their plain names

// ---------------------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------------------
static size_t tile_bytes(int h, int w) { return (size_t)(h + 2) * (w + 4) * sizeof(double); }   // pitch <= w + 4 (aligned staging)
static size_t field_bytes(const GmcPair& p) { return (size_t)p.h * p.pitchc * sizeof(double2); }

int gmc_step_configure(gmc_ctx* c) {
    // layout: [ field / new residuals | tile ]; the tile starts after the largest field so one offset serves all pairs
    size_t fmax = 0, tmax = 0;
    for (const GmcPair& p : c->h_pairs) {
        fmax = std::max(fmax, field_bytes(p));
        tmax = std::max(tmax, tile_bytes(p.h, p.w));
    }
    fmax = (fmax + 15) & ~(size_t)15;
    const size_t need = fmax + ((tmax + 15) & ~(size_t)15);
    cudaDeviceProp prop;
    GMC_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (need > prop.sharedMemPerBlockOptin)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_set_blocks: a %dx%d block needs %zu B of shared memory; the device offers %zu B",
                 c->max_h, c->max_w, need, (size_t)prop.sharedMemPerBlockOptin);
    GMC_CUDA(cudaFuncSetAttribute(run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    GMC_CUDA(cudaFuncSetAttribute(field_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    GMC_CUDA(cudaFuncSetAttribute(field_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    c->step_smem_bytes = (int)need;
    c->step_tile_off = (int)(fmax / sizeof(double));
    // randomization method: field [h][w] at 0 (<= fmax), mode tables share the tile region (the tile is staged after them)
    const size_t rm_tab = (size_t)4 * RM_CHUNK * RM_EDGE * sizeof(double);
    const size_t rm_need = fmax + std::max((tmax + 15) & ~(size_t)15, rm_tab);
    c->rm_smem_bytes = 0;
    if (rm_need <= prop.sharedMemPerBlockOptin) {
        GMC_CUDA(cudaFuncSetAttribute(run_randmeth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rm_need));
        GMC_CUDA(cudaFuncSetAttribute(field_randmeth_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rm_need));
        GMC_CUDA(cudaFuncSetAttribute(field_randmeth_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rm_need));
        c->rm_smem_bytes = (int)rm_need;
        c->rm_tile_off = c->step_tile_off;
    }
    int nb = 0;
    GMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, run_kernel, GMC_STEP_THREADS, need));
    c->step_ctas_per_sm = nb;
    // the 512-thread variant (one CTA per SM) for launches with no more chains than SMs
    GMC_CUDA(cudaFuncSetAttribute(tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    GMC_CUDA(cudaFuncSetAttribute(field_producer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fmax));
    c->step_wide_ctas = 0;
    if (cudaFuncSetAttribute(t512::run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) == cudaSuccess) {
        int nw = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nw, t512::run_kernel, 512, need) == cudaSuccess) c->step_wide_ctas = nw;
    }
    cudaGetLastError();
    return GMC_OK;
}

static int check_step(gmc_ctx* c, int C, const char* who, bool need_blocks) {
    if (!c) GMC_FAIL(GMC_EINVAL, "%s: ctx is NULL", who);
    if (!c->have_static) GMC_FAIL(GMC_ESTATE, "%s: call gmc_set_static first", who);
    if (need_blocks && (!c->have_blocks || !c->have_model))
        GMC_FAIL(GMC_ESTATE, "%s: call gmc_set_field_model and gmc_set_blocks first", who);
    if (C < 1 || C > c->max_chains) GMC_FAIL(GMC_ESHAPE, "%s: C=%d outside [1,%d]", who, C, c->max_chains);
    GMC_CUDA(cudaSetDevice(c->device));
    return GMC_OK;
}

extern "C" int gmc_field_spectral(gmc_ctx* c, int n, const int32_t* pair, const double* scale, const double* nug,
                                  const double* range_x, const double* range_y, const double* z_re, const double* z_im,
                                  const double* z_nug, const uint64_t* seeds, uint64_t iter, int apply_taper,
                                  double* f_out, int64_t stride, void* stream) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_field_spectral: ctx is NULL");
    if (!c->have_blocks || !c->have_model) GMC_FAIL(GMC_ESTATE, "gmc_field_spectral: call gmc_set_field_model and gmc_set_blocks first");
    if (n < 1) GMC_FAIL(GMC_EINVAL, "gmc_field_spectral: n must be >= 1");
    if (!pair || !scale || !nug || !range_x || !range_y || !f_out) GMC_FAIL(GMC_EINVAL, "gmc_field_spectral: NULL argument");
    if (stride < (int64_t)c->max_h * c->max_w) GMC_FAIL(GMC_ESHAPE, "gmc_field_spectral: stride %lld < max block %d", (long long)stride, c->max_h * c->max_w);
    const bool inject = z_re || z_im || z_nug;
    if (inject && !(z_re && z_im && z_nug)) GMC_FAIL(GMC_EINVAL, "gmc_field_spectral: z_re, z_im, z_nug must be all given or all NULL");
    if (!inject && !seeds) GMC_FAIL(GMC_EINVAL, "gmc_field_spectral: seeds required when no normals are injected");
    GMC_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (inject)
        field_kernel<true><<<n, GMC_STEP_THREADS, c->step_smem_bytes, st>>>(c->dev, pair, scale, nug, range_x, range_y, z_re, z_im, z_nug, seeds, iter, apply_taper, f_out, stride);
    else
        field_kernel<false><<<n, GMC_STEP_THREADS, c->step_smem_bytes, st>>>(c->dev, pair, scale, nug, range_x, range_y, z_re, z_im, z_nug, seeds, iter, apply_taper, f_out, stride);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_field_randmeth(gmc_ctx* c, int n, const int32_t* pair, const double* scale, const double* nug,
                                  const double* range_x, const double* range_y, const double* angle_deg, int n_modes,
                                  const double* modes, const double* z_nug, const uint64_t* seeds, uint64_t iter,
                                  int apply_taper, double* f_out, int64_t stride, void* stream) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: ctx is NULL");
    if (!c->have_blocks || !c->have_model) GMC_FAIL(GMC_ESTATE, "gmc_field_randmeth: call gmc_set_field_model and gmc_set_blocks first");
    if (n < 1) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: n must be >= 1");
    if (n_modes < 1 || n_modes > (1 << 20)) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: n_modes=%d outside [1, 2^20]", n_modes);
    if (!pair || !scale || !nug || !range_x || !range_y || !angle_deg || !f_out) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: NULL argument");
    if (stride < (int64_t)c->max_h * c->max_w) GMC_FAIL(GMC_ESHAPE, "gmc_field_randmeth: stride %lld < max block %d", (long long)stride, c->max_h * c->max_w);
    const bool inject = modes || z_nug;
    if (inject && !(modes && z_nug)) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: modes and z_nug must be both given or both NULL");
    if (!inject && !seeds) GMC_FAIL(GMC_EINVAL, "gmc_field_randmeth: seeds required when no modes are injected");
    if (!c->rm_smem_bytes) GMC_FAIL(GMC_EUNSUPPORTED, "gmc_field_randmeth: the block table needs more shared memory than the device offers");
    GMC_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (inject)
        field_randmeth_kernel<true><<<n, GMC_STEP_THREADS, c->rm_smem_bytes, st>>>(c->dev, n_modes, c->field_res, c->rm_tile_off, pair, scale, nug, range_x, range_y, angle_deg, modes, z_nug, seeds, iter, apply_taper, f_out, stride);
    else
        field_randmeth_kernel<false><<<n, GMC_STEP_THREADS, c->rm_smem_bytes, st>>>(c->dev, n_modes, c->field_res, c->rm_tile_off, pair, scale, nug, range_x, range_y, angle_deg, modes, z_nug, seeds, iter, apply_taper, f_out, stride);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_step_injected(gmc_ctx* c, double* bed, double* mcres, double* ssq, const double* f, int64_t f_stride,
                                 const int32_t* hw, const int32_t* centre, const double* u, int hmax, int wmax,
                                 uint8_t* accepted_out, double* loss_out, double* loss_next_out, int32_t* resampled, int C,
                                 void* stream) {
    int rc = check_step(c, C, "gmc_step_injected", false);
    if (rc) return rc;
    if (!bed || !mcres || !ssq || !f || !hw || !centre || !u) GMC_FAIL(GMC_EINVAL, "gmc_step_injected: NULL argument");
    if (hmax < 2 || wmax < 2 || hmax > c->H || wmax > c->W)
        GMC_FAIL(GMC_ESHAPE, "gmc_step_injected: hmax x wmax = %dx%d must lie in [2, grid %dx%d]", hmax, wmax, c->H, c->W);
    if (f_stride < (int64_t)hmax * wmax) GMC_FAIL(GMC_ESHAPE, "gmc_step_injected: f_stride smaller than hmax*wmax");
    const size_t smem = (((size_t)hmax * wmax * sizeof(double) + tile_bytes(hmax, wmax)) + 15) & ~(size_t)15;
    cudaDeviceProp prop;
    GMC_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (smem > prop.sharedMemPerBlockOptin)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_step_injected: a %dx%d block needs %zu B of shared memory", hmax, wmax, smem);
    GMC_CUDA(cudaFuncSetAttribute(step_injected_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    step_injected_kernel<<<C, GMC_STEP_THREADS, smem, (cudaStream_t)stream>>>(c->dev, bed, mcres, ssq, f, f_stride, hw, centre, u,
                                                                             accepted_out, loss_out, loss_next_out, resampled,
                                                                             hmax, wmax);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

extern "C" int gmc_run(gmc_ctx* c, double* bed, double* mcres, double* ssq, const uint64_t* seeds, uint64_t iter0,
                       int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache, int64_t cache_stride,
                       int64_t cache_offset, int32_t* resampled, int resync_every, int C, void* stream) {
    int rc = check_step(c, C, "gmc_run", true);
    if (rc) return rc;
    if (!bed || !mcres || !ssq || !seeds) GMC_FAIL(GMC_EINVAL, "gmc_run: NULL argument");
    if (n_steps < 0 || resync_every < 0) GMC_FAIL(GMC_EINVAL, "gmc_run: negative n_steps or resync_every");
    if ((loss_cache || step_cache || blocks_cache) && (cache_offset < 0 || cache_offset + n_steps > cache_stride))
        GMC_FAIL(GMC_ESHAPE, "gmc_run: cache window [%lld, %lld) exceeds stride %lld", (long long)cache_offset,
                 (long long)(cache_offset + n_steps), (long long)cache_stride);
    if (n_steps == 0) return GMC_OK;
    if (!c->spectral) {                      // RandField.set_generation_method(False): A5 proposal
        if (!c->rm_smem_bytes) GMC_FAIL(GMC_EUNSUPPORTED, "gmc_run: the block table needs more shared memory than the device offers");
        run_randmeth_kernel<<<C, GMC_STEP_THREADS, c->rm_smem_bytes, (cudaStream_t)stream>>>(
            c->dev, c->n_modes, c->field_res, c->rm_tile_off, bed, mcres, ssq, seeds, iter0, n_steps, loss_cache, step_cache,
            blocks_cache, cache_stride, cache_offset, resampled, resync_every, c->rm_tile_off);
        c->launches++;
        GMC_CUDA(cudaGetLastError());
        return GMC_OK;
    }
    rc = gmc_check_device_error(c, "gmc_run");      // a flag left by an earlier (finished) launch: refuse to build on its state
    if (rc) return rc;
    // more chains than resident CTAs: cut the run into (chunk, chain) items handed out dynamically, so that the tail of the
    // last wave does not idle the GPU (e.g. 512 chains on 296 slots: 1.73 instead of 2 waves)
    const int slots = std::max(1, c->step_ctas_per_sm) * c->sm_count;
    const bool vec = (c->W % 2 == 0) && (((uintptr_t)bed & 15) == 0);
    int grid = C, chunk = n_steps;
    int* sched = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    if (C > slots && vec && n_steps > 1 && !getenv("GMC_STATIC_SCHED")) {
        // launches for disjoint chain ranges may run concurrently on different streams (run_pipelined): each takes the next
        // of GMC_SCHED_SLOTS scheduler areas (a launch that finds its area still in use queues behind that user)
        rc = gmc_sched_acquire(c, C, st, &sched);
        if (rc) return rc;
        grid = slots;
        // ~32 chunks per chain, at most 256 iterations each: short chunks keep the tail small and the waits on a
        // predecessor chunk (bounded spin in the kernel) in the millisecond range however long the launch is
        chunk = std::min(256, std::max(4, (n_steps + 31) / 32));
    }
    // Launches that under-fill the GPU (auto mode; gmc_set_step_cta(1) = "this launch shares the GPU" disables both):
    //  * no more chains than SMs: 512-thread CTAs, one per SM;
    //  * opt-in, fewer than half as many chains as CTA slots: SPLIT mode - a producer CTA synthesises the fields of the
    //    coming steps while the consumer CTA of the same chain runs the Metropolis tail (two kernels, the producer on the
    //    context's auxiliary stream, fenced to the caller's stream by events).
    static const char* wide_env = getenv("GMC_STEP_WIDE");            // "0" / "1" force the 512-thread choice (A/B runs)
    static const char* split_env = getenv("GMC_STEP_SPLIT");          // "0" / "1" force the split choice (A/B runs)
    // (split is opt-in - mode 3 / GMC_STEP_SPLIT=1: measured 3.57 M chain-steps/s for 128 x 500^2 and 2.79 M for 128 x 2000^2
    // against 4.86 M / 4.36 M with 512-thread CTAs: the producer's per-step scalar preparation, hidden behind the residual
    // phase by the helper warp in the fused kernel, is exposed there)
    bool split = false;
    if (c->step_cta_mode == 3) split = !sched && 2 * C < slots;
    if (split_env) split = !sched && 2 * C < slots && split_env[0] == '1';
    bool wide = !split && !sched && c->step_wide_ctas >= 1 && C <= c->sm_count;
    if (c->step_cta_mode == 1 || c->step_cta_mode == 3) wide = false;  // gmc_set_step_cta: the launch shares the GPU / split only
    if (c->step_cta_mode == 2) wide = !sched && c->step_wide_ctas >= 1;
    if (wide_env) wide = !split && !sched && c->step_wide_ctas >= 1 && wide_env[0] == '1';
    if (split) {
        const int depth = 3;
        const int64_t fstride = (int64_t)c->max_h * c->max_w;
        const size_t need_ring = (size_t)C * depth * fstride * sizeof(double);
        if (need_ring > c->ring_bytes) {
            cudaFree(c->d_ring);
            c->d_ring = nullptr;
            c->ring_bytes = 0;
            GMC_CUDA(cudaMalloc(&c->d_ring, need_ring));
            c->ring_bytes = need_ring;
        }
        if (!c->d_pflags) GMC_CUDA(cudaMalloc(&c->d_pflags, (size_t)2 * c->max_chains * sizeof(int)));
        if (!c->aux_stream) GMC_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
        if (!c->ev_fork) GMC_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        if (!c->ev_join) GMC_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
        GMC_CUDA(cudaMemsetAsync(c->d_pflags, 0, (size_t)2 * C * sizeof(int), st));
        GMC_CUDA(cudaEventRecord(c->ev_fork, st));
        GMC_CUDA(cudaStreamWaitEvent(c->aux_stream, c->ev_fork, 0));
        field_producer_kernel<<<C, GMC_STEP_THREADS, (size_t)c->step_tile_off * sizeof(double), c->aux_stream>>>(
            c->dev, seeds, iter0, n_steps, c->d_ring, fstride, depth, c->d_pflags, c->d_err, c->spin_limit);
        tail_kernel<<<C, GMC_STEP_THREADS, c->step_smem_bytes, st>>>(
            c->dev, bed, mcres, ssq, seeds, iter0, n_steps, loss_cache, step_cache, blocks_cache, cache_stride, cache_offset, resampled,
            resync_every, c->step_tile_off, c->d_ring, fstride, depth, c->d_pflags, c->d_err, c->spin_limit);
        GMC_CUDA(cudaEventRecord(c->ev_join, c->aux_stream));
        GMC_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
        c->launches += 2;
        GMC_CUDA(cudaGetLastError());
        return GMC_OK;
    }
    if (wide)
        t512::run_kernel<<<grid, 512, c->step_smem_bytes, st>>>(
            c->dev, bed, mcres, ssq, seeds, iter0, n_steps, loss_cache, step_cache, blocks_cache, cache_stride, cache_offset, resampled,
            resync_every, c->step_tile_off, c->d_phase, C, sched, chunk, c->d_err, c->spin_limit);
    else
        run_kernel<<<grid, GMC_STEP_THREADS, c->step_smem_bytes, st>>>(
            c->dev, bed, mcres, ssq, seeds, iter0, n_steps, loss_cache, step_cache, blocks_cache, cache_stride, cache_offset, resampled,
            resync_every, c->step_tile_off, c->d_phase, C, sched, chunk, c->d_err, c->spin_limit);
    if (sched) gmc_sched_release(c, st);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}

// ---- debug: per-phase cycle accounting of run_kernel ---------------------------------------------------------
extern "C" int gmc_debug_phase_timing(gmc_ctx* c, int enable, int64_t* cycles_out) {
    if (!c) GMC_FAIL(GMC_EINVAL, "gmc_debug_phase_timing: ctx is NULL");
    GMC_CUDA(cudaSetDevice(c->device));
    if (cycles_out && c->d_phase) {
        GMC_CUDA(cudaDeviceSynchronize());
        GMC_CUDA(cudaMemcpy(cycles_out, c->d_phase, GMC_N_PHASES * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    if (enable && !c->d_phase) GMC_CUDA(cudaMalloc(&c->d_phase, GMC_N_PHASES * sizeof(long long)));
    if (enable) GMC_CUDA(cudaMemset(c->d_phase, 0, GMC_N_PHASES * sizeof(long long)));
    if (!enable && c->d_phase) {
        cudaFree(c->d_phase);
        c->d_phase = nullptr;
    }
    return GMC_OK;
}
