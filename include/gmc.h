/* gmc.h — C ABI of libgmc.so: the B200 (sm_100a) many-chain MCMC step of gstatsMCMC.
 *
 * The reference (tylerrleee/mcmc-gpu, pure Python) has no FFI layer; its boundary for this path is the Python API
 * of gstatsMCMC/MCMC.py and gstatsMCMC/Topography.py.  Each entry point below names the reference code it replaces
 * (file:line into the reference tree).  The Python mirror of that API (mcmc_gpu_b200/MCMC.py, Topography.py) binds
 * these symbols with ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 (GMC_OK) or a negative gmc_status; gmc_last_error() gives the thread-local message;
 *     nothing throws across the boundary.
 *   - arrays are float64, row-major, chains outermost: bed[C][H][W]; masks are uint8 holding 0/1.
 *   - "dev" pointers are CUDA device pointers owned by the caller (torch.Tensor.data_ptr()).  "any" pointers may be
 *     host or device (copied with cudaMemcpyDefault during setup calls).
 *   - all compute calls are asynchronous on the cudaStream_t passed as `stream` (void*; NULL = legacy default).
 *   - a context is bound to one device and is not thread-safe: one context per GPU per process.
 *   - there is no CPU fallback: without a usable CUDA device every call fails with GMC_ECUDA.
 */
#ifndef GMC_H_
#define GMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GMC_API __attribute__((visibility("default")))
#else
#define GMC_API
#endif

typedef struct gmc_ctx gmc_ctx;

typedef enum gmc_status {
    GMC_OK = 0,
    GMC_EINVAL = -1,       /* bad argument value / NULL pointer                          */
    GMC_ESHAPE = -2,       /* shape mismatch (chains > capacity, block larger than grid) */
    GMC_ECUDA = -3,        /* CUDA runtime error, or no device                           */
    GMC_ENCCL = -4,        /* NCCL error or libnccl not loadable                         */
    GMC_EUNSUPPORTED = -5, /* valid in the reference but not implemented here            */
    GMC_ESTATE = -6        /* call order: set_static / set_field_model / set_blocks first */
} gmc_status;

/* covariance model of the proposal field: RandField.model_name, MCMC.py:496-500 */
typedef enum gmc_model { GMC_GAUSSIAN = 0, GMC_EXPONENTIAL = 1, GMC_MATERN = 2 } gmc_model;

GMC_API const char* gmc_last_error(void);
GMC_API int gmc_version(void);

/* ---- context ------------------------------------------------------------------------------------------------ */

/* One context = one grid shape on one device, shared by up to max_chains chains.
 * Replaces the per-process chain object state of init_lsc_chain_by_instance (MCMC.py:359-379). */
GMC_API int gmc_create(gmc_ctx** out, int device, int H, int W, int max_chains);
GMC_API int gmc_destroy(gmc_ctx* ctx);

/* Chain-independent inputs (identical for all chains in the reference too: largeScaleChain_multiprocessing.py:51-57).
 *   surf, velx, vely, dhdt, smb : [H][W] f64 (any)               chain.__init__, MCMC.py:808-845
 *   gate_mask  [H][W] u8 (any)  : region_mask when update_in_region else grounded_ice_mask; cells a proposal may
 *                                 change, guard cells and resampled_times increment   MCMC.py:1287-1290,1324-1329,1349-1352
 *   mc_mask    [H][W] u8 (any)  : mc_region_mask of set_loss_type                    MCMC.py:1004-1007,1041
 *   centre_cells [n] i32 (any)  : linear indices (row*W+col) the block centre is drawn from (cells with
 *                                 region_mask==1); NULL/0 = whole grid                MCMC.py:1253-1261
 *   crf_weight [H][W] f64 (any) : set_crf_data_weight; NULL selects block_type 'RF'  MCMC.py:1124-1134,1279-1282
 *   resolution, sigma_mc        : chain.resolution, set_loss_type(sigma_mc)           MCMC.py:1016 */
GMC_API int gmc_set_static(gmc_ctx* ctx, const double* surf, const double* velx, const double* vely, const double* dhdt,
                   const double* smb, const uint8_t* gate_mask, const uint8_t* mc_mask, const int32_t* centre_cells,
                   int64_t n_centre_cells, const double* crf_weight, double resolution, double sigma_mc);

/* RandField.__init__ parameters, MCMC.py:462-510.  smoothness is ignored unless model == GMC_MATERN. */
GMC_API int gmc_set_field_model(gmc_ctx* ctx, int model, double smoothness, int isotropic, double range_min_x,
                        double range_max_x, double range_min_y, double range_max_y, double scale_min,
                        double scale_max, double nugget_max);

/* RandField.set_block_sizes / get_edge_masks, MCMC.py:524-623.
 *   pair_w[i], pair_h[i] : RandField.pairs[0][i], pairs[1][i] (even, >= 2); the field of pair i is [pair_h][pair_w]
 *   edge_masks (any)     : the n_pairs tapers concatenated, pair i at offsets[i], row-major [pair_h][pair_w]
 *   field_resolution     : RandField.resolution (spacing of the fftfreq grid, MCMC.py:221-222) */
GMC_API int gmc_set_blocks(gmc_ctx* ctx, int n_pairs, const int32_t* pair_w, const int32_t* pair_h, const double* edge_masks,
                   const int64_t* offsets, double field_resolution);

/* ---- A1/A2: residual and loss ------------------------------------------------------------------------------- */

/* Topography.get_mass_conservation_residual (Topography.py:592-600) for C beds at once.
 * bed, res_out: dev [C][H][W].  Needs gmc_set_static. */
GMC_API int gmc_residual(gmc_ctx* ctx, const double* bed, double* res_out, int C, void* stream);

/* Fused residual + chain.loss (Topography.py:592-600 + MCMC.py:1041): res_out may be NULL.
 * loss_out: dev [C] = nansum(res[mc_mask==1]^2) / (2 sigma_mc^2); ssq_out (dev [C], may be NULL) = the nansum. */
GMC_API int gmc_residual_loss(gmc_ctx* ctx, const double* bed, double* res_out, double* loss_out, double* ssq_out, int C,
                      void* stream);

/* Same for the chain range [chain0, chain0+C) of the context (pointers already point at that range): calls for disjoint
 * ranges use disjoint workspace and may run concurrently on different streams (pipelined host<->device transfers). */
GMC_API int gmc_residual_loss_range(gmc_ctx* ctx, const double* bed, double* res_out, double* loss_out, double* ssq_out,
                                    int C, int chain0, void* stream);

/* chain.loss (MCMC.py:1021-1044) on given residuals: res dev [C][H][W] -> loss_out dev [C] (ssq_out optional). */
GMC_API int gmc_loss(gmc_ctx* ctx, const double* res, double* loss_out, double* ssq_out, int C, void* stream);

/* ---- A3/A4: proposal field ---------------------------------------------------------------------------------- */

/* spectral_synthesis_field (+ edge taper of get_rfblock when apply_taper), MCMC.py:176-254, 742-778, for n fields.
 *   pair[n] i32 dev        : block-size index of each field
 *   scale[n], nug[n], range_x[n], range_y[n] f64 dev : the sampled parameters (scale already divided by 3)
 *   z_re, z_im, z_nug      : dev [n][stride] unit normals laid out [pair_h][pair_w] per field, or all NULL to draw
 *                            them from Philox streams keyed by seeds[n] (u64 dev) at iteration `iter`
 *   f_out                  : dev [n][stride], field i written row-major [pair_h][pair_w] at f_out + i*stride */
GMC_API int gmc_field_spectral(gmc_ctx* ctx, int n, const int32_t* pair, const double* scale, const double* nug,
                       const double* range_x, const double* range_y, const double* z_re, const double* z_im,
                       const double* z_nug, const uint64_t* seeds, uint64_t iter, int apply_taper, double* f_out,
                       int64_t stride, void* stream);

/* RandField.set_generation_method (MCMC.py:514-522).  spectral != 0 (default): gmc_run proposes with the FFT synthesis
 * (A3).  spectral == 0: with the randomization method (A5, below) using n_modes wave vectors per field (gstools SRF
 * default mode_no = 1000). */
GMC_API int gmc_set_generation_method(gmc_ctx* ctx, int spectral, int n_modes);

/* A5 — RandField.get_random_field (MCMC.py:625-687: gstools.SRF(model).structured([X, Y]).T * scale), n fields:
 *   field[y][x] = scale * ( sqrt(1/n_modes) * sum_m [ z1_m cos(kx_m x res + ky_m y res) + z2_m sin(..) ] + sqrt(nug) n[y][x] )
 * (times the edge taper of get_rfblock when apply_taper).
 *   pair[n] i32 dev : block-size index;  scale[n], nug[n], range_x[n], range_y[n], angle_deg[n] f64 dev : the sampled
 *                     parameters (scale already / 3; gstools len_scale = range / sqrt(3) | 3 | 2 by model, angle in degrees)
 *   modes           : dev [n][n_modes][4] = (kx, ky, z1, z2), wave vectors in the grid frame (rad per unit of the
 *                     field resolution), with z_nug dev [n][stride] unit normals — or both NULL: wave vectors are drawn
 *                     on the device from the model's radial spectral distribution (closed-form inversion in 2-D, see
 *                     DESIGN.md) and Philox streams keyed by seeds[n] (u64 dev) at iteration `iter`
 *   f_out           : dev [n][stride], field i row-major [pair_h][pair_w] at f_out + i*stride
 * gstools (1.7.0 in the reference environment) is not vendored and its RNG is unseeded in the reference: parity of
 * this branch is unpinned; the summation is checked against oracle/randmeth_oracle.py, the sampling distributionally. */
GMC_API int gmc_field_randmeth(gmc_ctx* ctx, int n, const int32_t* pair, const double* scale, const double* nug,
                       const double* range_x, const double* range_y, const double* angle_deg, int n_modes,
                       const double* modes, const double* z_nug, const uint64_t* seeds, uint64_t iter, int apply_taper,
                       double* f_out, int64_t stride, void* stream);

/* ---- A6: Metropolis step ------------------------------------------------------------------------------------ */

/* One chain_crf.run loop body (MCMC.py:1263-1360) for C chains with the proposal injected:
 *   bed, mcres : dev [C][H][W] state (updated in place on accept);  ssq : dev [C] nansum state (loss*2 sigma^2)
 *   f          : dev [C][f_stride], chain c's tapered field [h][w] row-major at f + c*f_stride (the value
 *                RandField.get_rfblock returns);  hw : dev [C][2] i32 = f.shape;  centre : dev [C][2] i32 = (indexx, indexy)
 *   u          : dev [C] uniforms of MCMC.py:1336;  hmax, wmax : upper bounds of hw (size the shared-memory tile)
 *   accepted_out : dev [C] u8;  loss_out : dev [C] loss after the decision;  loss_next_out : dev [C] candidate loss (may be NULL)
 *   resampled  : dev [C][H][W] i32 accepted-block coverage counts, or NULL (MCMC.py:1349-1352, times the mask value) */
GMC_API int gmc_step_injected(gmc_ctx* ctx, double* bed, double* mcres, double* ssq, const double* f, int64_t f_stride,
                      const int32_t* hw, const int32_t* centre, const double* u, int hmax, int wmax,
                      uint8_t* accepted_out, double* loss_out, double* loss_next_out, int32_t* resampled, int C,
                      void* stream);

/* n_steps fused free-running iterations for C chains (proposal synthesis + step on device, no host round trips).
 *   seeds  : dev [C] u64 per-chain Philox keys (the reference's rng_seeds, largeScaleChain_multiprocessing.py:55-57)
 *   iter0  : index of the first iteration (Philox counter; makes runs resumable and independent of GPU count)
 *   loss_cache, step_cache (f64 / u8, dev [C][cache_stride]) and blocks_cache (i32 dev [C][cache_stride][4] =
 *   indexx, indexy, h, w) receive entries [iter0 - cache_iter0 + k] for step k; any may be NULL      MCMC.py:1163-1170
 *   resync_every : recompute ssq from the tracked mcres every this many iterations (0 = never)      */
GMC_API int gmc_run(gmc_ctx* ctx, double* bed, double* mcres, double* ssq, const uint64_t* seeds, uint64_t iter0, int n_steps,
            double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache, int64_t cache_stride, int64_t cache_offset,
            int32_t* resampled, int resync_every, int C, void* stream);

/* ---- S1-S5: small-scale chain (block re-simulation by SGS) ---------------------------------------------------- */

/* Tables of the SGS chain (chain_sgs setters, MCMC.py:1465-1597).  All pointers "any" (host or device), copied.
 *   trend [H][W] or NULL          : set_trend (MCMC.py:1481-1496); the chain state is bed - trend
 *   zcond [H][W]                  : normal-scored conditioning data, NaN where none (MCMC.py:1653-1661)
 *   grounded [H][W] u8            : grounded_ice_mask of the full-grid thickness guard (MCMC.py:1789-1795)
 *   quantiles, references [n_q]   : QuantileTransformer.quantiles_[:,0], references_ (NULL/NULL: do_transform=False)
 *   oct_off [8][lmax][2] i16      : for octant b = -4..3 of neighbors.py:52-60 the window offsets (di, dj) with distance <
 *                                   the WIDEST search radius, sorted by (distance, di, dj); hw = window half width in cells
 *   oct_cnt [n_levels][8], n_levels: prefix lengths of those lists for radius, radius + 100 km, ...: a node that finds no
 *                                   conditioned cell within `radius` searches again 100 km wider (MCMC.py:149-155)
 *   num_points                    : set_sgs_param neighbours (num_points//8 per octant)
 *   lut [(4hw+1)][(4hw+1)]        : covariance of the offset (di, dj), di,dj in [-2hw, 2hw] (covariance.py models)
 *   sill; block sizes             : variogram sill; set_block_sizes (sizes drawn from [min, max), MCMC.py:1755-1756) */
GMC_API int gmc_sgs_setup(gmc_ctx* ctx, const double* trend, const double* zcond, const uint8_t* grounded,
                          const double* quantiles, const double* references, int n_quantiles, const int16_t* oct_off,
                          const int32_t* oct_cnt, int n_levels, int lmax, int hw, int num_points, const double* lut, double sill,
                          int block_min_x, int block_max_x, int block_min_y, int block_max_y);

/* QuantileTransformer(output_distribution='normal').transform / inverse_transform of n values (dev), S5. */
GMC_API int gmc_sgs_transform(gmc_ctx* ctx, const double* in, double* out, int64_t n, int inverse, void* stream);

/* Chain state from C full beds (dev [C][H][W]): bedc = bed - trend, z = normal score of bedc, mcres/ssq = residual and
 * nansum of bedc + trend, nviol[C] = cells violating the thickness guard.  scratch_full: dev [C][H][W] work array.
 * MCMC.py:1636-1669. */
GMC_API int gmc_sgs_init(gmc_ctx* ctx, const double* bed, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol,
                         double* scratch_full, int C, void* stream);

/* One chain_sgs.run loop body (MCMC.py:1747-1829) for C chains with every random input injected:
 *   centre [C][2] i32 (indexx, indexy); block_size [C][2] i32; path [C][path_stride] i32 = the shuffled visiting order
 *   as block-local cell indices (row * block_width + col); znorm [C][path_stride] unit normals by path position; u [C].
 *   err_flag (dev i32, may be NULL) is OR-ed with 1 if a node found no neighbour inside the search radius. */
GMC_API int gmc_sgs_step_injected(gmc_ctx* ctx, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol,
                                  const int32_t* centre, const int32_t* block_size, const int32_t* path, const double* znorm,
                                  int64_t path_stride, const double* u, uint8_t* accepted_out, double* loss_out,
                                  double* loss_next_out, int32_t* resampled, int32_t* err_flag, int C, void* stream);

/* n_steps free-running SGS iterations for C chains (device Philox: centre, block size, path permutation, normals, u). */
GMC_API int gmc_sgs_run(gmc_ctx* ctx, double* bedc, double* z, double* mcres, double* ssq, int32_t* nviol, const uint64_t* seeds,
                        uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache,
                        int64_t cache_stride, int64_t cache_offset, int32_t* resampled, int32_t* err_flag, int C, void* stream);

/* ---- whole-grid SGS: the initial beds of the large-scale chains (gstatsim_custom/interpolate.py:92-191) ---------- */

/* QuantileTransformer(output_distribution='normal') transform / inverse_transform from explicit tables (dev [nq]), as
 * interpolate.sgs applies them to the data, the bounds and the result (utilities.py:21-24, interpolate.py:240-246, 185). */
GMC_API int gmc_nst_transform(int device, const double* quantiles, const double* references, int n_quantiles, const double* in,
                      double* out, int64_t n, int inverse, void* stream);

/* Kriging records of n_real realisations (the loop body of interpolate.py:130-163 without the draw), all nodes in parallel:
 *   ord  dev [n_real][H*W] i32 : -1 at conditioning cells, else the cell's position in the realisation's path
 *   path dev [n_real][n_path] i32 : row-major cell indices in simulation order (the shuffled `inds`, :125)
 *   oct_off / lmax / hw : octant search lists as for gmc_sgs_setup (neighbors.py:52-60), built for the WIDEST radius;
 *   oct_cnt dev [n_levels][8] : per radius level (radius, radius + 100 km, ...: the reference widens the search of a node
 *   that finds no data, interpolate.py:149-155) the number of list entries closer than that radius;  n_levels <= 4
 *   lut : covariance of integer offsets, [(4hw+1)^2];  num_points <= 48
 *   rec_n [n_real][n_path] i32 (-1: conditioning cell), rec_idx, rec_w [n_real][n_path][48], rec_sd [n_real][n_path]
 *   err_flag dev i32: bit 0 set if a node found no neighbour even within the widest radius. */
GMC_API int gmc_sgs_grid_solve(int device, int H, int W, const int32_t* ord, const int32_t* path, int64_t n_path, int n_real,
                       const int16_t* oct_off, const int32_t* oct_cnt, int n_levels, int lmax, int hw, int num_points,
                       const double* lut, double sill, int32_t* rec_n, int32_t* rec_idx, double* rec_w, double* rec_sd, int32_t* err_flag,
                       void* stream);

/* The draws in path order (interpolate.py:166-183): z dev [n_real][H*W] holds the normal-scored data (anything elsewhere: the cells to simulate are
 * first marked NaN, the readiness marker of the dependency-driven walk - thirty-two warps per realisation)
 * and receives the simulated normal scores; noise dev [n_real][n_path]: the standard normal of node t (no bounds) or
 * the uniform of its truncated-normal draw; bound_lo / bound_hi dev [H*W] normal-scored bounds, or both NULL. */
GMC_API int gmc_sgs_grid_values(int device, int H, int W, double* z, const int32_t* path, int64_t n_path, int n_real,
                        const int32_t* rec_n, const int32_t* rec_idx, const double* rec_w, const double* rec_sd,
                        const double* noise, const double* bound_lo, const double* bound_hi, void* stream);

/* ---- ensemble statistics (new) ---------------------------------------------------------------- */

/* Local part of the posterior mean/variance: sum_out[H][W] = sum_c (bed_c - ref), sumsq_out = sum_c (bed_c - ref)^2. */
GMC_API int gmc_ensemble_moments(gmc_ctx* ctx, const double* bed, const double* ref_bed, double* sum_out, double* sumsq_out,
                         int C, void* stream);

/* In-place SUM all-reduce of the two [H][W] moment arrays and count[1] over an existing ncclComm_t (the only
 * collective on the path).  libnccl.so.2 is resolved at run time from the process (torch ships it). */
GMC_API int gmc_allreduce_moments(gmc_ctx* ctx, void* nccl_comm, double* sum, double* sumsq, double* count, void* stream);

/* ---- setup helper ------------------------------------------------------------------------------------------------ */

/* Utilities.min_dist_from_mask (Utilities.py:21-24): out[q] = min_p sqrt((qx-px)^2 + (qy-py)^2), bit-identical to the
 * reference's KD-tree query.  All pointers dev.  Used for the block tapers (MCMC.py:583-623) and the conditioning weight
 * (MCMC.py:689-714).  Brute force, O(N*M). */
GMC_API int gmc_min_dist(int device, const double* px, const double* py, int64_t M, const double* qx, const double* qy,
                         int64_t N, double* out, void* stream);

/* PIL ImageFilter.ModeFilter(size) as Topography.get_highvel_boundary applies it to the binary region mask
 * (Topography.py:551-553): in/out dev [H][W] u8 with values 0 / 255 (any non-zero input counts as 255); window
 * (2*(size/2)+1)^2 clipped to the image, majority wins, 0 on ties. */
GMC_API int gmc_mode_filter_binary(int device, const uint8_t* in, uint8_t* out, int H, int W, int size, void* stream);

/* ---- introspection for tests and bench ---------------------------------------------------------------------- */

/* Number of kernel launches issued through this context since creation. */
GMC_API int64_t gmc_launch_count(const gmc_ctx* ctx);
/* Dynamic shared memory (bytes) and threads per CTA of the fused step kernel for the current block table. */
GMC_API int gmc_step_kernel_info(const gmc_ctx* ctx, int* smem_bytes, int* threads, int* ctas_per_sm);

/* Launch shape policy of gmc_run's step kernels: 0 = auto (a launch with no more chains than the GPU has SMs gets 512-thread
 * CTAs, one per SM, so a chain uses all the warps of its SM; otherwise 256-thread CTAs, two per SM), 1 = always 256 threads
 * (for launches that share the GPU with other launches: chain ranges on several streams, several steps in flight),
 * 2 = 512 threads whenever the launch is not chunk-scheduled, 3 = SPLIT when fewer than half as many chains as CTA slots
 * (experimental: per chain a producer CTA that synthesises the proposal fields of the coming steps and a consumer CTA that
 * runs the Metropolis tail, pipelined across steps; measured slower than mode 2).  Trajectories do not depend
 * on the choice (bit-identical). */
GMC_API int gmc_set_step_cta(gmc_ctx* ctx, int mode);

/* Reports - then clears - the device-error flag (synchronize != 0: after waiting for the device; 0: as stored by the
 * launches that have finished, for callers that synchronised their own streams): GMC_ECUDA when a kernel of an earlier launch
 * gave up one of its bounded in-kernel waits (a (chunk, chain) item waiting for its chain's previous chunk, or a tile copy
 * that never completed) instead of running on with stale state; gmc_run / gmc_sgs_run also refuse to start while the flag
 * is set.  The reference has no counterpart (its chains are OS processes, largeScaleChain_multiprocessing.py:78-79); this
 * is the error path of the (chunk, chain) scheduler that replaces them. */
GMC_API int gmc_check(gmc_ctx* ctx, int synchronize);

/* Measures the FP64 FMA rate of the device with a register-resident DFMA loop (TFLOP/s, 2 flops per FMA): the
 * denominator of the FP64-pipe fractions quoted for the step and kriging kernels. */
GMC_API int gmc_debug_fp64_peak(gmc_ctx* ctx, double* tflops_out);

/* Debug: per-phase SM-cycle accounting of the fused step kernel (thread 0 of every CTA, summed over CTAs and steps).
 * enable != 0 allocates/zeroes the counters, 0 releases them; cycles_out (host, 8 x int64, may be NULL) receives the
 * counters accumulated so far: 0 scalars+prefetch, 1 spectrum fill, 2 column DFT, 3 row recombination, 4 row DFT,
 * 5 candidate tile, 6 block residual+loss+decision, 7 write-back. */
GMC_API int gmc_debug_phase_timing(gmc_ctx* ctx, int enable, int64_t* cycles_out);

/* Debug/test: compares the stencil kernels' division-by-constant (reciprocal + FMA residual correction) with the
 * correctly rounded division for n dividends x (dev) and one divisor; *mismatches_out = number of differing bit patterns. */
GMC_API int gmc_debug_div_check(gmc_ctx* ctx, const double* x, int64_t n, double divisor, int64_t* mismatches_out);

#ifdef __cplusplus
}
#endif
#endif /* GMC_H_ */
