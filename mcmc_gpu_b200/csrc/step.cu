// step.cu — K1 (proposal field synthesis) and K4 (Metropolis step) of the large-scale chain, fused.
//
// One CTA owns one chain.  Per iteration everything lives in shared memory: the h x w complex spectrum is filled
// from counter-based Philox normals, inverse-transformed in place (mixed-radix decimation in time, digit-reversed
// load), standardised and tapered into the proposal f (MCMC.py:176-254, 742-778); the candidate bed of the clipped
// block plus a one-cell halo is staged as a tile, the residual is recomputed on the block only, the loss changes by
// the block's delta, and the accept/reject decision plus in-place write-back happen without leaving the kernel
// (MCMC.py:1263-1360).  HBM sees the bed halo tile, the old block residual and, on accept, the block write-back.
#include "common.cuh"

struct StepScalars {
    int pair, h, w, ix, iy;
    int x0, x1, y0, y1, mx0, my0;
    double scale, nug, range_x, range_y, u;
    int accept;
};

// ---------------------------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }   // a * (+i)

// inverse-direction (e^{+i...}) DFT cores, in place on v[0..R)
__device__ __forceinline__ void dft2(double2* v) {
    const double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
__device__ __forceinline__ void dft4(double2* v) {
    const double2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    const double2 t2 = cadd(v[1], v[3]), t3 = cmuli(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
__device__ __forceinline__ void dft8(double2* v) {
    double2 e[4] = {v[0], v[2], v[4], v[6]};
    double2 o[4] = {v[1], v[3], v[5], v[7]};
    dft4(e);
    dft4(o);
    const double s = 0.70710678118654752440;
    o[1] = make_double2((o[1].x - o[1].y) * s, (o[1].x + o[1].y) * s);     // * e^{+i pi/4}
    o[2] = cmuli(o[2]);                                                    // * e^{+i pi/2}
    o[3] = make_double2((-o[3].x - o[3].y) * s, (o[3].x - o[3].y) * s);    // * e^{+i 3pi/4}
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = cadd(e[k], o[k]);
        v[k + 4] = csub(e[k], o[k]);
    }
}
// odd prime radix R: out[k] = sum_q v[q] W^{qk}, W = e^{+2 pi i / R}, using the q <-> R-q symmetry:
//   out[k], out[R-k] = v0 + sum_q cos(2 pi qk/R) (v[q]+v[R-q])  +/-  i sum_q sin(2 pi qk/R) (v[q]-v[R-q])
template <int R>
__device__ __forceinline__ void dft_odd(double2* v, const double2* __restrict__ tw, int step_r) {
    constexpr int HALF = (R - 1) / 2;
    double2 w[HALF + 1];
#pragma unroll
    for (int m = 1; m <= HALF; ++m) w[m] = __ldg(tw + m * step_r);
    double2 a[HALF + 1], b[HALF + 1];
    double2 sum0 = v[0];
#pragma unroll
    for (int q = 1; q <= HALF; ++q) {
        a[q] = cadd(v[q], v[R - q]);
        b[q] = csub(v[q], v[R - q]);
        sum0 = cadd(sum0, a[q]);
    }
    const double2 v0 = v[0];
    v[0] = sum0;
#pragma unroll
    for (int k = 1; k <= HALF; ++k) {
        double2 re = v0, im = make_double2(0.0, 0.0);
#pragma unroll
        for (int q = 1; q <= HALF; ++q) {
            const int m = (q * k) % R;                       // compile-time after unrolling
            const double c = (m <= HALF) ? w[m].x : w[R - m].x;
            const double sn = (m <= HALF) ? w[m].y : -w[R - m].y;
            re.x += c * a[q].x;
            re.y += c * a[q].y;
            im.x += sn * b[q].x;
            im.y += sn * b[q].y;
        }
        v[k] = make_double2(re.x - im.y, re.y + im.x);       // re + i*im
        v[R - k] = make_double2(re.x + im.y, re.y - im.x);   // re - i*im
    }
}

// One radix-R stage over `count` independent lines.  ALONG_ROW: the transform runs along x (contiguous) and lanes
// map to different rows (pitch is odd in double2 units => conflict-free); otherwise it runs along y and lanes map to
// adjacent columns.
template <int R, bool ALONG_ROW>
__device__ __forceinline__ void fft_stage(double2* Z, int pitch, int n, int count, int L, const double2* __restrict__ tw) {
    const int M = L / R;
    const int step = n / L;
    const int items = (n / R) * count;
    for (int t = threadIdx.x; t < items; t += GMC_STEP_THREADS) {
        const int line = t % count;
        const int bf = t / count;
        const int blk = bf / M, k1 = bf - blk * M;
        const int base = blk * L + k1;
        double2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int p = base + q * M;
            v[q] = ALONG_ROW ? Z[line * pitch + p] : Z[p * pitch + line];
        }
        if (L > R) {   // first stage has all twiddles == 1
            const int i1 = k1 * step;
            int iq = i1;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                v[q] = cmul(v[q], __ldg(tw + iq));
                iq += i1;
                if (iq >= n) iq -= n;
            }
        }
        if (R == 2) dft2(v);
        else if (R == 4) dft4(v);
        else if (R == 8) dft8(v);
        else dft_odd<R>(v, tw, n / R);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int p = base + q * M;
            if (ALONG_ROW) Z[line * pitch + p] = v[q];
            else Z[p * pitch + line] = v[q];
        }
    }
}

// Fallback for any other prime radix R <= GMC_MAX_RADIX (block edges such as 58 = 2*29): O(R^2) per butterfly with the
// inputs parked in local memory.  Slow but rare; the default block sizes never take it.
template <bool ALONG_ROW>
__device__ __noinline__ void fft_stage_generic(double2* Z, int pitch, int n, int count, int L, int R,
                                               const double2* __restrict__ tw) {
    const int M = L / R;
    const int step = n / L;
    const int step_r = n / R;
    const int items = (n / R) * count;
    double2 v[GMC_MAX_RADIX];
    for (int t = threadIdx.x; t < items; t += GMC_STEP_THREADS) {
        const int line = t % count;
        const int bf = t / count;
        const int blk = bf / M, k1 = bf - blk * M;
        const int base = blk * L + k1;
        const int i1 = k1 * step;
        int iq = 0;
        for (int q = 0; q < R; ++q) {
            const int p = base + q * M;
            const double2 x = ALONG_ROW ? Z[line * pitch + p] : Z[p * pitch + line];
            v[q] = (q == 0 || L == R) ? x : cmul(x, __ldg(tw + iq));
            iq += i1;
            if (iq >= n) iq -= n;
        }
        for (int k = 0; k < R; ++k) {
            double2 acc = v[0];
            int m = 0;
            for (int q = 1; q < R; ++q) {
                m += k;
                if (m >= R) m -= R;
                acc = cadd(acc, cmul(v[q], __ldg(tw + m * step_r)));
            }
            const int p = base + k * M;
            if (ALONG_ROW) Z[line * pitch + p] = acc;
            else Z[p * pitch + line] = acc;
        }
    }
}

template <bool ALONG_ROW>
__device__ void fft_lines(double2* Z, int pitch, const GmcFftPlan& plan, int count, const double2* __restrict__ tw_all) {
    const double2* tw = tw_all + plan.tw_off;
    int L = 1;
    for (int s = 0; s < plan.n_factors; ++s) {
        const int r = plan.radix[s];
        L *= r;
        switch (r) {
            case 2: fft_stage<2, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 3: fft_stage<3, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 4: fft_stage<4, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 5: fft_stage<5, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 7: fft_stage<7, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 8: fft_stage<8, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 11: fft_stage<11, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            case 13: fft_stage<13, ALONG_ROW>(Z, pitch, plan.n, count, L, tw); break;
            default: fft_stage_generic<ALONG_ROW>(Z, pitch, plan.n, count, L, r, tw); break;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// spectral amplitude sqrt(S(k))                                                          MCMC.py:209-239
// ---------------------------------------------------------------------------------------------------------------
struct SpecParams {
    int model;
    double a;        // sqrt(len_x*len_y)
    double nu;
    double constant; // Matern prefactor
    double kappa;    // 2 nu / a^2
};

__device__ __forceinline__ SpecParams make_spec(const GmcFieldModel& fm, double range_x, double range_y) {
    SpecParams sp;
    sp.model = fm.model;
    double len_x, len_y;
    if (fm.model == GMC_GAUSSIAN) {
        const double s3 = sqrt(3.0);
        len_x = div_rn(range_x, s3);
        len_y = div_rn(range_y, s3);
    } else if (fm.model == GMC_EXPONENTIAL) {
        len_x = div_rn(range_x, 3.0);
        len_y = div_rn(range_y, 3.0);
    } else {
        len_x = div_rn(range_x, 2.0);
        len_y = div_rn(range_y, 2.0);
    }
    sp.a = sqrt(mul_rn(len_x, len_y));
    sp.nu = fm.smoothness;
    sp.constant = 1.0;
    sp.kappa = 0.0;
    if (fm.model == GMC_MATERN) {
        sp.constant = div_rn(fm.matern_num, mul_rn(fm.matern_gamma, pow(sp.a, mul_rn(2.0, sp.nu))));
        sp.kappa = div_rn(mul_rn(2.0, sp.nu), mul_rn(sp.a, sp.a));
    }
    return sp;
}

__device__ __forceinline__ double spec_amp(const SpecParams& sp, double ksq_sum) {
    const double k = add_rn(sqrt(ksq_sum), 1e-10);
    double S;
    if (sp.model == GMC_GAUSSIAN) {
        const double ak = mul_rn(sp.a, k);
        S = exp(mul_rn(-0.5, mul_rn(ak, ak)));
    } else if (sp.model == GMC_EXPONENTIAL) {
        const double ak = mul_rn(sp.a, k);
        S = div_rn(1.0, pow(add_rn(1.0, mul_rn(ak, ak)), 1.5));
    } else {
        const double four_pi = 4 * 3.141592653589793;
        S = mul_rn(sp.constant, pow(add_rn(sp.kappa, mul_rn(four_pi, mul_rn(k, k))), sub_rn(-sp.nu, 1.0)));
    }
    return sqrt(S);
}

// ---------------------------------------------------------------------------------------------------------------
// K1: synthesise one field into shared memory.  On return (after its final __syncthreads) buf[0 .. h*w) holds
// f[h][w] row-major (tapered when apply_taper).
// ---------------------------------------------------------------------------------------------------------------
template <bool INJECT>
__device__ void synth_field(const GmcDev& d, double* buf, double* scratch, int pair_idx, double scale, double nug,
                            double range_x, double range_y, const Philox& rng, uint32_t it_lo, uint32_t it_hi,
                            const double* __restrict__ z_re, const double* __restrict__ z_im,
                            const double* __restrict__ z_nug, bool apply_taper) {
    __shared__ GmcFftPlan s_plan[2];
    const GmcPair pr = d.pairs[pair_idx];
    const int h = pr.h, w = pr.w;
    if (threadIdx.x < 2) s_plan[threadIdx.x] = d.plans[threadIdx.x == 0 ? pr.plan_h : pr.plan_w];
    __syncthreads();
    const GmcFftPlan& ph = s_plan[0];
    const GmcFftPlan& pw = s_plan[1];
    const int pitch = w + 1;
    double2* Z = reinterpret_cast<double2*>(buf);
    const SpecParams sp = make_spec(d.fm, range_x, range_y);
    const int16_t* posY = d.pos + ph.pos_off;
    const int16_t* posX = d.pos + pw.pos_off;
    const double* ksqY = d.ksq + ph.ksq_off;
    const double* ksqX = d.ksq + pw.ksq_off;

    // (1) fill the spectrum: one item per (|ky|, |kx|) class, up to four mirrored entries share sqrt(S)
    const int hq = h / 2 + 1, wq = w / 2 + 1;
    for (int q = threadIdx.x; q < hq * wq; q += GMC_STEP_THREADS) {
        const int a = q / wq, b = q - a * wq;
        // MCMC.py:224: k = sqrt(kxv**2 + kyv**2) + 1e-10
        const double amp = spec_amp(sp, add_rn(__ldg(ksqX + b), __ldg(ksqY + a)));
        const int na = (a == 0 || a == h / 2) ? 1 : 2;
        const int nb = (b == 0 || b == w / 2) ? 1 : 2;
        for (int sa = 0; sa < na; ++sa) {
            const int ky = sa ? h - a : a;
            for (int sb = 0; sb < nb; ++sb) {
                const int kx = sb ? w - b : b;
                const int e = ky * w + kx;
                double zr, zi;
                if (INJECT) {
                    zr = z_re[e];
                    zi = z_im[e];
                } else {
                    box_muller(rng((uint32_t)e, it_lo, it_hi, GMC_STREAM_NOISE), zr, zi);
                }
                Z[__ldg(posY + ky) * pitch + __ldg(posX + kx)] = make_double2(zr * amp, zi * amp);
            }
        }
    }
    __syncthreads();

    // (2) inverse 2-D DFT in place: last axis first like numpy's ifft2
    fft_lines<true>(Z, pitch, pw, h, d.twiddle);
    fft_lines<false>(Z, pitch, ph, w, d.twiddle);

    // (3) standardise: (x - mean) / (std + 1e-12), population std                       MCMC.py:247-248
    const int n = h * w;
    const double inv_n = 1.0 / (double)n;
    double acc = 0.0;
    for (int e = threadIdx.x; e < n; e += GMC_STEP_THREADS) {
        const int y = e / w, x = e - y * w;
        acc += Z[y * pitch + x].x;
    }
    const double mean = block_sum<GMC_STEP_THREADS>(acc, scratch) * inv_n * inv_n;   // includes the 1/(h w) of ifft2
    acc = 0.0;
    for (int e = threadIdx.x; e < n; e += GMC_STEP_THREADS) {
        const int y = e / w, x = e - y * w;
        const double dv = Z[y * pitch + x].x * inv_n - mean;
        acc += dv * dv;
    }
    const double sd = sqrt(block_sum<GMC_STEP_THREADS>(acc, scratch) * inv_n);
    const double inv_sd = 1.0 / (sd + 1e-12);
    const double sq_nug = sqrt(nug);
    const double* taper = d.edge_masks + pr.mask_off;

    // (4) scale, nugget, taper, and compact the real parts to buf[0..n) in ascending waves (in-place safe: the
    // destination of element e only overlaps sources of elements <= e/2, all read in this or an earlier wave)
    constexpr int CH = 8;
    for (int wave = 0; wave * CH * GMC_STEP_THREADS < n; ++wave) {
        double v[CH];
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int e = (wave * CH + k) * GMC_STEP_THREADS + threadIdx.x;
            if (e < n) {
                const int y = e / w, x = e - y * w;
                double f = (Z[y * pitch + x].x * inv_n - mean) * inv_sd;
                double nz = 0.0;
                if (nug > 0.0) {
                    double z0, z1;
                    if (INJECT) z0 = z_nug[e];
                    else box_muller(rng((uint32_t)e, it_lo, it_hi, GMC_STREAM_NUGGET), z0, z1);
                    nz = sq_nug * z0;
                }
                f = add_rn(mul_rn(f, scale), nz);                                     // MCMC.py:250
                if (apply_taper) f = mul_rn(f, __ldg(taper + e));                     // MCMC.py:778
                v[k] = f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            const int e = (wave * CH + k) * GMC_STEP_THREADS + threadIdx.x;
            if (e < n) buf[e] = v[k];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4: the Metropolis step given f                                                        MCMC.py:1263-1360
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_window(StepScalars& s, int H, int W) {
    // MCMC.py:1267-1276 (h, w even so h/2 is exact)
    const int h2 = s.h / 2, w2 = s.w / 2;
    s.x0 = max(0, s.ix - h2);
    s.x1 = min(H, s.ix + h2);
    s.y0 = max(0, s.iy - w2);
    s.y1 = min(W, s.iy + w2);
    s.mx0 = max(s.h - s.x1, 0);
    s.my0 = max(s.w - s.y1, 0);
}

__device__ __forceinline__ double sq_or_zero(double v) { return (v == v) ? mul_rn(v, v) : 0.0; }

// f: [h][w] with row pitch f_pitch (shared or global).  tile: (bh+2)x(bw+2) doubles; newres: bh*bw doubles (may alias f
// only if f is dead after the tile is built, which holds: f is read in phase A only).
__device__ void step_tail(const GmcDev& d, StepScalars* sc, double* scratch, const double* f, int f_pitch, double* tile,
                          double* newres, double* bed, double* mcres, double& ssq, int32_t* resampled,
                          double* loss_next_out) {
    const int H = d.H, W = d.W;
    const StepScalars s = *sc;
    const int bh = s.x1 - s.x0, bw = s.y1 - s.y0;
    const int tp = bw + 2;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);

    // phase A: candidate tile = block + one-cell halo                                  MCMC.py:1279-1290
    for (int t = threadIdx.x; t < (bh + 2) * tp; t += GMC_STEP_THREADS) {
        const int ti = t / tp, tj = t - ti * tp;
        const int i = s.x0 - 1 + ti, j = s.y0 - 1 + tj;
        double v = qnan;
        if (i >= 0 && i < H && j >= 0 && j < W) {
            const int64_t idx = (int64_t)i * W + j;
            v = __ldcg(bed + idx);
            if (ti >= 1 && ti <= bh && tj >= 1 && tj <= bw && (__ldg(d.flags + idx) & FLAG_GATE)) {
                double p = f[(s.mx0 + ti - 1) * f_pitch + (s.my0 + tj - 1)];
                if (d.crf_weight) p = mul_rn(p, __ldg(d.crf_weight + idx));
                v = add_rn(v, p);
            }
        }
        tile[t] = v;
    }
    __syncthreads();

    // phase B: residual on the block, loss delta, thickness guard                       MCMC.py:1292-1329
    double delta = 0.0;
    int bad = 0;
    for (int e = threadIdx.x; e < bh * bw; e += GMC_STEP_THREADS) {
        const int bi = e / bw, bj = e - bi * bw;
        const int i = s.x0 + bi, j = s.y0 + bj;
        const int64_t idx = (int64_t)i * W + j;
        const double* tc = tile + (bi + 1) * tp + (bj + 1);
        double dx, dy;
        {
            int jl = j - 1, jr = j + 1;
            double den = d.two_res;
            if (j == 0) { jl = 0; den = d.res; }
            else if (j == W - 1) { jr = W - 1; den = d.res; }
            const int64_t r = (int64_t)i * W;
            const double fr = mul_rn(__ldg(d.velx + r + jr), sub_rn(__ldg(d.surf + r + jr), tc[jr - j]));
            const double fl = mul_rn(__ldg(d.velx + r + jl), sub_rn(__ldg(d.surf + r + jl), tc[jl - j]));
            dx = div_rn(sub_rn(fr, fl), den);
        }
        {
            int iu = i - 1, id = i + 1;
            double den = d.two_res;
            if (i == 0) { iu = 0; den = d.res; }
            else if (i == H - 1) { id = H - 1; den = d.res; }
            const double fd = mul_rn(__ldg(d.vely + (int64_t)id * W + j), sub_rn(__ldg(d.surf + (int64_t)id * W + j), tc[(id - i) * tp]));
            const double fu = mul_rn(__ldg(d.vely + (int64_t)iu * W + j), sub_rn(__ldg(d.surf + (int64_t)iu * W + j), tc[(iu - i) * tp]));
            dy = div_rn(sub_rn(fd, fu), den);
        }
        const double rnew = sub_rn(add_rn(add_rn(dx, dy), __ldg(d.dhdt + idx)), __ldg(d.smb + idx));
        newres[e] = rnew;
        const uint8_t fl = __ldg(d.flags + idx);
        if (fl & FLAG_MC) {
            const double rold = __ldcg(mcres + idx);
            if (rnew == rnew && rold == rold) delta += (rnew - rold) * (rnew + rold);
            else delta += sq_or_zero(rnew) - sq_or_zero(rold);
        }
        if ((fl & FLAG_GATE) && sub_rn(__ldg(d.surf + idx), tc[0]) <= 0.0) bad = 1;
    }
    const double dsum = block_sum<GMC_STEP_THREADS>(delta, scratch);
    bad = __syncthreads_or(bad);

    // decision                                                                          MCMC.py:1331-1337
    if (threadIdx.x == 0) {
        const double ssq_next = ssq + dsum;
        const double loss_prev = div_rn(ssq, d.two_sigma2);
        double loss_next = div_rn(ssq_next, d.two_sigma2);
        if (bad) loss_next = __longlong_as_double(0x7ff0000000000000LL);
        double acc;
        if (loss_prev > loss_next) acc = 1.0;
        else {
            const double ex = exp(loss_prev - loss_next);
            acc = (ex < 1.0) ? ex : 1.0;      // python min(1, ex)
        }
        sc->accept = (s.u <= acc) ? 1 : 0;
        scratch[34] = ssq_next;
        scratch[35] = loss_next;
    }
    __syncthreads();
    const int accept = sc->accept;
    if (loss_next_out && threadIdx.x == 0) *loss_next_out = scratch[35];
    if (accept) {
        ssq = scratch[34];
        for (int e = threadIdx.x; e < bh * bw; e += GMC_STEP_THREADS) {
            const int bi = e / bw, bj = e - bi * bw;
            const int64_t idx = (int64_t)(s.x0 + bi) * W + (s.y0 + bj);
            __stcg(bed + idx, tile[(bi + 1) * tp + (bj + 1)]);
            __stcg(mcres + idx, newres[e]);
            if (resampled && (__ldg(d.flags + idx) & FLAG_GATE)) resampled[idx] += 1;
        }
    }
    __syncthreads();   // write-back visible to the next iteration's tile load; smem free for reuse
}

// full masked nansum of the tracked residual (fixed order)
__device__ double resync_ssq(const GmcDev& d, const double* mcres, double* scratch) {
    const int64_t n = (int64_t)d.H * d.W;
    double acc = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += GMC_STEP_THREADS) {
        const double v = __ldcg(mcres + k);
        if ((__ldg(d.flags + k) & FLAG_MC) && v == v) acc += v * v;
    }
    return block_sum<GMC_STEP_THREADS>(acc, scratch);
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char gmc_smem[];

__global__ void __launch_bounds__(GMC_STEP_THREADS, 2)
    run_kernel(GmcDev d, double* bed_all, double* mcres_all, double* ssq_all, const uint64_t* __restrict__ seeds,
               uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache,
               int64_t cache_stride, int64_t cache_offset, int32_t* resampled_all, int resync_every) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int c = blockIdx.x;
    const int64_t plane = (int64_t)d.H * d.W;
    double* bed = bed_all + c * plane;
    double* mcres = mcres_all + c * plane;
    int32_t* resampled = resampled_all ? resampled_all + c * plane : nullptr;
    const Philox rng(seeds[c]);
    double ssq = ssq_all[c];

    for (int k = 0; k < n_steps; ++k) {
        const uint64_t it = iter0 + (uint64_t)k;
        const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
        if (resync_every > 0 && it % (uint64_t)resync_every == 0) ssq = resync_ssq(d, mcres, scratch);
        if (threadIdx.x == 0) {
            const GmcFieldModel& fm = d.fm;
            // RandField stream: block size, scale, nugget, range(s)                   MCMC.py:755, 200-207
            const uint4 r0 = rng(0u, it_lo, it_hi, GMC_STREAM_RF_SCALARS);
            const uint4 r1 = rng(1u, it_lo, it_hi, GMC_STREAM_RF_SCALARS);
            sc.pair = (int)bounded_u64(r0.x, r0.y, (uint64_t)d.n_pairs);
            sc.scale = div_rn(add_rn(fm.scale_min, mul_rn(sub_rn(fm.scale_max, fm.scale_min), u01_halfopen(r0.z, r0.w))), 3.0);
            sc.nug = add_rn(0.0, mul_rn(fm.nugget_max, u01_halfopen(r1.x, r1.y)));
            sc.range_x = add_rn(fm.range_min_x, mul_rn(sub_rn(fm.range_max_x, fm.range_min_x), u01_halfopen(r1.z, r1.w)));
            if (fm.isotropic) sc.range_y = sc.range_x;
            else {
                const uint4 r2 = rng(2u, it_lo, it_hi, GMC_STREAM_RF_SCALARS);
                sc.range_y = add_rn(fm.range_min_y, mul_rn(sub_rn(fm.range_max_y, fm.range_min_y), u01_halfopen(r2.x, r2.y)));
            }
            // chain stream: block centre (uniform over the allowed cells) and the acceptance uniform   MCMC.py:1253-1261, 1336
            const uint4 c0 = rng(0u, it_lo, it_hi, GMC_STREAM_CHAIN);
            if (d.n_centre_cells > 0) {
                const int32_t cell = d.centre_cells[bounded_u64(c0.x, c0.y, (uint64_t)d.n_centre_cells)];
                sc.ix = cell / d.W;
                sc.iy = cell - sc.ix * d.W;
            } else {
                sc.ix = (int)bounded_u64(c0.x, c0.y, (uint64_t)d.H);
                sc.iy = (int)bounded_u64(c0.z, c0.w, (uint64_t)d.W);
            }
            const uint4 c1 = rng(1u, it_lo, it_hi, GMC_STREAM_CHAIN);
            sc.u = u01_halfopen(c1.x, c1.y);
            sc.h = d.pairs[sc.pair].h;
            sc.w = d.pairs[sc.pair].w;
            block_window(sc, d.H, d.W);
        }
        __syncthreads();
        synth_field<false>(d, buf, scratch, sc.pair, sc.scale, sc.nug, sc.range_x, sc.range_y, rng, it_lo, it_hi, nullptr,
                           nullptr, nullptr, true);
        const int n = sc.h * sc.w;
        step_tail(d, &sc, scratch, buf, sc.w, buf + n, buf, bed, mcres, ssq, resampled, nullptr);
        if (threadIdx.x == 0) {
            const int64_t slot = (int64_t)c * cache_stride + cache_offset + k;
            if (loss_cache) loss_cache[slot] = div_rn(ssq, d.two_sigma2);
            if (step_cache) step_cache[slot] = (uint8_t)sc.accept;
            if (blocks_cache) {
                int4 b = make_int4(sc.ix, sc.iy, sc.h, sc.w);
                reinterpret_cast<int4*>(blocks_cache)[slot] = b;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ssq_all[c] = ssq;
}

__global__ void __launch_bounds__(GMC_STEP_THREADS)
    step_injected_kernel(GmcDev d, double* bed_all, double* mcres_all, double* ssq_all, const double* __restrict__ f_all,
                         int64_t f_stride, const int32_t* __restrict__ hw, const int32_t* __restrict__ centre,
                         const double* __restrict__ u, uint8_t* accepted_out, double* loss_out, double* loss_next_out,
                         int32_t* resampled_all, int hmax, int wmax) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int c = blockIdx.x;
    const int64_t plane = (int64_t)d.H * d.W;
    if (threadIdx.x == 0) {
        sc.h = hw[2 * c];
        sc.w = hw[2 * c + 1];
        sc.ix = centre[2 * c];
        sc.iy = centre[2 * c + 1];
        sc.u = u[c];
        sc.pair = -1;
        block_window(sc, d.H, d.W);
    }
    __syncthreads();
    double ssq = ssq_all[c];
    step_tail(d, &sc, scratch, f_all + c * f_stride, sc.w, buf + (int64_t)hmax * wmax, buf, bed_all + c * plane,
              mcres_all + c * plane, ssq, resampled_all ? resampled_all + c * plane : nullptr,
              loss_next_out ? loss_next_out + c : nullptr);
    if (threadIdx.x == 0) {
        ssq_all[c] = ssq;
        if (accepted_out) accepted_out[c] = (uint8_t)sc.accept;
        if (loss_out) loss_out[c] = div_rn(ssq, d.two_sigma2);
    }
}

template <bool INJECT>
__global__ void __launch_bounds__(GMC_STEP_THREADS, 2)
    field_kernel(GmcDev d, const int32_t* __restrict__ pair, const double* __restrict__ scale, const double* __restrict__ nug,
                 const double* __restrict__ range_x, const double* __restrict__ range_y, const double* __restrict__ z_re,
                 const double* __restrict__ z_im, const double* __restrict__ z_nug, const uint64_t* __restrict__ seeds,
                 uint64_t iter, int apply_taper, double* __restrict__ f_out, int64_t stride) {
    __shared__ double scratch[40];
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int i = blockIdx.x;
    const int p = pair[i];
    const Philox rng(INJECT ? 0ull : seeds[i]);
    synth_field<INJECT>(d, buf, scratch, p, scale[i], nug[i], range_x[i], range_y[i], rng, (uint32_t)iter,
                        (uint32_t)(iter >> 32), INJECT ? z_re + i * stride : nullptr, INJECT ? z_im + i * stride : nullptr,
                        INJECT ? z_nug + i * stride : nullptr, apply_taper != 0);
    const int n = d.pairs[p].h * d.pairs[p].w;
    for (int e = threadIdx.x; e < n; e += GMC_STEP_THREADS) f_out[i * stride + e] = buf[e];
}

// ---------------------------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------------------------
static size_t tail_bytes(int h, int w) { return ((size_t)h * w + (size_t)(h + 2) * (w + 2)) * sizeof(double); }

int gmc_step_configure(gmc_ctx* c) {
    size_t need = 0;
    for (const GmcPair& p : c->h_pairs) {
        const size_t z = (size_t)p.h * (p.w + 1) * sizeof(double2);
        need = std::max(need, std::max(z, tail_bytes(p.h, p.w)));
    }
    need = (need + 15) & ~(size_t)15;
    cudaDeviceProp prop;
    GMC_CUDA(cudaGetDeviceProperties(&prop, c->device));
    if (need > prop.sharedMemPerBlockOptin)
        GMC_FAIL(GMC_EUNSUPPORTED, "gmc_set_blocks: a %dx%d block needs %zu B of shared memory; the device offers %zu B",
                 c->max_h, c->max_w, need, (size_t)prop.sharedMemPerBlockOptin);
    GMC_CUDA(cudaFuncSetAttribute(run_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    GMC_CUDA(cudaFuncSetAttribute(field_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    GMC_CUDA(cudaFuncSetAttribute(field_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
    c->step_smem_bytes = (int)need;
    int nb = 0;
    GMC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, run_kernel, GMC_STEP_THREADS, need));
    c->step_ctas_per_sm = nb;
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
    const size_t smem = (tail_bytes(hmax, wmax) + 15) & ~(size_t)15;
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
    run_kernel<<<C, GMC_STEP_THREADS, c->step_smem_bytes, (cudaStream_t)stream>>>(c->dev, bed, mcres, ssq, seeds, iter0, n_steps,
                                                                                 loss_cache, step_cache, blocks_cache,
                                                                                 cache_stride, cache_offset, resampled,
                                                                                 resync_every);
    c->launches++;
    GMC_CUDA(cudaGetLastError());
    return GMC_OK;
}
