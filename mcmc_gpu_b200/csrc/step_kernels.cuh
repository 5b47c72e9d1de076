// step_kernels.cuh - device code of K1 (proposal field synthesis) and K4 (Metropolis step); see step.cu for the overview.
// Included by step.cu once per CTA size inside a namespace (t256: 256 threads, 2 CTAs per SM - the default; t512: 512
// threads, one CTA per SM - used when a launch has no more chains than the GPU has SMs, so that a chain gets all the warps
// of its SM instead of half of them).  GMC_STEP_THREADS / GMC_STEP_MIN_CTAS select the size; no include guard on purpose.
struct SpecParams {
    int model;
    double a;        // sqrt(len_x*len_y)
    double nu;
    double constant; // Matern prefactor
    double kappa;    // 2 nu / a^2
    double amp0;     // sqrt(S) = amp0 * exp(qexp * L): L = (a k)^2 (Gaussian), log(1 + (a k)^2) (Exponential),
    double qexp;     //                                  log(kappa + 4 pi k^2) (Matern)
};

struct StepScalars {
    int pair, h, w, ix, iy;
    int x0, x1, y0, y1, mx0, my0;
    int tp, toff, tc0;     // candidate tile: row pitch (doubles), offset of column y0-1 inside a row, grid column of tile column 0
    double scale, nug, range_x, range_y, u;
    int accept;
    SpecParams spec;
};

// exact t / d for t * d < 2^32 with one multiply-high
struct FastDiv {
    uint32_t m, d;
    __device__ __forceinline__ explicit FastDiv(int dd) : m(dd == 1 ? 0u : 0xFFFFFFFFu / (uint32_t)dd + 1u), d((uint32_t)dd) {}
    __device__ __forceinline__ int div(int t) const { return d == 1 ? t : (int)__umulhi((uint32_t)t, m); }
};

// Optional per-phase cycle accounting (thread 0 of each CTA; enabled when the host passes a buffer): used by
// profiles/phase_timing.py to attribute the step time without relying on SASS line tables.
struct PhaseClock {
    long long* acc;      // [GMC_N_PHASES] global, atomically accumulated; nullptr = disabled
    long long t;
    __device__ __forceinline__ void start() { if (acc && threadIdx.x == 0) t = clock64(); }
    __device__ __forceinline__ void mark(int phase) {
        if (acc && threadIdx.x == 0) {
            const long long n = clock64();
            atomicAdd(reinterpret_cast<unsigned long long*>(acc + phase), (unsigned long long)(n - t));
            t = n;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------------
// Canonical reductions: the step kernels exist for two CTA sizes, and a chain's trajectory must not depend on which one
// advances it (bit-identical across batch compositions and GPU counts).  Every sum over a CTA is therefore defined over V
// virtual slots - item q belongs to slot q mod V, slots accumulate in increasing q, groups of 32 slots are summed by the
// warp shuffle tree, the V/32 group sums by one more shuffle tree - and a CTA with fewer than V working threads keeps
// V / workers accumulators per thread (thread t: slots t, t + workers, ...).
// ---------------------------------------------------------------------------------------------------------------
#define GMC_SUM_SLOTS 512                                  // fill power, residual re-sum: all threads work
#define GMC_SUM_NACC (GMC_SUM_SLOTS / GMC_STEP_THREADS)    // 2 (256 threads) or 1 (512 threads)
#define GMC_TAIL_SLOTS 448                                 // residual phase: 14 worker warps' worth
#define GMC_TAIL_WORKERS (GMC_STEP_THREADS == 256 ? 224 : 448)
#define GMC_TAIL_NACC (GMC_TAIL_SLOTS / GMC_TAIL_WORKERS)
static_assert(GMC_SUM_NACC * GMC_STEP_THREADS == GMC_SUM_SLOTS && GMC_TAIL_NACC * GMC_TAIL_WORKERS == GMC_TAIL_SLOTS, "CTA size");

// acc[a] = partial sum of slot a * WORKERS + threadIdx.x (threads >= WORKERS pass zeros).  Result valid in all threads.
template <int NACC, int WORKERS>
__device__ __forceinline__ double canon_block_sum(const double (&acc)[NACC], double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int GROUPS = NACC * WORKERS / 32;
    static_assert(GROUPS <= 32, "one final warp");
    __syncthreads();                       // protect scratch from a previous use
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        const double v = warp_sum(acc[a]);
        if (lane == 0 && wid < WORKERS / 32) scratch[a * (WORKERS / 32) + wid] = v;
    }
    __syncthreads();
    if (wid == 0) {
        double t = (lane < GROUPS) ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

// ---------------------------------------------------------------------------------------------------------------
// complex helpers
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmuli(double2 a) { return make_double2(-a.y, a.x); }   // a * (+i)
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

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
    for (int m = 1; m <= HALF; ++m) w[m] = tw[m * step_r];
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

// One radix-R decimation-in-time stage over `count` independent lines of length n, in place.  Element p of line l is
// Z[l * line_stride + p * elem_stride]: the column pass has (1, pitch) — lanes on adjacent columns — and the row pass
// (pitch, 1) with lanes on different rows (the pitch is odd in double2 units, so both are bank-conflict free).
template <int R>
__device__ __forceinline__ void fft_stage(double2* Z, int line_stride, int elem_stride, int n, int count, int L,
                                          const double2* __restrict__ tw) {
    const int M = L / R;
    const int step = n / L;
    const int items = (n / R) * count;
    const FastDiv dcount(count), dM(M);
    const int stride = M * elem_stride;
    for (int t = threadIdx.x; t < items; t += GMC_STEP_THREADS) {
        const int bf = dcount.div(t);
        const int line = t - bf * count;
        const int blk = dM.div(bf), k1 = bf - blk * M;
        const int base = blk * L + k1;
        double2* p0 = Z + line * line_stride + base * elem_stride;
        double2 v[R];
#pragma unroll
        for (int q = 0; q < R; ++q) v[q] = p0[q * stride];
        if (L > R) {   // the first stage has all twiddles == 1
            const int i1 = k1 * step;
            int iq = i1;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                v[q] = cmul(v[q], tw[iq]);
                iq += i1;
                if (iq >= n) iq -= n;
            }
        }
        if (R == 2) dft2(v);
        else if (R == 4) dft4(v);
        else if (R == 8) dft8(v);
        else dft_odd<R>(v, tw, n / R);
#pragma unroll
        for (int q = 0; q < R; ++q) p0[q * stride] = v[q];
    }
}

// Fallback for any other prime radix R < GMC_MAX_RADIX (block edges such as 58 = 2*29): O(R^2) per butterfly with the
// inputs parked in local memory.  Slow but rare; the default block sizes never take it.
__device__ __noinline__ void fft_stage_generic(double2* Z, int line_stride, int elem_stride, int n, int count, int L, int R,
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
        double2* p0 = Z + line * line_stride + (blk * L + k1) * elem_stride;
        const int i1 = k1 * step;
        int iq = 0;
        for (int q = 0; q < R; ++q) {
            const double2 x = p0[q * M * elem_stride];
            v[q] = (q == 0 || L == R) ? x : cmul(x, tw[iq]);
            iq += i1;
            if (iq >= n) iq -= n;
        }
        for (int k = 0; k < R; ++k) {
            double2 acc = v[0];
            int m = 0;
            for (int q = 1; q < R; ++q) {
                m += k;
                if (m >= R) m -= R;
                acc = cadd(acc, cmul(v[q], tw[m * step_r]));
            }
            p0[k * M * elem_stride] = acc;
        }
    }
}

// All stages of one pass.  Kept out of line: the column and the row pass share one copy of the stage code (the step
// kernel is instruction-cache sensitive).
__device__ __noinline__ void fft_lines(double2* Z, int line_stride, int elem_stride, const GmcFftPlan& plan, int count,
                                       const double2* __restrict__ tw) {
    int L = 1;
    for (int s = 0; s < plan.n_factors; ++s) {
        const int r = plan.radix[s];
        L *= r;
        switch (r) {
            case 2: fft_stage<2>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            case 3: fft_stage<3>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            case 4: fft_stage<4>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            case 5: fft_stage<5>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            case 7: fft_stage<7>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            case 8: fft_stage<8>(Z, line_stride, elem_stride, plan.n, count, L, tw); break;
            default: fft_stage_generic(Z, line_stride, elem_stride, plan.n, count, L, r, tw); break;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Lean FP64 elementary functions for the spectrum fill.  The fill is 38 percent of the step kernel's instructions and it is
// issue bound; libdevice's log / exp / sincospi / sqrt spend half of theirs on argument classes that cannot occur here
// (zero, negative, subnormal, inf, nan, huge).  These versions assume positive, normal, moderately sized arguments and keep
// full double accuracy (<= 3.2e-16 relative against 120-bit references on millions of points, the prototype is kept in
// profiles/lean_math_check.py): fdlibm's log kernel, a degree-13 Taylor exp on |r| <= ln2/2, Taylor sin/cos of pi r on
// |r| <= 1/4 after an exact reduction, and Newton square roots / reciprocals seeded by the single-precision units.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double lean_rcp(double d) {                 // 1/d, d in [1.7, 3.5]
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)d));
    double r = (double)rf;
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return r;
}
__device__ __forceinline__ double lean_sqrt(double x) {                // x > 0, normal, within float range
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)x));
    double y = (double)yf;
    const double h = 0.5 * x;
    y = y * fma(-(h * y), y, 1.5);
    y = y * fma(-(h * y), y, 1.5);
    const double g = x * y;
    return fma(fma(-g, g, x), 0.5 * y, g);
}
__device__ __forceinline__ double lean_log(double x) {                 // x > 0, normal
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    if (hi >= 0x3ff6a09f) {                                            // mantissa above sqrt(2): halve it
        hi -= 0x00100000;
        e += 1;
    }
    const double f = __hiloint2double(hi, lo) - 1.0;                   // in [-0.2929, 0.4143)
    const double s = f * lean_rcp(2.0 + f);
    const double z = s * s;
    double R = fma(z, 1.479819860511658591e-01, 1.531383769920937332e-01);
    R = fma(z, R, 1.818357216161805012e-01);
    R = fma(z, R, 2.222219843214978396e-01);
    R = fma(z, R, 2.857142874366239149e-01);
    R = fma(z, R, 3.999999999940941908e-01);
    R = fma(z, R, 6.666666666666735130e-01);
    R *= z;
    const double lm = fma(s, R, s + s);                                // log(1 + f) = 2 s + s R
    const double de = (double)e;
    return fma(de, 6.93147180369123816490e-01, fma(de, 1.90821492927058770002e-10, lm));
}
__device__ __forceinline__ double lean_exp(double y) {                 // |y| < 700
    const double magic = 6755399441055744.0;                           // 1.5 * 2^52: rounds to the nearest integer
    const double t = fma(y, 1.4426950408889634074, magic);
    const int k = __double2loint(t);
    const double kd = t - magic;
    double r = fma(-kd, 6.93147180369123816490e-01, y);
    r = fma(-kd, 1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 0.0001984126984126984);
    p = fma(p, r, 0.001388888888888889);
    p = fma(p, r, 0.008333333333333333);
    p = fma(p, r, 0.041666666666666664);
    p = fma(p, r, 0.16666666666666666);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// sin and cos of 2 pi u, u in (0, 1)
__device__ __forceinline__ void lean_sincos2pi(double u, double& sn, double& cs) {
    const double magic = 6755399441055744.0;
    const double t4 = 4.0 * u;                                         // quarter turns
    const double tm = t4 + magic;
    const int j = __double2loint(tm);                                  // nearest quarter turn, 0..4
    const double r = 0.5 * (t4 - (tm - magic));                        // exact; sin(pi r), cos(pi r) with |r| <= 1/4
    const double z = r * r;
    double sp = 7.952054001475508e-07;
    sp = fma(sp, z, -2.1915353447830204e-05);
    sp = fma(sp, z, 0.00046630280576761234);
    sp = fma(sp, z, -0.007370430945714348);
    sp = fma(sp, z, 0.08214588661112819);
    sp = fma(sp, z, -0.5992645293207919);
    sp = fma(sp, z, 2.550164039877345);
    sp = fma(sp, z, -5.167712780049969);
    sp = fma(sp, z, 3.141592653589793);
    sp *= r;
    double cp = -1.387895246221376e-07;
    cp = fma(cp, z, 4.303069587032944e-06);
    cp = fma(cp, z, -0.00010463810492484565);
    cp = fma(cp, z, 0.001929574309403922);
    cp = fma(cp, z, -0.02580689139001405);
    cp = fma(cp, z, 0.23533063035889312);
    cp = fma(cp, z, -1.3352627688545893);
    cp = fma(cp, z, 4.058712126416768);
    cp = fma(cp, z, -4.934802200544679);
    cp = fma(cp, z, 1.0);
    const double a = (j & 1) ? cp : sp, b = (j & 1) ? sp : cp;
    sn = __hiloint2double(__double2hiint(a) ^ ((j & 2) << 30), __double2loint(a));
    cs = __hiloint2double(__double2hiint(b) ^ (((j + 1) & 2) << 30), __double2loint(b));
}

// ---------------------------------------------------------------------------------------------------------------
// spectral amplitude sqrt(S(k))                                                          MCMC.py:209-239
// ---------------------------------------------------------------------------------------------------------------
// x^p for x > 0 as exp(p log x): relative error ~ |p log x| 2^-53 (<= 1e-14 here), about half the cost of pow()
__device__ __forceinline__ double pow_pos(double x, double p) { return exp(p * log(x)); }

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
        sp.constant = div_rn(fm.matern_num, mul_rn(fm.matern_gamma, pow_pos(sp.a, mul_rn(2.0, sp.nu))));
        sp.kappa = div_rn(mul_rn(2.0, sp.nu), mul_rn(sp.a, sp.a));
    }
    sp.amp0 = (fm.model == GMC_MATERN) ? sqrt(sp.constant) : 1.0;
    sp.qexp = (fm.model == GMC_GAUSSIAN) ? -0.25 : (fm.model == GMC_EXPONENTIAL ? -0.75 : -0.5 * (sp.nu + 1.0));
    return sp;
}

// sqrt(S(k)) for k^2 = ksq_sum (MCMC.py:224-239: k = sqrt(kx^2 + ky^2) + 1e-10, S by model, then sqrt): the square root of
// the density is folded into the exponent, amp0 * exp(qexp * L) - equal to the reference's sqrt(S) to ~1e-15 relative
// (the parity bar for fields is 1e-9 of max|f|).
__device__ __forceinline__ double spec_sqrt_density(const SpecParams& sp, double ksq_sum) {
    const double k = add_rn((ksq_sum > 0.0) ? lean_sqrt(ksq_sum) : 0.0, 1e-10);
    double L;
    if (sp.model == GMC_GAUSSIAN) {
        const double ak = mul_rn(sp.a, k);
        L = mul_rn(ak, ak);
    } else if (sp.model == GMC_EXPONENTIAL) {
        const double ak = mul_rn(sp.a, k);
        L = lean_log(add_rn(1.0, mul_rn(ak, ak)));
    } else {
        const double four_pi = 4 * 3.141592653589793;
        L = lean_log(add_rn(sp.kappa, mul_rn(four_pi, mul_rn(k, k))));
    }
    return sp.amp0 * lean_exp(sp.qexp * L);
}
__device__ __noinline__ double spec_amp(const SpecParams& sp, double ksq_sum) { return spec_sqrt_density(sp, ksq_sum); }

// One spectrum item of the device-RNG path: sqrt(S) and the two complex normals of its mirrored rows.  A single
// out-of-line body (one copy in the instruction cache) in which the three dependency chains - log/exp of the density,
// log/sqrt/sincos of the two Box-Muller draws - are independent, so the scheduler interleaves them.
struct FillItem {
    double amp, z0, z1, y0, y1;
};
__device__ __noinline__ void fill_item(const SpecParams& sp, const Philox& rng, uint32_t it_lo, uint32_t it_hi, double ks,
                                       uint32_t e0, uint32_t e1, FillItem& r) {
    const uint4 a = rng(e0, it_lo, it_hi, GMC_STREAM_NOISE), b = rng(e1, it_lo, it_hi, GMC_STREAM_NOISE);
    const double l0 = lean_log(u01_open(a.x, a.y)), l1 = lean_log(u01_open(b.x, b.y));
    double s0, c0, s1, c1;
    lean_sincos2pi(u01_open(a.z, a.w), s0, c0);
    lean_sincos2pi(u01_open(b.z, b.w), s1, c1);
    r.amp = spec_sqrt_density(sp, ks);
    const double q0 = lean_sqrt(-2.0 * l0), q1 = lean_sqrt(-2.0 * l1);
    r.z0 = q0 * c0;
    r.z1 = q0 * s0;
    r.y0 = q1 * c1;
    r.y1 = q1 * s1;
}

// ---------------------------------------------------------------------------------------------------------------
// K1: synthesise one field into shared memory.  On return (after its final __syncthreads) the standardised field times
// `scale` sits in buf as F[y * fpitch + x] with fpitch = 2 * pitchc doubles (rows are the complex rows of the half plane).
// Nugget noise and taper are applied by the consumer (field_value()).
// ---------------------------------------------------------------------------------------------------------------
struct FieldView {
    const double* F;        // shared memory
    int fpitch;
    int w;
    double sq_nug;          // sqrt(nug); 0 => no nugget term
    const double* taper;    // global [h][w] or nullptr
    const double* z_nug;    // injected unit normals [h][w] or nullptr (=> Philox)
};

template <bool INJECT>
__device__ __forceinline__ double field_value(const FieldView& fv, int y, int x, const Philox& rng, uint32_t it_lo,
                                              uint32_t it_hi) {
    double f = fv.F[y * fv.fpitch + x];
    if (fv.sq_nug > 0.0) {                                                            // MCMC.py:250
        const int e = y * fv.w + x;
        double z0, z1;
        if (INJECT) z0 = fv.z_nug[e];
        else box_muller(rng((uint32_t)e, it_lo, it_hi, GMC_STREAM_NUGGET), z0, z1);
        f = add_rn(f, mul_rn(fv.sq_nug, z0));
    }
    return f;
}

// Per-step copies of the current block size's small tables: every later access is a shared-memory read instead of an
// L2 round trip on the critical path (the SM's L1 is almost entirely carved out as shared memory).
struct StepTables {
    double2 tw_h[GMC_MAX_EDGE];            // e^{+2 pi i k/h}
    double2 tw_w[GMC_MAX_EDGE / 2];        // e^{+2 pi i k/(w/2)}
    double2 htw[GMC_MAX_EDGE / 2];         // e^{+2 pi i k/w}
    double ksq_y[GMC_MAX_EDGE / 2 + 1], ksq_x[GMC_MAX_EDGE / 2 + 1];
    int16_t pos_y[GMC_MAX_EDGE], pos_w[GMC_MAX_EDGE / 2];
};

__device__ __forceinline__ void stage_tables(const GmcDev& d, const GmcPair& pr, StepTables& T, int first = threadIdx.x,
                                             int stride = GMC_STEP_THREADS) {
    const int h = pr.h, n2 = pr.w / 2;
    for (int t = first; t < h; t += stride) {
        T.tw_h[t] = __ldg(d.twiddle + pr.ph.tw_off + t);
        T.pos_y[t] = __ldg(d.pos + pr.ph.pos_off + t);
        if (t <= h / 2) T.ksq_y[t] = __ldg(d.ksq + pr.ksq_off_h + t);
        if (t < n2) {
            T.tw_w[t] = __ldg(d.twiddle + pr.pw.tw_off + t);
            T.htw[t] = __ldg(d.twiddle + pr.pw.htw_off + t);
            T.pos_w[t] = __ldg(d.pos + pr.pw.pos_off + t);
        }
        if (t <= n2) T.ksq_x[t] = __ldg(d.ksq + pr.ksq_off_w + t);
    }
    for (int t = h + first; t <= n2; t += stride) {                       // w/2 >= h (wide blocks)
        if (t < n2) {
            T.tw_w[t] = __ldg(d.twiddle + pr.pw.tw_off + t);
            T.htw[t] = __ldg(d.twiddle + pr.pw.htw_off + t);
            T.pos_w[t] = __ldg(d.pos + pr.pw.pos_off + t);
        }
        T.ksq_x[t] = __ldg(d.ksq + pr.ksq_off_w + t);
    }
}

// pr and T live in shared memory; the caller has staged T (stage_tables + __syncthreads).
template <bool INJECT>
__device__ FieldView synth_field(const GmcDev& d, double* buf, double* scratch, const GmcPair& pr, const StepTables& T,
                                 double scale, double nug, const SpecParams& sp, const Philox& rng, uint32_t it_lo, uint32_t it_hi,
                                 const double* __restrict__ z_re, const double* __restrict__ z_im,
                                 const double* __restrict__ z_nug, bool apply_taper, PhaseClock& pc) {
    const int h = pr.h, w = pr.w, n2 = w / 2, hc = n2 + 1, pitchc = pr.pitchc;
    const GmcFftPlan& ph = pr.ph;
    const GmcFftPlan& pw = pr.pw;
    double2* Z = reinterpret_cast<double2*>(buf);
    const int16_t* posY = T.pos_y;
    const double* ksqY = T.ksq_y;
    const double* ksqX = T.ksq_x;
    const double inv_sqrt2 = 0.70710678118654752440;

    // (1) fill the Hermitian half plane (ky in [0,h), kx in [0,w/2]) at the digit-reversed row position of the column
    // pass; one item per (|ky| = a, kx): the entries ky = a and ky = h-a share sqrt(S).  Accumulate the power for the
    // variance (Parseval): interior columns count twice (their mirror images kx > w/2 are not stored).
    double power[GMC_SUM_NACC];                                    // canonical slots, see canon_block_sum
#pragma unroll
    for (int a = 0; a < GMC_SUM_NACC; ++a) power[a] = 0.0;
    int trip = 0;                                                  // this thread's trip count: selects the accumulator
    const FastDiv dhc(hc);
    const int n_items = (h / 2 + 1) * hc;
    // store one item's (up to) two entries and accumulate their power
    auto emit = [&](int a, int kx, double2 X0, double2 X1) {
        const bool self_y = (a == 0 || 2 * a == h);
        const bool edge_x = (kx == 0 || kx == n2);
        if (a == 0 && kx == 0) X0 = X1 = make_double2(0.0, 0.0);   // DC: removed by the mean subtraction (MCMC.py:248)
        const double wgt = edge_x ? 1.0 : 2.0;
        double pw = wgt * (X0.x * X0.x + X0.y * X0.y);
        Z[posY[a] * pitchc + kx] = X0;
        if (!self_y) {
            pw += wgt * (X1.x * X1.x + X1.y * X1.y);
            Z[posY[h - a] * pitchc + kx] = X1;
        }
        if (GMC_SUM_NACC == 1 || (trip & 1) == 0) power[0] += pw;
        else power[GMC_SUM_NACC - 1] += pw;
        ++trip;
    };
    if (INJECT) {
        for (int q = threadIdx.x; q < n_items; q += GMC_STEP_THREADS) {
            const int a = dhc.div(q), kx = q - a * hc;
            // MCMC.py:224: k = sqrt(kxv**2 + kyv**2) + 1e-10
            const double amp = spec_amp(sp, add_rn(ksqX[kx], ksqY[a]));
            const int a2 = (a == 0 || 2 * a == h) ? a : h - a;
            // X_h(k) = sqrt(S)/2 * ((A_k + A_-k) + i (B_k - B_-k)),  -k = ((h-ky)%h, (w-kx)%w)
            const int kxm = (kx == 0) ? 0 : w - kx;
            const int e0 = a * w + kx, m0 = ((a == 0) ? 0 : h - a) * w + kxm;
            const int e1 = a2 * w + kx, m1 = ((a2 == 0) ? 0 : h - a2) * w + kxm;
            emit(a, kx, make_double2(0.5 * amp * (z_re[e0] + z_re[m0]), 0.5 * amp * (z_im[e0] - z_im[m0])),
                 make_double2(0.5 * amp * (z_re[e1] + z_re[m1]), 0.5 * amp * (z_im[e1] - z_im[m1])));
        }
    } else {
        // the second draw of an item is unused for self-conjugate rows / Hermitian columns (< 10 % of the items)
        for (int q = threadIdx.x; q < n_items; q += GMC_STEP_THREADS) {
            const int a = dhc.div(q), kx = q - a * hc;
            const bool self_y = (a == 0 || 2 * a == h), edge_x = (kx == 0 || kx == n2);
            const int a2 = self_y ? a : h - a;
            FillItem fi;
            fill_item(sp, rng, it_lo, it_hi, add_rn(ksqX[kx], ksqY[a]), (uint32_t)(a * w + kx), (uint32_t)(a2 * w + kx), fi);
            const double am = fi.amp * inv_sqrt2;
            const bool real_pt = self_y && edge_x;                 // self-conjugate point: real, variance S
            const double2 X0 = make_double2(real_pt ? fi.amp * fi.z0 : am * fi.z0, real_pt ? 0.0 : am * fi.z1);
            const double2 X1 = edge_x ? cconj(X0) : make_double2(am * fi.y0, am * fi.y1);   // kx = 0, w/2: Hermitian in ky
            emit(a, kx, X0, X1);
        }
    }
    const double power_sum = canon_block_sum<GMC_SUM_NACC, GMC_STEP_THREADS>(power, scratch);   // also orders the fill before the column pass
    // field = (1/(hw)) sum_k X_h e^{...};  var = power/(hw)^2;  (x - mean)/(std + 1e-12) * scale   MCMC.py:247-250
    const double inv_n = 1.0 / ((double)h * (double)w);
    const double sd = sqrt(power_sum) * inv_n;
    const double cscale = scale / (sd + 1e-12) * inv_n;
    pc.mark(1);

    // (2) inverse DFT along y for the w/2+1 stored columns
    fft_lines(Z, 1, pitchc, ph, hc, T.tw_h);
    pc.mark(2);

    // (3) real-row recombination: Y_k = (X_k + conj X_{n2-k}) + i (X_k - conj X_{n2-k}) e^{+2 pi i k/w}, k < n2, stored
    // at the digit-reversed position of the row pass.  One warp per row; a lane owns the pair (k, n2-k), reads both, and
    // only then writes, so the in-place permutation is safe (npairs <= 64).
    {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const int16_t* posW = T.pos_w;
        const double2* htw = T.htw;
        const int npairs = n2 / 2 + 1;
        for (int y = wid; y < h; y += GMC_STEP_THREADS / 32) {
            double2* row = Z + y * pitchc;
            double2 ya[2], yb[2];
            int ka[2], kb[2];
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                const int k = lane + 32 * it;
                ka[it] = kb[it] = -1;
                if (k < npairs) {
                    const int kp = n2 - k;                         // partner (k = 0 pairs with the Nyquist entry n2)
                    const double2 xk = row[k], xp = row[kp];
                    const double2 E = cadd(xk, cconj(xp));
                    const double2 O = cmul(csub(xk, cconj(xp)), (k == 0) ? make_double2(1.0, 0.0) : htw[k]);
                    ya[it] = make_double2((E.x - O.y) * cscale, (E.y + O.x) * cscale);          // E + iO
                    ka[it] = k;
                    if (k != 0 && kp != k) {
                        yb[it] = make_double2((E.x + O.y) * cscale, (-E.y + O.x) * cscale);     // conj(E) + i conj(O)
                        kb[it] = kp;
                    }
                }
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 2; ++it) {
                if (ka[it] >= 0) row[posW[ka[it]]] = ya[it];
                if (kb[it] >= 0) row[posW[kb[it]]] = yb[it];
            }
        }
        __syncthreads();
    }
    pc.mark(3);

    // (4) inverse DFT of length w/2 along x: row y now holds (f[y][2m], f[y][2m+1]) as its m-th complex entry
    fft_lines(Z, pitchc, 1, pw, h, T.tw_w);
    pc.mark(4);

    FieldView fv;
    fv.F = buf;
    fv.fpitch = 2 * pitchc;
    fv.w = w;
    fv.sq_nug = (nug > 0.0) ? sqrt(nug) : 0.0;
    fv.taper = apply_taper ? d.edge_masks + pr.mask_off : nullptr;
    fv.z_nug = z_nug;
    return fv;
}

// ---------------------------------------------------------------------------------------------------------------
// K4: the Metropolis step given f                                                        MCMC.py:1263-1360
// ---------------------------------------------------------------------------------------------------------------
// vec: the bed rows can be staged with 16-byte copies (even W, 16-byte aligned base): the tile then starts at an even grid
// column (one column left of y0-1 if needed) and has an even pitch
__device__ __forceinline__ void block_window(StepScalars& s, int H, int W, bool vec) {
    // MCMC.py:1267-1276 (h, w even so h/2 is exact)
    const int h2 = s.h / 2, w2 = s.w / 2;
    s.x0 = max(0, s.ix - h2);
    s.x1 = min(H, s.ix + h2);
    s.y0 = max(0, s.iy - w2);
    s.y1 = min(W, s.iy + w2);
    s.mx0 = max(s.h - s.x1, 0);
    s.my0 = max(s.w - s.y1, 0);
    const int bw = s.y1 - s.y0;
    if (vec) {
        s.tc0 = (s.y0 - 1) & ~1;                              // floor to even, also for -1 -> -2
        s.toff = (s.y0 - 1) - s.tc0;
        s.tp = (s.toff + bw + 3) & ~1;
    } else {
        s.tc0 = s.y0 - 1;
        s.toff = 0;
        s.tp = bw + 2;
    }
}

__device__ __forceinline__ double sq_or_zero(double v) { return (v == v) ? mul_rn(v, v) : 0.0; }

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
// ---- bulk asynchronous copies (TMA, 1-D) completing on an mbarrier: one instruction moves a whole tile row --------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, unsigned parity) {
    unsigned done;
    asm volatile(
        "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gmem, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}
// orders earlier generic-proxy accesses of shared memory before later asynchronous-proxy (TMA) writes to it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Start the tail's HBM traffic at the top of the step, while the field is synthesised: the bed of the block plus its
// one-cell halo lands asynchronously in the tile region, which is idle until the candidate is built - by TMA bulk copies
// completing on an mbarrier when `bar` is given (run_kernel, even W), else by cp.async - and the lines of the old block
// residual are pulled into L2.  Cells outside the grid become NaN (never used by the edge rules).
__device__ __forceinline__ void stage_block_async(const StepScalars& s, int H, int W, const double* bed, const double* mcres,
                                                  double* tile, bool vec, uint64_t* bar = nullptr) {
    const int bh = s.x1 - s.x0, bw = s.y1 - s.y0, tp = s.tp;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    if (bar && vec) {
        // One bulk copy per tile row, issued by one thread each (~80 instructions instead of ~250 cp.async warp
        // instructions; staging cost at the issuing threads 6.3 k -> 5.2 k cycles), one bulk L2 prefetch per residual row.
        // Columns [j_lo, j_hi) of the tile lie inside the grid: an even range, so every copy is 16-byte aligned and sized.
        const int j_lo = max(s.tc0, 0), j_hi = min(s.tc0 + tp, W);
        const int r_lo = max(s.x0 - 1, 0), r_hi = min(s.x1 + 1, H);
        const unsigned bytes = (unsigned)(j_hi - j_lo) * 8u;
        // rows are issued by the LAST threads of the CTA: the spectrum fill that follows hands the first threads one more
        // trip than the last ones, so the few thousand cycles these instructions take disappear in that slack
        const int rt = GMC_STEP_THREADS - 1 - (int)threadIdx.x;
        if (rt == 0) mbar_arrive_expect_tx(bar, (unsigned)(r_hi - r_lo) * bytes);
        for (int ti = rt; ti < bh + 2; ti += GMC_STEP_THREADS) {
            const int i = s.x0 - 1 + ti;
            double* dst = tile + ti * tp;
            if (i >= r_lo && i < r_hi) {
                fence_proxy_async();
                bulk_g2s(dst + (j_lo - s.tc0), bed + (int64_t)i * W + j_lo, bytes, bar);
                if (ti >= 1 && ti <= bh) bulk_prefetch_l2(mcres + (int64_t)i * W + j_lo, bytes);
                for (int tj = 0; tj < j_lo - s.tc0; ++tj) dst[tj] = qnan;
                for (int tj = j_hi - s.tc0; tj < tp; ++tj) dst[tj] = qnan;
            } else {
                for (int tj = 0; tj < tp; ++tj) dst[tj] = qnan;
            }
        }
        return;
    }
    for (int ti = wid; ti < bh + 2; ti += GMC_STEP_THREADS / 32) {          // one warp per tile row: coalesced
        const int i = s.x0 - 1 + ti;
        const bool row_in = i >= 0 && i < H;
        const double* src = bed + (int64_t)i * W + s.tc0;
        double* dst = tile + ti * tp;
        if (vec) {
            // column pairs (tc0 + 2m, tc0 + 2m + 1): with W and tc0 even a pair is entirely inside or outside the grid;
            // 16-byte copies halve the number of asynchronous-copy instructions, which is what this phase costs
            for (int m = lane; 2 * m < tp; m += 32) {
                const int j = s.tc0 + 2 * m;
                if (row_in && j >= 0 && j < W) cp_async16(dst + 2 * m, src + 2 * m);
                else dst[2 * m] = dst[2 * m + 1] = qnan;
            }
        } else {
            for (int tj = lane; tj < tp; tj += 32) {
                const int j = s.tc0 + tj;
                if (row_in && j >= 0 && j < W) cp_async8(dst + tj, src + tj);
                else dst[tj] = qnan;
            }
        }
        if (ti >= 1 && ti <= bh) {                                          // old residual of this block row -> L2
            const char* base = reinterpret_cast<const char*>(mcres + (int64_t)i * W + s.y0);
            for (int off = lane * 128; off < bw * 8 + 127; off += 32 * 128)
                prefetch_l2(base + min(off, bw * 8 - 8));
        }
    }
    cp_async_commit();
}

// What the helper warp of run_kernel prepares during the residual phase: the next step's scalars, block record, tables.
struct StepTables;
struct NextStep {
    StepScalars* sc;
    GmcPair* pair;
    StepTables* tab;
    uint64_t it;
    bool vec;
};
__device__ void prepare_step(const GmcDev& d, const Philox& rng, uint64_t it, StepScalars& sc, GmcPair& s_pair, StepTables& s_tab,
                             bool vec);

// Field source for the tail: either the synthesised field in shared memory (FieldView) or an injected f in global memory.
// HELPER: the last warp does not take residual cells; it prepares the next step instead (run_kernel).
template <bool INJECT_F, bool HELPER = false>
__device__ void step_tail(const GmcDev& d, StepScalars* sc, double* scratch, const FieldView& fv, const double* f_inj,
                          int f_pitch, const Philox& rng, uint32_t it_lo, uint32_t it_hi, double* tile, double* newres,
                          double* bed, double* mcres, double& ssq, int32_t* resampled, double* loss_next_out,
                          PhaseClock& pc, const NextStep* next = nullptr, uint64_t* tile_bar = nullptr, unsigned tile_parity = 0,
                          int* err = nullptr, int* consumed_flag = nullptr, int consumed_val = 0) {
    const int H = d.H, W = d.W;
    const StepScalars s = *sc;
    const int bh = s.x1 - s.x0, bw = s.y1 - s.y0;
    const int tp = s.tp;
    tile += s.toff;                                           // tile[(bi+1)*tp + (bj+1)] is the cell (x0+bi, y0+bj)
    const FastDiv dbw(bw);

    // phase A: candidate = bed + perturbation on the gated block cells (the tile already holds the bed, staged
    // asynchronously at the top of the step); loads are issued in batches so one L2 round trip serves U cells.
    if (tile_bar) {
        unsigned spins = 0;
        while (!mbar_try_wait(tile_bar, tile_parity))
            if (++spins > (1u << 22)) {                       // byte counts match by construction; never hang, but never go on
                if (err) *(volatile int*)err = GMC_DEVERR_WAIT_TIMEOUT;   // silently either: the host reports GMC_ECUDA
                break;
            }
    } else cp_async_wait_all();
    __syncthreads();
    {
        // One warp per block row, lanes across the columns (chunks of 32): the row's base pointers are formed once and every
        // access is base + lane (+ 32, + 64, ...), instead of a division and three 64-bit address computations per cell -
        // this phase is issue bound (ncu: 79 thread-instructions per cell before, 13.5 % of the kernel's instructions).
        // Two rows per trip keep up to 3 loads x 3 chunks x 2 rows in flight per lane.
        constexpr int NWARP = GMC_STEP_THREADS / 32;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        const bool use_taper = !INJECT_F && fv.taper != nullptr;
        const bool use_cw = d.crf_weight != nullptr;
        for (int bi0 = wid; bi0 < bh; bi0 += 2 * NWARP) {
            uint8_t fl[2][3];
            double tpv[2][3], cw[2][3];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int bi = bi0 + r * NWARP;
                const int64_t row = (int64_t)(s.x0 + bi) * W + s.y0;
                const double* tprow = use_taper ? fv.taper + (s.mx0 + bi) * fv.w + s.my0 : nullptr;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int bj = lane + 32 * c;
                    const bool on = bi < bh && bj < bw;
                    fl[r][c] = on ? __ldg(d.flags + row + bj) : 0;
                    cw[r][c] = (on && use_cw) ? __ldg(d.crf_weight + row + bj) : 1.0;
                    tpv[r][c] = (on && use_taper) ? __ldg(tprow + bj) : 1.0;
                }
            }
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int bi = bi0 + r * NWARP;
                if (bi >= bh) break;                                                   // warp-uniform
                const int fy = s.mx0 + bi;
                double* trow = tile + (bi + 1) * tp + 1;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const int bj = lane + 32 * c;
                    if (bj < bw && (fl[r][c] & FLAG_GATE)) {
                        double pv;
                        if (INJECT_F) pv = __ldcg(f_inj + fy * f_pitch + s.my0 + bj);      // L2: a producer CTA may have written it
                        else {
                            pv = field_value<false>(fv, fy, s.my0 + bj, rng, it_lo, it_hi);
                            if (use_taper) pv = mul_rn(pv, tpv[r][c]);                     // MCMC.py:778
                        }
                        if (use_cw) pv = mul_rn(pv, cw[r][c]);                            // MCMC.py:1279-1282
                        trow[bj] = add_rn(trow[bj], pv);                                  // MCMC.py:1285-1290
                    }
                }
                // blocks wider than 96 columns (non-default block tables): the remaining chunks, one cell at a time
                for (int bj = lane + 96; bj < bw; bj += 32) {
                    const int64_t idx = (int64_t)(s.x0 + bi) * W + s.y0 + bj;
                    if (__ldg(d.flags + idx) & FLAG_GATE) {
                        double pv;
                        if (INJECT_F) pv = __ldcg(f_inj + fy * f_pitch + s.my0 + bj);
                        else {
                            pv = field_value<false>(fv, fy, s.my0 + bj, rng, it_lo, it_hi);
                            if (use_taper) pv = mul_rn(pv, __ldg(fv.taper + fy * fv.w + s.my0 + bj));
                        }
                        if (use_cw) pv = mul_rn(pv, __ldg(d.crf_weight + idx));
                        trow[bj] = add_rn(trow[bj], pv);
                    }
                }
            }
        }
    }
    __syncthreads();
    // split mode (tail_kernel): every thread has read its cells of the injected field - its ring slot may be refilled
    if (consumed_flag && threadIdx.x == 0) atomicExch(consumed_flag, consumed_val);
    pc.mark(5);

    // phase B: residual on the block, loss delta, thickness guard                       MCMC.py:1292-1329
    double delta[GMC_TAIL_NACC];                                   // canonical slots, see canon_block_sum
#pragma unroll
    for (int a = 0; a < GMC_TAIL_NACC; ++a) delta[a] = 0.0;
    int trip = 0;
    int bad = 0;
    // workers: 224 threads (7 warps) of a 256-thread CTA, 448 (14 warps) of a 512-thread one; with HELPER the last warp
    // prepares the next step instead (and warp 14 of the wide CTA idles here)
    constexpr int B_THREADS = GMC_TAIL_WORKERS;
    if (HELPER && threadIdx.x >= GMC_STEP_THREADS - 32) prepare_step(d, rng, next->it, *next->sc, *next->pair, *next->tab, next->vec);
    // (unrolling this loop by two so that the loads of two trips overlap measured 4 % SLOWER at 128 registers: 6.12 vs 6.38 M
    // chain-steps/s at 256 x 500^2 - profiles/r2/step_ab.txt)
    for (int e = (threadIdx.x >= B_THREADS) ? bh * bw : threadIdx.x; e < bh * bw; e += B_THREADS, ++trip) {
        const int bi = dbw.div(e), bj = e - bi * bw;
        const int i = s.x0 + bi, j = s.y0 + bj;
        const double* tc = tile + (bi + 1) * tp + (bj + 1);
        // np.gradient: one-sided at the grid edge (neighbour index clamped, divisor res), central elsewhere
        const int jl = max(j - 1, 0), jr = min(j + 1, W - 1);
        const int iu = max(i - 1, 0), id = min(i + 1, H - 1);
        const bool ex = (j == 0) || (j == W - 1), ey = (i == 0) || (i == H - 1);
        const double denx = ex ? d.res : d.two_res, rdx = ex ? d.r_res : d.r_two_res;
        const double deny = ey ? d.res : d.two_res, rdy = ey ? d.r_res : d.r_two_res;
        const int64_t r = (int64_t)i * W;
        // all global loads first: {surf, velx} / {surf, vely} / {dhdt, smb} pairs are one 16 B load each
        const double2 xr = __ldg(d.sv + r + jr), xl = __ldg(d.sv + r + jl);
        const double2 yd = __ldg(d.sy + (int64_t)id * W + j), yu = __ldg(d.sy + (int64_t)iu * W + j);
        const double2 hs = __ldg(d.ds + r + j);
        const double sc0 = __ldg(d.surf + r + j);
        const uint8_t fl = __ldg(d.flags + r + j);
        const double rold = __ldcg(mcres + r + j);
        const double fr = mul_rn(xr.y, sub_rn(xr.x, tc[jr - j]));
        const double fl_ = mul_rn(xl.y, sub_rn(xl.x, tc[jl - j]));
        const double dx = div_const(sub_rn(fr, fl_), denx, rdx);
        const double fd = mul_rn(yd.y, sub_rn(yd.x, tc[(id - i) * tp]));
        const double fu = mul_rn(yu.y, sub_rn(yu.x, tc[(iu - i) * tp]));
        const double dy = div_const(sub_rn(fd, fu), deny, rdy);
        const double rnew = sub_rn(add_rn(add_rn(dx, dy), hs.x), hs.y);
        newres[e] = rnew;
        if (fl & FLAG_MC) {
            const double dl = (rnew == rnew && rold == rold) ? (rnew - rold) * (rnew + rold) : sq_or_zero(rnew) - sq_or_zero(rold);
            if (GMC_TAIL_NACC == 1 || (trip & 1) == 0) delta[0] += dl;
            else delta[GMC_TAIL_NACC - 1] += dl;
        }
        if ((fl & FLAG_GATE) && sub_rn(sc0, tc[0]) <= 0.0) bad = 1;
    }
    const double dsum = canon_block_sum<GMC_TAIL_NACC, GMC_TAIL_WORKERS>(delta, scratch);
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
    pc.mark(6);
    const int accept = sc->accept;
    if (loss_next_out && threadIdx.x == 0) *loss_next_out = scratch[35];
    if (accept) {
        ssq = scratch[34];
        for (int e = threadIdx.x; e < bh * bw; e += GMC_STEP_THREADS) {
            const int bi = dbw.div(e), bj = e - bi * bw;
            const int64_t idx = (int64_t)(s.x0 + bi) * W + (s.y0 + bj);
            __stcg(bed + idx, tile[(bi + 1) * tp + (bj + 1)]);
            __stcg(mcres + idx, newres[e]);
            // coverage count (MCMC.py:1349-1352): a fire-and-forget L2 reduction instead of a load + store round trip
            if (resampled && (__ldg(d.flags + idx) & FLAG_GATE)) atomicAdd(resampled + idx, 1);
        }
    }
    __syncthreads();   // write-back visible to the next iteration's tile load; smem free for reuse
    pc.mark(7);
}

// full masked nansum of the tracked residual (fixed order)
__device__ double resync_ssq(const GmcDev& d, const double* mcres, double* scratch) {
    const int64_t n = (int64_t)d.H * d.W;
    double acc[GMC_SUM_NACC];
#pragma unroll
    for (int a = 0; a < GMC_SUM_NACC; ++a) acc[a] = 0.0;
    int trip = 0;
    for (int64_t k = threadIdx.x; k < n; k += GMC_STEP_THREADS, ++trip) {
        const double v = __ldcg(mcres + k);
        const double sq = ((__ldg(d.flags + k) & FLAG_MC) && v == v) ? v * v : 0.0;
        if (GMC_SUM_NACC == 1 || (trip & 1) == 0) acc[0] += sq;
        else acc[GMC_SUM_NACC - 1] += sq;
    }
    return canon_block_sum<GMC_SUM_NACC, GMC_STEP_THREADS>(acc, scratch);
}

// Everything a step needs before its field can be synthesised, computed by ONE warp: the step's scalars (block size, scale,
// nugget, range, centre, acceptance uniform, spectral constants), the block size's record and its small tables.  Nothing
// here depends on the chain state, so run_kernel lets a helper warp prepare step k+1 while the other warps are in the
// latency-bound residual phase of step k (the tables of step k are dead by then).
__device__ __noinline__ void prepare_step(const GmcDev& d, const Philox& rng, uint64_t it, StepScalars& sc, GmcPair& s_pair,
                                          StepTables& s_tab, bool vec) {
    const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
    const int l = threadIdx.x & 31;
    // the five Philox blocks of the step's scalars are drawn by five lanes in parallel, then gathered by lane 0
    uint4 r = make_uint4(0, 0, 0, 0);
    if (l < 5) r = rng(l < 3 ? (uint32_t)l : (uint32_t)(l - 3), it_lo, it_hi, l < 3 ? GMC_STREAM_RF_SCALARS : GMC_STREAM_CHAIN);
    uint4 g[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
        g[k] = make_uint4(__shfl_sync(0xffffffffu, r.x, k), __shfl_sync(0xffffffffu, r.y, k),
                          __shfl_sync(0xffffffffu, r.z, k), __shfl_sync(0xffffffffu, r.w, k));
    if (l == 0) {
        const GmcFieldModel& fm = d.fm;
        const uint4 r0 = g[0], r1 = g[1], r2 = g[2], c0 = g[3], c1 = g[4];
        // RandField stream: block size, scale, nugget, range(s)                   MCMC.py:755, 200-207
        sc.pair = (int)bounded_u64(r0.x, r0.y, (uint64_t)d.n_pairs);
        sc.scale = div_rn(add_rn(fm.scale_min, mul_rn(sub_rn(fm.scale_max, fm.scale_min), u01_halfopen(r0.z, r0.w))), 3.0);
        sc.nug = add_rn(0.0, mul_rn(fm.nugget_max, u01_halfopen(r1.x, r1.y)));
        sc.range_x = add_rn(fm.range_min_x, mul_rn(sub_rn(fm.range_max_x, fm.range_min_x), u01_halfopen(r1.z, r1.w)));
        if (fm.isotropic) sc.range_y = sc.range_x;
        else sc.range_y = add_rn(fm.range_min_y, mul_rn(sub_rn(fm.range_max_y, fm.range_min_y), u01_halfopen(r2.x, r2.y)));
        // chain stream: block centre (uniform over the allowed cells) and the acceptance uniform   MCMC.py:1253-1261, 1336
        if (d.n_centre_cells > 0) {
            const int32_t cell = d.centre_cells[bounded_u64(c0.x, c0.y, (uint64_t)d.n_centre_cells)];
            sc.ix = cell / d.W;
            sc.iy = cell - sc.ix * d.W;
        } else {
            sc.ix = (int)bounded_u64(c0.x, c0.y, (uint64_t)d.H);
            sc.iy = (int)bounded_u64(c0.z, c0.w, (uint64_t)d.W);
        }
        sc.u = u01_halfopen(c1.x, c1.y);
        s_pair = d.pairs[sc.pair];             // one trip: sizes, table offsets and both FFT plans
        sc.h = s_pair.h;
        sc.w = s_pair.w;
        block_window(sc, d.H, d.W, vec);
        sc.spec = make_spec(fm, sc.range_x, sc.range_y);
    }
    __syncwarp();
    stage_tables(d, s_pair, s_tab, l, 32);
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char gmc_smem[];

// sched (may be NULL): work distribution for launches with more chains than resident CTAs.  Without it CTA b advances
// chain b by all n_steps.  With it the iterations are cut into chunks of `chunk` and the (chunk, chain) items - chunk-major,
// so that a chain's previous chunk was handed out gridDim.x * ... items earlier - are drawn from the counter sched[0]; an
// item waits until its chain has completed the previous chunk (sched[1 + chain]).  Every CTA of the grid is resident and an
// item only ever waits for an item drawn earlier, so the waits cannot deadlock.  A chain migrates between CTAs (and SMs):
// its state is written with L2 stores + __threadfence() before the completion count is published, and read back only
// through L2 (TMA bulk copies, ld.cg).
#ifdef GMC_STEP_MAXNREG   // A/B: cap the 256-thread kernel below 128 registers so that small kernels fit next to two step CTAs
__global__ void __maxnreg__(GMC_STEP_THREADS == 256 ? GMC_STEP_MAXNREG : 128)
#else
__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
#endif
    run_kernel(GmcDev d, double* bed_all, double* mcres_all, double* ssq_all, const uint64_t* __restrict__ seeds,
               uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache,
               int64_t cache_stride, int64_t cache_offset, int32_t* resampled_all, int resync_every, int tile_off,
               long long* phase_acc, int C, int* sched, int chunk, int* err, unsigned spin_limit) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc, sc_next;
    __shared__ GmcPair s_pair;
    __shared__ StepTables s_tab;
    __shared__ __align__(8) uint64_t tile_bar;                // completion barrier of the bulk copies that stage the bed tile
    __shared__ long long s_item;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int64_t plane = (int64_t)d.H * d.W;
    PhaseClock pc;
    pc.acc = phase_acc;
    pc.start();
    if (threadIdx.x == 0) mbar_init(&tile_bar, 1);
    const bool vec = (d.W % 2 == 0) && ((reinterpret_cast<uintptr_t>(bed_all) & 15) == 0);
    const int n_chunks = sched ? (n_steps + chunk - 1) / chunk : 1;
    const long long n_items = (long long)n_chunks * C;
    unsigned tile_uses = 0;                                   // phase parity of tile_bar

    for (long long item = blockIdx.x;; item += gridDim.x) {
        if (sched) {
            if (threadIdx.x == 0) {
                const long long it2 = atomicAdd(reinterpret_cast<unsigned int*>(sched), 1u);
                if (it2 < n_items) {
                    const int cc = (int)(it2 % C), jj = (int)(it2 / C);
                    volatile int* done = sched + 1 + cc;
                    unsigned spins = 0;
                    while (*done < jj) {
                        if (*(volatile int*)err) break;       // a wait already gave up: the launch is void, do not wait again
                        __nanosleep(200);
                        if (++spins > spin_limit) {           // never hang the device - and never carry on silently: the
                            *(volatile int*)err = GMC_DEVERR_WAIT_TIMEOUT;   // host turns the flag into GMC_ECUDA
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
        double* bed = bed_all + c * plane;
        double* mcres = mcres_all + c * plane;
        int32_t* resampled = resampled_all ? resampled_all + c * plane : nullptr;
        const Philox rng(seeds[c]);
        double ssq = __ldcg(ssq_all + c);

        // step k0 is prepared up front; every later step by the helper warp during the previous step's residual phase
        if (threadIdx.x < 32) prepare_step(d, rng, iter0 + (uint64_t)k0, sc_next, s_pair, s_tab, vec);
        __syncthreads();
        if (threadIdx.x == 0) sc = sc_next;
        __syncthreads();
        // iterations until the next re-sum of the tracked residual (it % resync_every == 0), counted down instead of a
        // 64-bit modulo per step
        int64_t to_resync = -1;
        if (resync_every > 0)
            to_resync = (int64_t)((uint64_t)resync_every - (iter0 + (uint64_t)k0) % (uint64_t)resync_every) % resync_every;
        for (int k = k0; k < k1; ++k) {
            const uint64_t it = iter0 + (uint64_t)k;
            const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
            if (to_resync == 0) {
                ssq = resync_ssq(d, mcres, scratch);
                to_resync = resync_every;
            }
            --to_resync;
            stage_block_async(sc, d.H, d.W, bed, mcres, buf + tile_off, vec, vec ? &tile_bar : nullptr);
            pc.mark(0);
            const FieldView fv = synth_field<false>(d, buf, scratch, s_pair, s_tab, sc.scale, sc.nug, sc.spec, rng, it_lo, it_hi,
                                                    nullptr, nullptr, nullptr, true, pc);
            // tile after the field; the new residuals reuse the field's storage once the tile is built (f is dead by then)
            const NextStep next = {&sc_next, &s_pair, &s_tab, it + 1, vec};
            step_tail<false, true>(d, &sc, scratch, fv, nullptr, 0, rng, it_lo, it_hi, buf + tile_off, buf, bed, mcres, ssq,
                                   resampled, nullptr, pc, &next, vec ? &tile_bar : nullptr, tile_uses & 1u, err);
            ++tile_uses;
            if (threadIdx.x == 0) {
                const int64_t slot = (int64_t)c * cache_stride + cache_offset + k;
                if (loss_cache) loss_cache[slot] = div_rn(ssq, d.two_sigma2);
                if (step_cache) step_cache[slot] = (uint8_t)sc.accept;
                if (blocks_cache) reinterpret_cast<int4*>(blocks_cache)[slot] = make_int4(sc.ix, sc.iy, sc.h, sc.w);
                sc = sc_next;                  // nobody reads sc between the tail's last barrier and the one below
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) __stcg(ssq_all + c, ssq);
        if (!sched) break;
        __threadfence();                                      // this thread's state writes are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(sched + 1 + c, j + 1);   // ... before the chain's next chunk may start anywhere
    }
}

#ifndef GMC_STEP_RUN_ONLY   // the replay / field / randomization-method kernels exist for the default CTA size only
// ---------------------------------------------------------------------------------------------------------------
// Split mode: TWO CTAs per chain, pipelined across steps.  Nothing a field needs depends on the chain state (Philox is
// counter addressed), so a PRODUCER CTA synthesises the tapered proposal fields of steps k, k+1, ... into a small ring in
// global memory (L2 resident: depth x max block doubles per chain) while the CONSUMER CTA runs the Metropolis tail of step
// k on the field it finds there (the injected-field instantiation of step_tail).  A chain then advances at the pace of the
// slower half (~half a fused step) instead of their sum - used when a launch has at most half as many chains as the GPU has
// CTA slots, i.e. exactly when one CTA per chain leaves the GPU under-filled.  Same arithmetic in the same order as
// run_kernel: bit-identical trajectories (tests/test_gpu_fullsize.py).  flags[2c] = fields produced, flags[2c+1] = fields
// consumed by chain c in this launch; both sides only ever wait for a resident CTA, with a bound (error flag, no hang).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_wait(volatile int* flag, int need, int* err, unsigned spin_limit) {
    unsigned spins = 0;
    while (*flag < need) {
        if (*(volatile int*)err) break;                       // a wait already gave up: the launch is void, do not wait again
        __nanosleep(100);
        if (++spins > spin_limit) {
            *(volatile int*)err = GMC_DEVERR_WAIT_TIMEOUT;
            __threadfence_system();
            break;
        }
    }
    __threadfence();
}

__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
    field_producer_kernel(GmcDev d, const uint64_t* __restrict__ seeds, uint64_t iter0, int n_steps, double* ring, int64_t fstride,
                          int depth, int* flags, int* err, unsigned spin_limit) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc;
    __shared__ GmcPair s_pair;
    __shared__ StepTables s_tab;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int c = blockIdx.x;
    const Philox rng(seeds[c]);
    PhaseClock pc;
    pc.acc = nullptr;
    if (threadIdx.x < 32) prepare_step(d, rng, iter0, sc, s_pair, s_tab, false);
    __syncthreads();
    for (int k = 0; k < n_steps; ++k) {
        const uint64_t it = iter0 + (uint64_t)k;
        const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
        const FieldView fv = synth_field<false>(d, buf, scratch, s_pair, s_tab, sc.scale, sc.nug, sc.spec, rng, it_lo, it_hi, nullptr,
                                                nullptr, nullptr, true, pc);
        const int h = s_pair.h, w = s_pair.w;                 // in registers: warp 0 overwrites the records below
        // the slot of step k was last used by step k - depth: wait until the consumer has read that field
        if (threadIdx.x == 32 && k >= depth) split_wait(flags + 2 * c + 1, k - depth + 1, err, spin_limit);
        __syncthreads();
        // warp 0 prepares step k+1 (scalars, block record, tables - those of step k are dead after the row pass) while the
        // other warps write the field out; it joins them for its own share afterwards
        if (threadIdx.x < 32 && k + 1 < n_steps) prepare_step(d, rng, it + 1, sc, s_pair, s_tab, false);
        double* slot = ring + ((int64_t)c * depth + (k % depth)) * fstride;
        const FastDiv dw(w);
        for (int e = threadIdx.x; e < h * w; e += GMC_STEP_THREADS) {
            const int y = dw.div(e), x = e - y * w;
            double f = field_value<false>(fv, y, x, rng, it_lo, it_hi);
            if (fv.taper) f = mul_rn(f, __ldg(fv.taper + e));                               // MCMC.py:778
            __stcg(slot + e, f);
        }
        __threadfence();                                      // this thread's field values are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(flags + 2 * c, k + 1);   // ... before the consumer is told
    }
}

__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
    tail_kernel(GmcDev d, double* bed_all, double* mcres_all, double* ssq_all, const uint64_t* __restrict__ seeds, uint64_t iter0,
                int n_steps, double* loss_cache, uint8_t* step_cache, int32_t* blocks_cache, int64_t cache_stride,
                int64_t cache_offset, int32_t* resampled_all, int resync_every, int tile_off, const double* ring, int64_t fstride,
                int depth, int* flags, int* err, unsigned spin_limit) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc, sc_next;
    __shared__ GmcPair s_pair;
    __shared__ StepTables s_tab;
    __shared__ __align__(8) uint64_t tile_bar;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int64_t plane = (int64_t)d.H * d.W;
    const int c = blockIdx.x;
    PhaseClock pc;
    pc.acc = nullptr;
    if (threadIdx.x == 0) mbar_init(&tile_bar, 1);
    const bool vec = (d.W % 2 == 0) && ((reinterpret_cast<uintptr_t>(bed_all) & 15) == 0);
    double* bed = bed_all + c * plane;
    double* mcres = mcres_all + c * plane;
    int32_t* resampled = resampled_all ? resampled_all + c * plane : nullptr;
    const Philox rng(seeds[c]);
    double ssq = __ldcg(ssq_all + c);
    if (threadIdx.x < 32) prepare_step(d, rng, iter0, sc_next, s_pair, s_tab, vec);
    __syncthreads();
    if (threadIdx.x == 0) sc = sc_next;
    __syncthreads();
    int64_t to_resync = -1;
    if (resync_every > 0) to_resync = (int64_t)((uint64_t)resync_every - iter0 % (uint64_t)resync_every) % resync_every;
    const FieldView fv = {};
    for (int k = 0; k < n_steps; ++k) {
        const uint64_t it = iter0 + (uint64_t)k;
        const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
        if (to_resync == 0) {
            ssq = resync_ssq(d, mcres, scratch);
            to_resync = resync_every;
        }
        --to_resync;
        stage_block_async(sc, d.H, d.W, bed, mcres, buf + tile_off, vec, vec ? &tile_bar : nullptr);
        if (threadIdx.x == 0) split_wait(flags + 2 * c, k + 1, err, spin_limit);      // the producer has published field k
        __syncthreads();
        const double* f = ring + ((int64_t)c * depth + (k % depth)) * fstride;
        const NextStep next = {&sc_next, &s_pair, &s_tab, it + 1, vec};
        step_tail<true, true>(d, &sc, scratch, fv, f, sc.w, rng, it_lo, it_hi, buf + tile_off, buf, bed, mcres, ssq, resampled,
                              nullptr, pc, &next, vec ? &tile_bar : nullptr, (unsigned)k & 1u, err, flags + 2 * c + 1, k + 1);
        if (threadIdx.x == 0) {
            const int64_t slot = (int64_t)c * cache_stride + cache_offset + k;
            if (loss_cache) loss_cache[slot] = div_rn(ssq, d.two_sigma2);
            if (step_cache) step_cache[slot] = (uint8_t)sc.accept;
            if (blocks_cache) reinterpret_cast<int4*>(blocks_cache)[slot] = make_int4(sc.ix, sc.iy, sc.h, sc.w);
            sc = sc_next;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) __stcg(ssq_all + c, ssq);
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
    // 16-byte tile staging needs an even offset of the tile inside the dynamic shared memory as well
    const bool vec = (d.W % 2 == 0) && ((reinterpret_cast<uintptr_t>(bed_all) & 15) == 0) && (((int64_t)hmax * wmax) % 2 == 0);
    if (threadIdx.x == 0) {
        sc.h = hw[2 * c];
        sc.w = hw[2 * c + 1];
        sc.ix = centre[2 * c];
        sc.iy = centre[2 * c + 1];
        sc.u = u[c];
        sc.pair = -1;
        block_window(sc, d.H, d.W, vec);
    }
    __syncthreads();
    double ssq = ssq_all[c];
    const Philox rng(0ull);
    stage_block_async(sc, d.H, d.W, bed_all + c * plane, mcres_all + c * plane, buf + (int64_t)hmax * wmax, vec);
    FieldView fv = {};
    PhaseClock pc;
    pc.acc = nullptr;
    step_tail<true>(d, &sc, scratch, fv, f_all + c * f_stride, sc.w, rng, 0u, 0u, buf + (int64_t)hmax * wmax, buf,
                    bed_all + c * plane, mcres_all + c * plane, ssq, resampled_all ? resampled_all + c * plane : nullptr,
                    loss_next_out ? loss_next_out + c : nullptr, pc);
    if (threadIdx.x == 0) {
        ssq_all[c] = ssq;
        if (accepted_out) accepted_out[c] = (uint8_t)sc.accept;
        if (loss_out) loss_out[c] = div_rn(ssq, d.two_sigma2);
    }
}

template <bool INJECT>
__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
    field_kernel(GmcDev d, const int32_t* __restrict__ pair, const double* __restrict__ scale, const double* __restrict__ nug,
                 const double* __restrict__ range_x, const double* __restrict__ range_y, const double* __restrict__ z_re,
                 const double* __restrict__ z_im, const double* __restrict__ z_nug, const uint64_t* __restrict__ seeds,
                 uint64_t iter, int apply_taper, double* __restrict__ f_out, int64_t stride) {
    __shared__ double scratch[40];
    __shared__ GmcPair s_pair;
    __shared__ StepTables s_tab;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int i = blockIdx.x;
    const int p = pair[i];
    if (threadIdx.x == 0) s_pair = d.pairs[p];
    __syncthreads();
    stage_tables(d, s_pair, s_tab);
    __syncthreads();
    const Philox rng(INJECT ? 0ull : seeds[i]);
    const uint32_t it_lo = (uint32_t)iter, it_hi = (uint32_t)(iter >> 32);
    PhaseClock pc;
    pc.acc = nullptr;
    const SpecParams sp = make_spec(d.fm, range_x[i], range_y[i]);
    const FieldView fv = synth_field<INJECT>(d, buf, scratch, s_pair, s_tab, scale[i], nug[i], sp, rng, it_lo, it_hi,
                                             INJECT ? z_re + i * stride : nullptr, INJECT ? z_im + i * stride : nullptr,
                                             INJECT ? z_nug + i * stride : nullptr, apply_taper != 0, pc);
    const int h = s_pair.h, w = s_pair.w;
    const FastDiv dw(w);
    for (int e = threadIdx.x; e < h * w; e += GMC_STEP_THREADS) {
        const int y = dw.div(e), x = e - y * w;
        double f = field_value<INJECT>(fv, y, x, rng, it_lo, it_hi);
        if (fv.taper) f = mul_rn(f, __ldg(fv.taper + e));                               // MCMC.py:778
        f_out[i * stride + e] = f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// A5: randomization-method proposal (RandField.get_random_field, MCMC.py:625-687 -> gstools SRF / RandMeth, mode_no
// wave vectors):  field(p) = sqrt(var / N) * sum_m [ z1_m cos(k_m . p) + z2_m sin(k_m . p) ],  var = 1, p = (x res, y res).
// On the regular block grid the phase separates, k.p = kx x res + ky y res, so with a = kx x res, b = ky y res
//   z1 cos(a+b) + z2 sin(a+b) = cos b (z1 cos a + z2 sin a) + sin b (z2 cos a - z1 sin a)
// and the field is the product of an [h x 2N] table (cos b, sin b) with a [2N x w] table (P, Q): 2 N (h + w) sincos
// instead of 2 N h w, and a rank-2N update held in registers (5 x 5 outputs per thread, 16 x 16 threads).
// The wave vectors are isotropic-frame samples k' = r (cos phi, sin phi): phi uniform, r by inversion of the model's
// radial spectral distribution in two dimensions (gstools model definitions, rescale factors included):
//   Gaussian     rho = exp(-(pi/4)(d/l)^2)                        r = sqrt(pi)/l * sqrt(-ln(1-u))
//   Exponential  rho = exp(-d/l)                                  r = sqrt(1/(1-u)^2 - 1) / l
//   Matern       rho = 2^(1-nu)/Gamma(nu) (sqrt(nu) d/l)^nu K_nu   r = sqrt(nu ((1-u)^(-1/nu) - 1)) / l
// mapped to the grid frame by the model's rotation and anisotropy: k = R(theta) diag(1, l1/l2) k'.
// ---------------------------------------------------------------------------------------------------------------
constexpr int RM_CHUNK = 16;                 // modes per shared-memory table chunk
constexpr int RM_EDGE = 80;                  // rows / columns of one register-tile pass
constexpr int RM_T = RM_EDGE / 16;           // outputs per thread per axis
static_assert(GMC_STEP_THREADS == 256, "the randomization-method register tiling assumes 16 x 16 threads");

struct RandMethParams {
    int model, n_modes;
    double nu, len, inv_anis, cos_t, sin_t;
};

__device__ __forceinline__ RandMethParams make_randmeth(const GmcFieldModel& fm, int n_modes, double range_x, double range_y,
                                                        double angle_deg) {
    RandMethParams rp;
    rp.model = fm.model;
    rp.n_modes = n_modes;
    rp.nu = fm.smoothness;
    // len_scale = [range1, range2] / sqrt(3) | / 3 | / 2 (MCMC.py:657-676); main length l1, anisotropy ratio l2 / l1
    const double dv = (fm.model == GMC_GAUSSIAN) ? sqrt(3.0) : (fm.model == GMC_EXPONENTIAL ? 3.0 : 2.0);
    const double l1 = div_rn(range_x, dv), l2 = div_rn(range_y, dv);
    rp.len = l1;
    rp.inv_anis = div_rn(l1, l2);
    sincos(div_rn(mul_rn(angle_deg, 3.141592653589793), 180.0), &rp.sin_t, &rp.cos_t);   // angles = angle*np.pi/180
    return rp;
}

// mode m of the step: out = (kx, ky, z1, z2), wave vector in rad per length unit of `res`
__device__ __noinline__ void rm_mode(const RandMethParams& rp, const Philox& rng, uint32_t m, uint32_t it_lo, uint32_t it_hi,
                                     double* out) {
    const uint4 a = rng(m, it_lo, it_hi, GMC_STREAM_RM_MODE);
    const double u = u01_open(a.x, a.y);
    double s, c;
    sincospi(2.0 * u01_open(a.z, a.w), &s, &c);
    double r;
    if (rp.model == GMC_GAUSSIAN) r = sqrt(-log1p(-u)) * 1.7724538509055159 / rp.len;
    else if (rp.model == GMC_EXPONENTIAL) r = sqrt(u * (2.0 - u)) / (1.0 - u) / rp.len;
    else r = sqrt(rp.nu * expm1(-log1p(-u) / rp.nu)) / rp.len;
    const double k0 = r * c, k1 = r * s * rp.inv_anis;
    out[0] = rp.cos_t * k0 - rp.sin_t * k1;
    out[1] = rp.sin_t * k0 + rp.cos_t * k1;
    box_muller(rng(m, it_lo, it_hi, GMC_STREAM_RM_AMP), out[2], out[3]);
}

// Synthesises the field (times `scale`) into buf as F[y * w + x]; the tables live at buf + tab_off (4 RM_CHUNK RM_EDGE
// doubles, past the largest field).  Nugget noise (times scale, as the reference scales the whole gstools field) and
// taper are applied by the consumer through the returned view.
template <bool INJECT>
__device__ FieldView synth_randmeth(const GmcDev& d, double* buf, int tab_off, const GmcPair& pr, double res, double scale,
                                    double nug, const RandMethParams& rp, const Philox& rng, uint32_t it_lo, uint32_t it_hi,
                                    const double* __restrict__ modes_inj, const double* __restrict__ z_nug, bool apply_taper) {
    __shared__ double s_mode[RM_CHUNK][4];
    const int h = pr.h, w = pr.w;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double* A = buf + tab_off;                    // [2 RM_CHUNK][RM_EDGE]: cos b, sin b
    double* B = A + 2 * RM_CHUNK * RM_EDGE;       // [2 RM_CHUNK][RM_EDGE]: P, Q
    const double amp = scale * sqrt(1.0 / (double)rp.n_modes);
    for (int py = 0; py < h; py += RM_EDGE)
        for (int px = 0; px < w; px += RM_EDGE) {
            const int hp = min(RM_EDGE, h - py), wp = min(RM_EDGE, w - px);
            const FastDiv dline(hp + wp);
            double acc[RM_T][RM_T];
#pragma unroll
            for (int i = 0; i < RM_T; ++i)
#pragma unroll
                for (int j = 0; j < RM_T; ++j) acc[i][j] = 0.0;
            for (int m0 = 0; m0 < rp.n_modes; m0 += RM_CHUNK) {
                if (threadIdx.x < RM_CHUNK) {
                    const int m = m0 + threadIdx.x;
                    double v[4] = {0.0, 0.0, 0.0, 0.0};
                    if (m < rp.n_modes) {
                        if (INJECT) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) v[q] = modes_inj[4 * m + q];
                        } else rm_mode(rp, rng, (uint32_t)m, it_lo, it_hi, v);
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) s_mode[threadIdx.x][q] = v[q];
                }
                __syncthreads();
                for (int e = threadIdx.x; e < RM_CHUNK * (hp + wp); e += GMC_STEP_THREADS) {
                    const int mm = dline.div(e), p = e - mm * (hp + wp);
                    double sn, cs;
                    if (p < hp) {
                        sincos(s_mode[mm][1] * ((double)(py + p) * res), &sn, &cs);
                        A[(2 * mm) * RM_EDGE + p] = cs;
                        A[(2 * mm + 1) * RM_EDGE + p] = sn;
                    } else {
                        const int x = p - hp;
                        sincos(s_mode[mm][0] * ((double)(px + x) * res), &sn, &cs);
                        const double z1 = s_mode[mm][2], z2 = s_mode[mm][3];
                        B[(2 * mm) * RM_EDGE + x] = z1 * cs + z2 * sn;
                        B[(2 * mm + 1) * RM_EDGE + x] = z2 * cs - z1 * sn;
                    }
                }
                __syncthreads();
#pragma unroll 2
                for (int k = 0; k < 2 * RM_CHUNK; ++k) {
                    double a[RM_T], b[RM_T];
#pragma unroll
                    for (int i = 0; i < RM_T; ++i) a[i] = A[k * RM_EDGE + ty + 16 * i];
#pragma unroll
                    for (int j = 0; j < RM_T; ++j) b[j] = B[k * RM_EDGE + tx + 16 * j];
#pragma unroll
                    for (int i = 0; i < RM_T; ++i)
#pragma unroll
                        for (int j = 0; j < RM_T; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
                }
            }
            // rows / columns past (hp, wp) multiplied stale table entries: never stored
#pragma unroll
            for (int i = 0; i < RM_T; ++i)
#pragma unroll
                for (int j = 0; j < RM_T; ++j) {
                    const int y = ty + 16 * i, x = tx + 16 * j;
                    if (y < hp && x < wp) buf[(py + y) * w + (px + x)] = amp * acc[i][j];
                }
        }
    __syncthreads();
    FieldView fv;
    fv.F = buf;
    fv.fpitch = w;
    fv.w = w;
    fv.sq_nug = (nug > 0.0) ? sqrt(nug) * scale : 0.0;
    fv.taper = apply_taper ? d.edge_masks + pr.mask_off : nullptr;
    fv.z_nug = z_nug;
    return fv;
}

template <bool INJECT>
__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
    field_randmeth_kernel(GmcDev d, int n_modes, double res, int tab_off, const int32_t* __restrict__ pair,
                          const double* __restrict__ scale, const double* __restrict__ nug, const double* __restrict__ range_x,
                          const double* __restrict__ range_y, const double* __restrict__ angle_deg,
                          const double* __restrict__ modes, const double* __restrict__ z_nug,
                          const uint64_t* __restrict__ seeds, uint64_t iter, int apply_taper, double* __restrict__ f_out,
                          int64_t stride) {
    __shared__ GmcPair s_pair;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int i = blockIdx.x;
    if (threadIdx.x == 0) s_pair = d.pairs[pair[i]];
    __syncthreads();
    const Philox rng(INJECT ? 0ull : seeds[i]);
    const uint32_t it_lo = (uint32_t)iter, it_hi = (uint32_t)(iter >> 32);
    const RandMethParams rp = make_randmeth(d.fm, n_modes, range_x[i], range_y[i], angle_deg[i]);
    const FieldView fv = synth_randmeth<INJECT>(d, buf, tab_off, s_pair, res, scale[i], nug[i], rp, rng, it_lo, it_hi,
                                                INJECT ? modes + (int64_t)i * n_modes * 4 : nullptr,
                                                INJECT ? z_nug + i * stride : nullptr, apply_taper != 0);
    const int h = s_pair.h, w = s_pair.w;
    for (int e = threadIdx.x; e < h * w; e += GMC_STEP_THREADS) {
        double f = field_value<INJECT>(fv, e / w, e % w, rng, it_lo, it_hi);
        if (fv.taper) f = mul_rn(f, __ldg(fv.taper + e));                               // MCMC.py:778
        f_out[i * stride + e] = f;
    }
}

// chain_crf.run with the randomization-method proposal: same step as run_kernel, other field source.  The field costs
// ~100x the FFT synthesis (as in the reference), so the block's HBM traffic is simply staged after it.
__global__ void __launch_bounds__(GMC_STEP_THREADS, GMC_STEP_MIN_CTAS)
    run_randmeth_kernel(GmcDev d, int n_modes, double res, int tab_off, double* bed_all, double* mcres_all, double* ssq_all,
                        const uint64_t* __restrict__ seeds, uint64_t iter0, int n_steps, double* loss_cache, uint8_t* step_cache,
                        int32_t* blocks_cache, int64_t cache_stride, int64_t cache_offset, int32_t* resampled_all,
                        int resync_every, int tile_off) {
    __shared__ double scratch[40];
    __shared__ StepScalars sc;
    __shared__ GmcPair s_pair;
    __shared__ RandMethParams s_rp;
    double* buf = reinterpret_cast<double*>(gmc_smem);
    const int c = blockIdx.x;
    const int64_t plane = (int64_t)d.H * d.W;
    double* bed = bed_all + c * plane;
    double* mcres = mcres_all + c * plane;
    int32_t* resampled = resampled_all ? resampled_all + c * plane : nullptr;
    const Philox rng(seeds[c]);
    double ssq = ssq_all[c];
    const bool vec = (d.W % 2 == 0) && ((reinterpret_cast<uintptr_t>(bed_all) & 15) == 0) && (tile_off % 2 == 0);
    PhaseClock pc;
    pc.acc = nullptr;

    for (int k = 0; k < n_steps; ++k) {
        const uint64_t it = iter0 + (uint64_t)k;
        const uint32_t it_lo = (uint32_t)it, it_hi = (uint32_t)(it >> 32);
        if (resync_every > 0 && it % (uint64_t)resync_every == 0) ssq = resync_ssq(d, mcres, scratch);
        if (threadIdx.x == 0) {
            const GmcFieldModel& fm = d.fm;
            const uint4 r0 = rng(0u, it_lo, it_hi, GMC_STREAM_RF_SCALARS), r1 = rng(1u, it_lo, it_hi, GMC_STREAM_RF_SCALARS);
            const uint4 r2 = rng(2u, it_lo, it_hi, GMC_STREAM_RF_SCALARS);
            const uint4 c0 = rng(0u, it_lo, it_hi, GMC_STREAM_CHAIN), c1 = rng(1u, it_lo, it_hi, GMC_STREAM_CHAIN);
            // RandField stream: block size, scale, nugget, range(s), angle                MCMC.py:755, 642-653
            sc.pair = (int)bounded_u64(r0.x, r0.y, (uint64_t)d.n_pairs);
            sc.scale = div_rn(add_rn(fm.scale_min, mul_rn(sub_rn(fm.scale_max, fm.scale_min), u01_halfopen(r0.z, r0.w))), 3.0);
            sc.nug = add_rn(0.0, mul_rn(fm.nugget_max, u01_halfopen(r1.x, r1.y)));
            sc.range_x = add_rn(fm.range_min_x, mul_rn(sub_rn(fm.range_max_x, fm.range_min_x), u01_halfopen(r1.z, r1.w)));
            double angle = 0.0;
            if (fm.isotropic) sc.range_y = sc.range_x;
            else {
                sc.range_y = add_rn(fm.range_min_y, mul_rn(sub_rn(fm.range_max_y, fm.range_min_y), u01_halfopen(r2.x, r2.y)));
                angle = mul_rn(180.0, u01_halfopen(r2.z, r2.w));
            }
            // chain stream: identical to run_kernel                                        MCMC.py:1253-1261, 1336
            if (d.n_centre_cells > 0) {
                const int32_t cell = d.centre_cells[bounded_u64(c0.x, c0.y, (uint64_t)d.n_centre_cells)];
                sc.ix = cell / d.W;
                sc.iy = cell - sc.ix * d.W;
            } else {
                sc.ix = (int)bounded_u64(c0.x, c0.y, (uint64_t)d.H);
                sc.iy = (int)bounded_u64(c0.z, c0.w, (uint64_t)d.W);
            }
            sc.u = u01_halfopen(c1.x, c1.y);
            s_pair = d.pairs[sc.pair];
            sc.h = s_pair.h;
            sc.w = s_pair.w;
            block_window(sc, d.H, d.W, vec);
            s_rp = make_randmeth(fm, n_modes, sc.range_x, sc.range_y, angle);
        }
        __syncthreads();
        const FieldView fv = synth_randmeth<false>(d, buf, tab_off, s_pair, res, sc.scale, sc.nug, s_rp, rng, it_lo, it_hi,
                                                   nullptr, nullptr, true);
        stage_block_async(sc, d.H, d.W, bed, mcres, buf + tile_off, vec);
        step_tail<false>(d, &sc, scratch, fv, nullptr, 0, rng, it_lo, it_hi, buf + tile_off, buf, bed, mcres, ssq, resampled,
                         nullptr, pc);
        if (threadIdx.x == 0) {
            const int64_t slot = (int64_t)c * cache_stride + cache_offset + k;
            if (loss_cache) loss_cache[slot] = div_rn(ssq, d.two_sigma2);
            if (step_cache) step_cache[slot] = (uint8_t)sc.accept;
            if (blocks_cache) reinterpret_cast<int4*>(blocks_cache)[slot] = make_int4(sc.ix, sc.iy, sc.h, sc.w);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) ssq_all[c] = ssq;
}

#endif  // GMC_STEP_RUN_ONLY

