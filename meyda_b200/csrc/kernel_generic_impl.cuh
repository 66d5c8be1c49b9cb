// kernel_generic_impl.cuh -- body of the block-per-frame kernels, compiled once per CTA size
// (MB_GENERIC_THREADS) by kernel_generic.cu.  Not a standalone header.

constexpr int kThreads = MB_GENERIC_THREADS;
constexpr int kWarps = kThreads / 32;
// MB_GENERIC_WARP_LOCAL (kernel_exact_warp.cu, with MB_GENERIC_THREADS 32): the warps of a CTA work on frames of their
// own, so "the thread's index in its frame's team" is its lane and the epilogue's team is one warp.
#ifdef MB_GENERIC_WARP_LOCAL
#define MB_TID ((int)(threadIdx.x & 31))
#else
#define MB_TID ((int)threadIdx.x)
#endif
// One-warp CTAs (bufferSize <= 512) need no block barrier: warp-level ordering is enough.
__device__ __forceinline__ void block_sync() {
    if constexpr (kWarps == 1) __syncwarp();
    else __syncthreads();
}
constexpr int kLaneBandWarps = 2;  // CTAs of up to this many warps (bufferSize <= 2048) sum bands / filters one per lane
constexpr int kScalarThread = kThreads > 64 ? 64 : 0;  // the lane that turns the frame's sums into the number features

// One padding element every 32 and every 1024 entries keeps both the unit-stride
// butterfly accesses and the bit-reversed gather of the split pass off a
// single shared-memory bank.
__device__ __forceinline__ int pidx(int i) { return i + (i >> 5) + (i >> 10); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return mbx2::cmul(a, b); }  // FMUL2 + FFMA2 (mb_fft.cuh)
// real-FFT split of one bin: a = X[k], b = X[M-k], w = exp(+2 pi i k / N) -> Z[k] = ((a + conj b) / 2 + w O) sc with
// O = (a - conj b) / (2 i); seven packed instructions
__device__ __forceinline__ float2 split_bin(float2 a, float2 b, float2 w, float sc) {
    const float2 cb = make_float2(b.x, -b.y);
    const float2 E = mbx2::add(a, cb), F = mbx2::sub(a, cb);                           // (2 er, 2 ei), (-2 oi, 2 orr)
    const float2 Oo = mbx2::mul(mbx2::swap(F), make_float2(0.5f, -0.5f));              // (orr, oi)
    return mbx2::mul(mbx2::fma(E, mbx2::bc(0.5f), mbx2::cmul(Oo, w)), mbx2::bc(sc));
}

// ---- exact mode: the reference's butterfly (lib/jsfft/fft.js:151-161) on float32-stored values, in
// float64 without FMA contraction, rounded to float32 where the reference stores into its Float32Array.
__device__ __forceinline__ int xidx(int i) { return i + (i >> 5); }  // one pad word per 32: conflict-free strides

__device__ __forceinline__ void exact_bfly(float &l_r, float &l_i, float &r_r, float &r_i, const double2 f) {
    const double SQRT1_2 = 0.70710678118654752440;
    const double lr = (double)l_r, li = (double)l_i, xr = (double)r_r, xi = (double)r_i;
    const double rr = __dsub_rn(__dmul_rn(f.x, xr), __dmul_rn(f.y, xi));
    const double ri = __dadd_rn(__dmul_rn(f.y, xr), __dmul_rn(f.x, xi));
    l_r = (float)__dmul_rn(SQRT1_2, __dadd_rn(lr, rr));
    l_i = (float)__dmul_rn(SQRT1_2, __dadd_rn(li, ri));
    r_r = (float)__dmul_rn(SQRT1_2, __dsub_rn(lr, rr));
    r_i = (float)__dmul_rn(SQRT1_2, __dsub_rn(li, ri));
}

// Q consecutive radix-2 DIT stages (widths w, 2w, .. 2^(Q-1) w) of an n-point array fused in registers:
// the same butterflies in the same order of stages as FFT_2_Iterative, each result rounded to float32
// exactly where the stage-by-stage loop would have stored it, so the outcome is bit-identical.
template <int Q>
__device__ __forceinline__ void exact_pass(float *xre, float *xim, const double2 *__restrict__ tw_exact, int n,
                                           int log2w) {
    constexpr int R = 1 << Q;
    const int w = 1 << log2w;
    for (int idx = MB_TID; idx < (n >> Q); idx += kThreads) {
        const int j = idx & (w - 1);
        const int base = ((idx >> log2w) << (log2w + Q)) + j;
        float re[R], im[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            re[t] = xre[xidx(base + t * w)];
            im[t] = xim[xidx(base + t * w)];
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int h = 1 << q;  // partner distance in units of w; this stage has width h * w
            const double2 *__restrict__ tw = tw_exact + ((w << q) - 1);
#pragma unroll
            for (int t = 0; t < R; t++) {
                if (t & h) continue;
                exact_bfly(re[t], im[t], re[t + h], im[t + h], __ldg(&tw[j + (t & (h - 1)) * w]));
            }
        }
#pragma unroll
        for (int t = 0; t < R; t++) {
            xre[xidx(base + t * w)] = re[t];
            xim[xidx(base + t * w)] = im[t];
        }
    }
}

// stages of widths 2^first .. 2^(last-1) of an n-point array, three at a time
__device__ __forceinline__ void exact_stages(float *xre, float *xim, const double2 *__restrict__ tw_exact, int n,
                                             int first, int last) {
    int log2w = first;
    while (log2w < last) {
        const int q = last - log2w >= 3 ? 3 : last - log2w;
        if (q == 3) exact_pass<3>(xre, xim, tw_exact, n, log2w);
        else if (q == 2) exact_pass<2>(xre, xim, tw_exact, n, log2w);
        else exact_pass<1>(xre, xim, tw_exact, n, log2w);
        log2w += q;
        block_sync();
    }
}

// Q fused radix-2 DIF stages starting at span 2^log2s: 2^Q points per work item.
template <int Q>
__device__ __forceinline__ void fft_pass(float2 *work, const float2 *__restrict__ twM, int M, int log2M, int log2s) {
    constexpr int R = 1 << Q;
    const int items = M >> Q;
    const int log2sub = log2s - Q;
    const int sub = 1 << log2sub;
    for (int idx = MB_TID; idx < items; idx += kThreads) {
        const int j = idx & (sub - 1);
        const int b = (idx >> log2sub) << log2s;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; t++) v[t] = work[pidx(b + j + t * sub)];
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int h = R >> (q + 1);
            const int tw_shift = log2M - (log2s - q);  // M / (s >> q)
#pragma unroll
            for (int t = 0; t < R; t++) {
                if (t & h) continue;
                const int e = (j + (t & (h - 1)) * sub) << tw_shift;
                const float2 u = v[t], w = v[t + h];
                v[t] = mbx2::add(u, w);
                v[t + h] = cmul(mbx2::sub(u, w), __ldg(&twM[e]));
            }
        }
#pragma unroll
        for (int t = 0; t < R; t++) work[pidx(b + j + t * sub)] = v[t];
    }
}

// Block-wide sums: per-warp partials through shared memory, then every warp folds them with shuffles
// (kWarps <= 32), so all threads get the total.
__device__ __forceinline__ double block_sum(double v, double *scratch /*[kWarps]*/) {
    v = mb_warp_sum(v);
    if constexpr (kWarps == 1) return v;
    block_sync();
    if ((MB_TID & 31) == 0) scratch[MB_TID >> 5] = v;
    block_sync();
    const int lane = MB_TID & 31;
    return mb_warp_sum(lane < kWarps ? scratch[lane] : 0.0);
}

// Eight warp sums at once.  Step o = 16 / 8 / 4 halves the number of values a lane still carries (the lane keeps the
// ones its bit of o selects and hands the others to its partner), o = 2, 1 are plain butterflies: lane L ends with the
// total of v[(L >> 2) & 7], added in the same order as mb_warp_sum adds it.
__device__ __forceinline__ double mb_warp_sum8(const double (&v)[8]) {
    const int lane = MB_TID & 31;
    double w[4], x[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double keep = b3 ? w[i + 2] : w[i], send = b3 ? w[i] : w[i + 2];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    double y = (b2 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, b2 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}

__device__ __forceinline__ int block_sum_int(int v, int *scratch /*[kWarps]*/) {
    v = mb_warp_sum(v);
    if constexpr (kWarps == 1) return v;
    block_sync();
    if ((MB_TID & 31) == 0) scratch[MB_TID >> 5] = v;
    block_sync();
    const int lane = MB_TID & 31;
    return mb_warp_sum(lane < kWarps ? scratch[lane] : 0);
}

// float32-mode helpers: one MUFU each (<= 1 ulp / 2^-22 abs), subnormals handled (no .ftz)
__device__ __forceinline__ float sqrt_fast(float x) {
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float log2_fast(float x) {
    float r;
    asm("lg2.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct MomentAcc {
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, lg = 0;
    // mb_adaptive.cuh (float32 kernels of an adaptive plan): q0 = sum q(a_k), q4 = sum q(a_k) k^4 with
    // q(a) = min(1, (theta sigma / a)^2) from the exponent trick; cf = 0: off
    float q0 = 0.f, q4 = 0.f;
    int cf = 0;
    unsigned nq = 0;  // bins seen by add()
    template <bool FAST = false>
    __device__ __forceinline__ void add(float av, int k, bool want_log) {
        const double ad = (double)av, kd = (double)k;
        // (every fourth bin a thread visits, on an irregular pattern: a quarter of the spectrum is plenty for a bound)
        if (FAST && cf != 0 && (((nq++) * 5u >> 2) & 3u) == 0u) {
            float u;
            asm("mul.sat.f32 %0, %1, %1;" : "=f"(u) : "f"(__int_as_float(cf - __float_as_int(av))));
            const float k2 = (float)k * (float)k;
            q0 += u;
            q4 = fmaf(u, k2 * k2, q4);
        }
        double t = ad * kd;
        s0 += ad;
        s1 += t;
        t *= kd; s2 += t;
        t *= kd; s3 += t;
        t *= kd; s4 += t;
        // (one MUFU in every kernel family, exact mode included, so that a frame's flatness does not depend on which
        // kernel served it: 2^-22 per term, 2e-7 of the mean at worst; flatness is compared at 5e-6 in exact mode)
        if (want_log) lg += (double)log2_fast(av);
    }
};

// Everything after the amplitude spectrum: block reductions of the moment partials, rolloff, Bark bands,
// mel/log/DCT and the per-band outputs.  `amp` holds the N/2 amplitudes of the frame in shared memory.
struct Scratch {
    // mb_adaptive.cuh: what the decision needs.  The bands' / filters' bounds are summed as they are made, through
    // integer atomics in 2^-20 fixed point (order-independent, so the decision is reproducible; each term rounded up
    // and clipped to 14: "out" stays out) -- two words instead of two arrays: at bufferSize 4096 the arrays' 768 bytes
    // cost the sixth CTA per SM (21.8 -> 27 M frames/s)
    int noise_fx[2];
    float noise_q[2], noise_total, noise_sharp;
    double red_d[kWarps];
    float red_f[kWarps];
    int red_i[kWarps];
    double scan_d[kWarps];
    double red6[8][kWarps];
    double band_sum[MB_MAX_BARK_BANDS];
    float specific[MB_MAX_BARK_BANDS];
    float mel_log[MB_MAX_MEL_FILTERS];
};

__device__ __forceinline__ void noise_fx_add(int *acc, float u) {  // (a NaN or +inf clips to 14 as well)
    atomicAdd(acc, (int)fminf(u * 1048576.f, 1.5e7f) + 1);
}

// noise_sigma > 0: the float32 kernel of an adaptive plan also leaves the mb_adaptive.cuh bounds in `sc`.
template <bool EXACT>
__device__ __forceinline__ void frame_epilogue(const MbDevPlan &P, const mb_outputs &O, int64_t g, MbFrameSums &S,
                                               const MomentAcc &acc, const float *amp, Scratch &sc, float noise_sigma = 0.f) {
    const int M = P.M;
    const uint32_t mask = P.mask;
    const int nb = P.nb, nf = P.nf, nc = P.nc;  // 24 / 26 / 13 unless the plan was created with other parameters
    const int tid = MB_TID, lane = tid & 31, warp = tid >> 5;
    double *scan_d = sc.scan_d, *band_sum = sc.band_sum;
    int *red_i = sc.red_i;
    float *specific = sc.specific, *mel_log = sc.mel_log;
    const bool want_moments =
        mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE));
    const bool want_log = mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS);
    const bool want_bark = mask & (MB_FEATURE_BIT(MB_FEAT_LOUDNESS) | MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SPREAD) |
                                   MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SHARPNESS));
    if (tid == 0) sc.noise_fx[0] = sc.noise_fx[1] = 0;  // (read last by frame_gather, in front of the previous frame's final barrier)
    if (want_moments) {
        // the six sums (and the two floor measures of mb_adaptive.cuh) through ONE pair of barriers (each is reduced
        // exactly as block_sum would: same bits)
        // (mb_warp_sum8: the eight butterflies folded into one -- 9 exchanges instead of 40, same pairings and so the
        // same bits; the straight-line code of these kernels sits at the edge of the instruction cache, every
        // hundred instructions count: profiles/README.md, round 2)
        double v[8] = {acc.s0, acc.s1, acc.s2, acc.s3, acc.s4, acc.lg, (double)acc.q0, (double)acc.q4};
        double r = mb_warp_sum8(v);  // lane L: the warp's total of quantity (L >> 2) & 7
        if constexpr (kWarps > 1) {
            block_sync();
            if ((lane & 3) == 0) sc.red6[lane >> 2][warp] = r;
            block_sync();
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = lane < kWarps ? sc.red6[q][lane] : 0.0;
            r = mb_warp_sum8(v);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) v[q] = __shfl_sync(0xffffffffu, r, 4 * q);
        S.s0 = v[0]; S.s1 = v[1]; S.s2 = v[2]; S.s3 = v[3]; S.s4 = v[4];
        if (want_log) S.log2sum = v[5];
        if (!EXACT && tid == 0) {  // (4: every fourth bin was looked at; 1.6: the exponent trick's worst case, squared)
            sc.noise_q[0] = 6.4f * (float)v[6];
            sc.noise_q[1] = 6.4f * (float)v[7];
        }
    }
    block_sync();

    // ---- rolloff: prefix sums of the amplitude spectrum in double
    if (mb_has(mask, MB_FEAT_SPECTRAL_ROLLOFF)) {
        // warp w owns the contiguous bins [w L, (w+1) L); lanes read them 32 at a time (coalesced, no bank
        // conflicts): first the warp totals, then a running scan that counts the bins m with P[m] <= thr
        const int L = max(32, (M + kWarps - 1) / kWarps);
        const int k_lo = min(M, warp * L), k_hi = min(M, k_lo + L);
        double part = 0;
        for (int k = k_lo + lane; k < k_hi; k += 32) part += (double)amp[k];
        part = mb_warp_sum(part);
        if (lane == 0) scan_d[warp] = part;
        block_sync();
        const double mine = lane < kWarps ? scan_d[lane] : 0.0;
        const double total = mb_warp_sum(mine);
        double run = mb_warp_sum(lane < warp ? mine : 0.0);  // sum of amp[0 .. k_lo)
        const double thr = P.rolloff_frac * total;  // spectralRolloff.js:9 (0.99)
        // bins below this warp's range all count when its first prefix is under the threshold, none of its
        // own count when it is over; only the warp the threshold falls into scans bin by bin
        int cnt = 0;
        const bool starts_under = run <= thr, ends_under = run + part <= thr;
        const int nch = (k_hi - k_lo + 31) >> 5;  // chunks of 32 bins in this warp's range
        if (starts_under && ends_under) cnt = k_hi - k_lo;
        else if (starts_under && nch <= 32) {
            // two levels instead of nch sequential 32-bin scans (the other warps of the CTA wait at the next barrier
            // meanwhile: 18 % of the stall samples of the bufferSize-32768 kernel): lane c sums chunk c (rotated
            // reads: no bank conflicts), one scan over the chunk sums finds the chunk the threshold falls into,
            // and only that chunk is scanned bin by bin
            double csum = 0;
            if (lane < nch) {
                const int b = k_lo + 32 * lane;
#pragma unroll 8
                for (int j = 0; j < 32; j++) {
                    const int k = b + ((j + lane) & 31);
                    if (k < k_hi) csum += (double)amp[k];
                }
            }
            double cincl = csum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, cincl, o);
                if (lane >= o) cincl += y;
            }
            const double cstart = run + (cincl - csum);  // sum of amp[0 .. first bin of chunk `lane`)
            const unsigned under = __ballot_sync(0xffffffffu, lane < nch && cstart <= thr) | 1u;  // (chunk 0: starts_under)
            const int cs = 31 - __clz(under);
            const double crun = __shfl_sync(0xffffffffu, cstart, cs);
            const int k = k_lo + 32 * cs + lane;
            const double x = k < k_hi ? (double)amp[k] : 0.0;
            double incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            cnt = 32 * cs + __popc(__ballot_sync(0xffffffffu, k < k_hi && crun + (incl - x) <= thr));
        } else if (starts_under)
        for (int k0 = k_lo; k0 < k_hi; k0 += 32) {
            const int k = k0 + lane;
            const double x = k < k_hi ? (double)amp[k] : 0.0;
            double incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            cnt += __popc(__ballot_sync(0xffffffffu, k < k_hi && run + (incl - x) <= thr));
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        cnt = block_sum_int(lane == 0 ? cnt : 0, red_i);
        // spectralRolloff.js:11-15: the loop only runs while ec > threshold
        S.rolloff_bin = (total > thr) ? cnt - 1 : M;
    }

    // ---- bark band sums (loudness.js:55-63), one warp per band
    if (want_bark) {
        if constexpr (kWarps <= kLaneBandWarps) {  // small CTAs: one lane per band, ascending like sumArray (loudness.js:69-77)
            if (warp == 0)
                for (int b = lane; b < nb; b += 32) {
                    double s = 0;
                    for (int k = P.bb[b]; k < P.bb[b + 1]; k++) s += (double)amp[k];
                    band_sum[b] = s;
                }
        } else {
            for (int b = warp; b < nb; b += kWarps) {
                double s = 0;
                for (int k = P.bb[b] + lane; k < P.bb[b + 1]; k += 32) s += (double)amp[k];
                s = mb_warp_sum(s);
                if (lane == 0) band_sum[b] = s;
            }
        }
    }
    // ---- mel filterbank energies (mfcc.js:40-65)
    if (mb_has(mask, MB_FEAT_MFCC)) {
        if (EXACT) {  // the reference's order: one float32 running sum per filter
            for (int f = tid; f < nf; f += kThreads) {
                const int e0 = P.mel[f], e1 = P.mel[f + 1], e2 = P.mel[f + 2];
                // weights (i - lo) / (hi - lo) and (hi - i) / (hi - lo) come from the plan (float64 divisions
                // done once, mfcc.js:45-50); the float32 running sum keeps the reference's order (mfcc.js:56-62)
                const double *__restrict__ wgt = P.mel_w_exact + P.mel_w_off[f] - e0;
                (void)e1;
                // s <- float32(s + w p), kept as a float64 that sits on the float32 grid.  The running sum never
                // shrinks, so it stays in the binade of the previous step or moves up: (u + M) - M with the previous
                // step's M = 1.5 * 2^(e + 29) is the float32 rounding of u unless u left that binade (then, or for a
                // NaN, the general sequence below picks the new one).  The loop-carried chain is three DADDs instead
                // of an add and two conversions on the quarter-rate pipe.
                double s = 0.0, Mv = 0.0;
                int eb = -1;  // exponent field M was made for (-1: none yet)
                const int kend = min(e2, M);
                for (int k = e0; k < kend; k++) {
                    const float a = amp[k];
                    const double u = __dadd_rn(s, __dmul_rn(__ldg(wgt + k), (double)__fmul_rn(a, a)));
                    double y = __dsub_rn(__dadd_rn(u, Mv), Mv);  // speculative: issued before the binade check resolves
                    const int eu = max(__double2hiint(u) & 0x7FF00000, 0x38100000);
                    if (eu != eb) {  // rare: at most once per binade the sum climbs through
                        eb = eu;
                        Mv = __hiloint2double(eu + 0x01D80000, 0);
                        y = __dsub_rn(__dadd_rn(u, Mv), Mv);
                    }
                    s = y;  // (u >= +0 or NaN: no zero sign to restore)
                }
                mel_log[f] = (float)log((double)(float)s);
            }
        } else if constexpr (kWarps <= kLaneBandWarps) {  // small CTAs: one lane per filter (on the second warp where there is one)
            if (warp == (kWarps > 1 ? 1 : 0))
            for (int f = lane; f < nf; f += 32) {
                const int e0 = P.mel[f], e1 = P.mel[f + 1], e2 = min(P.mel[f + 2], M);
                const float up = __ldg(P.mel_inv_width + f), dn = __ldg(P.mel_inv_width + f + 1);
                float s = 0.f;
                for (int k = e0; k < e1; k++) {
                    const float a = amp[k];
                    s += (float)(k - e0) * up * (a * a);
                }
                for (int k = e1; k < e2; k++) {
                    const float a = amp[k];
                    s += (float)(e2 - k) * dn * (a * a);
                }
                mel_log[f] = s;  // (the logarithm is taken below, all filters at once)
            }
        } else {  // one warp per filter
            for (int f = warp; f < nf; f += kWarps) {
                const int e0 = P.mel[f], e1 = P.mel[f + 1], e2 = P.mel[f + 2];
                const float up = __ldg(P.mel_inv_width + f), dn = __ldg(P.mel_inv_width + f + 1);
                float s = 0.f;
                for (int k = e0 + lane; k < e1; k += 32) {
                    const float a = amp[k];
                    s += (float)(k - e0) * up * (a * a);
                }
                for (int k = e1 + lane; k < e2; k += 32) {
                    const float a = amp[k];
                    s += (float)(e2 - k) * dn * (a * a);
                }
                s = mb_warp_sum(s);
                if (lane == 0) mel_log[f] = s;
            }
        }
    }
    block_sync();
    // ln of the 26 filter energies by 26 threads at once (mfcc.js:63), not one filter at a time
    if (!EXACT && mb_has(mask, MB_FEAT_MFCC)) {
        for (int f = tid; f < nf; f += kThreads) {
            if (noise_sigma > 0.f) noise_fx_add(&sc.noise_fx[1], mb_noise_mel(mel_log[f], __ldg(&P.noise->mel_c1[f]), __ldg(&P.noise->mel_c2[f]), noise_sigma));
            mel_log[f] = (float)log((double)mel_log[f]);
        }
        block_sync();
    }

    if (want_bark) {
        for (int b = tid; b < nb; b += kThreads) {
            const float sp = (float)pow(band_sum[b], 0.23);
            specific[b] = sp;
            if (!EXACT && noise_sigma > 0.f) noise_fx_add(&sc.noise_fx[0], mb_noise_band((float)band_sum[b], sp, __ldg(&P.noise->band_c[b]), noise_sigma));
            if (mb_has(mask, MB_FEAT_LOUDNESS)) O.loudness_specific[g * nb + b] = sp;
        }
        block_sync();
        if (tid == 0) {
            double total = 0, mx = 0, sharp = 0;
            for (int i = 0; i < nb; i++) {
                const double sp = (double)specific[i];
                total += sp;
                if (sp > mx) mx = sp;
                if (i >= 1 && i <= 15) sharp += (double)i * sp;  // (i+1) * spec[i+1], i < 15
            }
            sharp += P.sharp_const;
            sc.noise_total = (float)total;
            sc.noise_sharp = (float)sharp;
            if (mb_has(mask, MB_FEAT_LOUDNESS)) O.loudness_total[g] = (float)total;
            if (mb_has(mask, MB_FEAT_PERCEPTUAL_SPREAD)) {
                const double r = (total - mx) / total;
                O.perceptual_spread[g] = (float)(r * r);
            }
            if (mb_has(mask, MB_FEAT_PERCEPTUAL_SHARPNESS))
                O.perceptual_sharpness[g] = (float)(sharp * (0.11 / total));
        }
    }
    constexpr int kDctThread0 = kThreads > 32 ? 32 : 0;  // (a second warp where there is one)
    if (mb_has(mask, MB_FEAT_MFCC) && tid >= kDctThread0)
        for (int c = tid - kDctThread0; c < nc; c += kThreads - kDctThread0) {
            double v = 0;
            for (int f = 0; f < nf; f++) v += (double)__ldg(P.dct + c + f * nc) * (double)mel_log[f];
            O.mfcc[g * nc + c] = (float)(v / (double)nc);
        }
}

// The exponent trick behind MomentAcc::cf needs theta sigma < 1; a frame beyond that (samples above ~2^18, or a NaN)
// is simply redone: reported to frame_finish as if it had been rescaled.
__device__ __forceinline__ bool acc_cf_missing(float noise_sigma) { return !(noise_sigma < 1.f / kMbNoiseTheta); }

// The frame's number features and, in the float32 kernel of an adaptive plan, the decision whether the exact-FFT
// kernel has to redo it (mb_adaptive.cuh): ONE thread's work, ~10 float64 divisions and roots.  It is split so that the
// CTA does not wait for it: frame_gather copies what the epilogue left in shared memory into the thread's registers in
// front of the frame's last barrier, frame_finish does the arithmetic and the stores behind it, while the other warps
// are already loading the next frame (the thread's own warp joins them a little later, well before the load's barrier).
__device__ __forceinline__ MbNoiseFrame frame_gather(const MbDevPlan &P, const Scratch &sc, float noise_sigma) {
    MbNoiseFrame NF;
    NF.sigma = noise_sigma;
    NF.q0 = NF.q4 = NF.sum_u = NF.sum_dl = NF.total = NF.sharp = 0.f;
    if (noise_sigma >= 0.f) {
        NF.q0 = sc.noise_q[0];
        NF.q4 = sc.noise_q[1];
        NF.sum_u = (float)sc.noise_fx[0] * (1.0f / 1048576.f);
        NF.sum_dl = (float)sc.noise_fx[1] * (1.0f / 1048576.f);
        NF.total = sc.noise_total;
        NF.sharp = sc.noise_sharp;
    }
    return NF;
}
__device__ __forceinline__ void frame_finish(const MbDevPlan &P, const MbClipTable &T, const mb_outputs &O, int64_t g,
                                             const MbFrameSums &S, const MbNoiseFrame &NF, int kscale) {
    MbMoments MO;
    mb_store_scalars(P, O, g, S, &MO);
    // (a frame rescaled by 2^kscale lies outside the range the bounds were made for: redo it)
    if (NF.sigma >= 0.f && T.fix_count != nullptr && (kscale != 0 || mb_noise_needs_exact(P, P.mask, S, MO, NF)))
        T.fix_list[atomicAdd(T.fix_count, 1)] = (int)g;
}

template <bool EXACT>
__global__ void __launch_bounds__(kThreads)
mb_generic_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                  const float *__restrict__ samples, const __grid_constant__ mb_outputs O) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.N, M = P.M, log2M = P.log2M;
    // fast: work = padded M complex, then amp[M].  exact: re[N], im[N], then amp[M].
    float2 *work = reinterpret_cast<float2 *>(smem_raw);
    float *xre = reinterpret_cast<float *>(smem_raw);
    float *xim = xre + xidx(N) + 1;
    float *amp = EXACT ? (xim + xidx(N) + 1) : reinterpret_cast<float *>(work + pidx(M) + 1);

    __shared__ Scratch sc;
    double *red_d = sc.red_d;
    float *red_f = sc.red_f;
    int *red_i = sc.red_i;

    const uint32_t mask = P.mask;
    const int tid = MB_TID, lane = tid & 31, warp = tid >> 5;
    const bool want_moments =
        mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE));
    const bool want_log = mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS);
    const bool want_time = mask & (MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                                   MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER));
    const uint32_t time_only = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                               MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    const bool want_spectrum = (mask & ~time_only) != 0;
    const bool adapt = !EXACT && want_spectrum && T.fix_count != nullptr;  // this launch lists the frames to be redone exactly

    // (sel_list: only the frames a float32-FFT kernel flagged, mb_adaptive.cuh)
    const int64_t n_work = T.sel_list ? (int64_t)*T.sel_count : T.total_frames;
    for (int64_t it = blockIdx.x; it < n_work; it += gridDim.x) {
        const int64_t g = T.sel_list ? (int64_t)T.sel_list[it] : it;
        const int64_t clip = mb_find_clip_warp(T, g);
        const MbFrameSrc src = mb_frame_src(T, samples, T.clip_off[clip] + (g - T.frame_start[clip]) * (int64_t)P.hop);

        MbFrameSums S;
        S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = S.log2sum = S.energy = 0;
        S.zcr = 0;
        S.rolloff_bin = M;

        // ---- time domain: buffer, energy, zcr; windowed frame into smem
        int kscale = 0;  // fast mode: power-of-two rescale of frames that would under/overflow float32 squares
        {
            double e = 0;
            int z = 0;
            float mxabs = 0.f;
            for (int i = tid; i < M; i += kThreads) {
                const float x0 = src[2 * i], x1 = src[2 * i + 1];
                mxabs = fmaxf(mxabs, fmaxf(fabsf(x0), fabsf(x1)));
                if (adapt && !want_time) e += (double)x0 * (double)x0 + (double)x1 * (double)x1;
                if (want_time) {
                    e += (double)x0 * (double)x0 + (double)x1 * (double)x1;
                    z += ((x0 >= 0.f) != (x1 >= 0.f)) && (x0 == x0) && (x1 == x1);
                    if (2 * i + 2 < N) {
                        const float x2 = src[2 * i + 2];
                        z += ((x1 >= 0.f) != (x2 >= 0.f)) && (x1 == x1) && (x2 == x2);
                    }
                    if (mb_has(mask, MB_FEAT_BUFFER)) {
                        O.buffer[g * N + 2 * i] = x0;
                        O.buffer[g * N + 2 * i + 1] = x1;
                    }
                }
                // computeWindow src/meyda.js:158-168: float32 store of the product
                const float w0 = __fmul_rn(x0, __ldg(P.window + 2 * i));
                const float w1 = __fmul_rn(x1, __ldg(P.window + 2 * i + 1));
                if (EXACT) {  // BitReverseComplexArray lib/jsfft/fft.js:185-208, imag zero
                    const int rshift = 32 - (log2M + 1);
                    const int r0 = (int)(__brev((unsigned)(2 * i)) >> rshift);
                    const int r1 = (int)(__brev((unsigned)(2 * i + 1)) >> rshift);
                    xre[xidx(r0)] = w0; xim[xidx(r0)] = 0.f;
                    xre[xidx(r1)] = w1; xim[xidx(r1)] = 0.f;
                } else {
                    work[pidx(i)] = make_float2(w0, w1);
                }
            }
            if (want_time) {
                S.energy = block_sum(e, red_d);
                S.zcr = block_sum_int(z, red_i);
            } else if (adapt) {
                S.energy = block_sum(e, red_d);
            }
            if (!EXACT && want_spectrum) {
                // the reference squares |Z| in float64; a frame far outside the float32 comfort zone is
                // rescaled by an exact power of two and scaled back on the way out
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, o));
                block_sync();
                if (lane == 0) red_f[warp] = mxabs;
                block_sync();
                float mx = 0.f;
#pragma unroll
                for (int w = 0; w < kWarps; w++) mx = fmaxf(mx, red_f[w]);
                if (mx > 0.f && mx < 3.0e38f && (mx < 0x1p-40f || mx > 0x1p40f)) {
                    int ex;
                    (void)frexpf(mx, &ex);
                    kscale = max(-100, min(100, -ex));
                    const float up = ldexpf(1.f, kscale);
                    for (int i = tid; i < M; i += kThreads) {
                        float2 t = work[pidx(i)];
                        work[pidx(i)] = make_float2(t.x * up, t.y * up);
                    }
                }
            }
        }
        const float unscale = ldexpf(1.f, -kscale);
        block_sync();

        // mb_adaptive.cuh: rms rounding error of one spectrum bin (the frame's own units); < 0: not an adaptive launch
        const float noise_sigma = adapt ? mb_noise_sigma((float)S.energy, 1.0f / (float)N) : -1.f;
        if (want_spectrum) {
            MomentAcc acc;
            if (adapt && noise_sigma > 0.f && noise_sigma < 1.f / kMbNoiseTheta)
                acc.cf = 0x7EF311C7 + __float_as_int(kMbNoiseTheta * noise_sigma) - 0x3F800000;
            if (EXACT) {
                // ---- FFT_2_Iterative lib/jsfft/fft.js:139-168: doubles, no FMA, f32 stage stores
                exact_stages(xre, xim, P.tw_exact, N, 0, log2M + 1);
                if (mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM)) {
                    for (int k = tid; k < N; k += kThreads) {
                        O.complex_real[g * N + k] = xre[xidx(k)];
                        O.complex_imag[g * N + k] = xim[xidx(k)];
                    }
                }
                for (int k = tid; k < M; k += kThreads) {  // computeAmplitude src/meyda.js:104-114
                    const double r = (double)xre[xidx(k)], i = (double)xim[xidx(k)];
                    const float av = (float)sqrt(__dadd_rn(__dmul_rn(r, r), __dmul_rn(i, i)));
                    amp[k] = av;
                    if (mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM)) O.amplitude_spectrum[g * M + k] = av;
                    if (mb_has(mask, MB_FEAT_POWER_SPECTRUM)) O.power_spectrum[g * M + k] = __fmul_rn(av, av);
                    if (want_moments) acc.add(av, k, want_log);
                }
            } else {
                // ---- in-place DIF FFT, output in bit-reversed positions
                int log2s = log2M;
                while (log2s > 0) {
                    const int q = log2s >= 3 ? 3 : log2s;
                    if (q == 3) fft_pass<3>(work, P.twM, M, log2M, log2s);
                    else if (q == 2) fft_pass<2>(work, P.twM, M, log2M, log2s);
                    else fft_pass<1>(work, P.twM, M, log2M, log2s);
                    log2s -= q;
                    block_sync();
                }
                // ---- real-FFT split, spectra out, amplitude into smem, moment partials
                const float sc = P.inv_sqrt_N;
                const int rshift = 32 - log2M;
                for (int k = tid; k < M; k += kThreads) {
                    const int kk = (M - k) & (M - 1);
                    const float2 a = work[pidx((int)(__brev((unsigned)k) >> rshift))];
                    const float2 b = work[pidx((int)(__brev((unsigned)kk) >> rshift))];
                    const float2 Z = split_bin(a, b, __ldg(&P.twN[k]), sc);
                    float zr = Z.x, zi = Z.y;
                    if (mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM)) {
                        float *re = O.complex_real + g * N, *im = O.complex_imag + g * N;
                        const float zro = zr * unscale, zio = zi * unscale;
                        re[k] = zro;
                        im[k] = zio;
                        if (k > 0) {
                            re[N - k] = zro;
                            im[N - k] = -zio;
                        } else {
                            re[M] = (a.x - a.y) * sc * unscale;  // Nyquist bin: (E[0] - O[0]) / sqrt(N)
                            im[M] = (a.x - a.y) * 0.f + 0.f;  // +0, or NaN when the frame holds a NaN
                        }
                    }
                    const float av = sqrt_fast(zr * zr + zi * zi) * unscale;
                    amp[k] = av;
                    if (mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM)) O.amplitude_spectrum[g * M + k] = av;
                    if (mb_has(mask, MB_FEAT_POWER_SPECTRUM)) O.power_spectrum[g * M + k] = __fmul_rn(av, av);
                    if (want_moments) acc.add<true>(av, k, want_log);
                }
            }
            frame_epilogue<EXACT>(P, O, g, S, acc, amp, sc, adapt ? noise_sigma : 0.f);
        }
        if (adapt) block_sync();  // (the epilogue's bounds are read by the one thread below)
        MbNoiseFrame NF;
        if (tid == kScalarThread) NF = frame_gather(P, sc, adapt ? noise_sigma : -1.f);
        block_sync();  // smem reused by the next frame
        if (tid == kScalarThread) frame_finish(P, T, O, g, S, NF, (adapt && acc_cf_missing(noise_sigma)) ? 1 : kscale);
    }
}


#if MB_GENERIC_THREADS == 256 || MB_GENERIC_THREADS == 1024  // the two CTA sizes the cluster kernel is launched with
// ---- exact-FFT mode for a frame that does not fit one CTA ---------------------------------------
// At bufferSize 32768 the reference's N-point complex transform needs 256 KB as float32 re/im -- more
// than one CTA's shared memory.  A 2-CTA thread-block cluster holds it: CTA r keeps positions
// [r N/2, (r+1) N/2) of the bit-reversed array.  Every radix-2 stage but the last pairs elements inside
// one half; the last stage (width N/2) pairs element j of CTA 0 with element j of CTA 1, which each CTA
// reads from its peer through distributed shared memory (cluster.map_shared_rank), after a
// cluster-wide barrier.  Amplitudes are gathered into CTA 0, which runs the feature epilogue.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads)
mb_exact_cluster_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                        const float *__restrict__ samples, const __grid_constant__ mb_outputs O) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int N = P.N, M = P.M, H = N / 2, log2N = P.log2M + 1;
    float *xre = reinterpret_cast<float *>(smem_raw), *xim = xre + xidx(H) + 1, *amp = xim + xidx(H) + 1;  // amp[M]: CTA 0
    __shared__ Scratch sc;
    const float *re0 = cluster.map_shared_rank(xre, 0), *im0 = cluster.map_shared_rank(xim, 0);
    const float *re1 = cluster.map_shared_rank(xre, 1), *im1 = cluster.map_shared_rank(xim, 1);
    float *amp0 = cluster.map_shared_rank(amp, 0);

    const uint32_t mask = P.mask;
    const int tid = MB_TID;
    const bool want_moments =
        mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE));
    const bool want_log = mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS);
    const bool want_time = mask & (MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                                   MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER));
    const uint32_t time_only = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                               MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    const bool want_spectrum = (mask & ~time_only) != 0;
    const double SQRT1_2 = 0.70710678118654752440;
    const int64_t n_clusters = gridDim.x / 2, cid = blockIdx.x / 2;

    const int64_t n_work = T.sel_list ? (int64_t)*T.sel_count : T.total_frames;
    for (int64_t it = cid; it < n_work; it += n_clusters) {
        const int64_t g = T.sel_list ? (int64_t)T.sel_list[it] : it;
        const int64_t clip = mb_find_clip_warp(T, g);
        const MbFrameSrc src = mb_frame_src(T, samples, T.clip_off[clip] + (g - T.frame_start[clip]) * (int64_t)P.hop);
        MbFrameSums S;
        S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = S.log2sum = S.energy = 0;
        S.zcr = 0;
        S.rolloff_bin = M;

        if (rank == 0 && want_time) {  // time-domain features over the whole raw frame
            double e = 0;
            int z = 0;
            for (int i = tid; i < N; i += kThreads) {
                const float x0 = src[i];
                e += (double)x0 * (double)x0;
                if (i + 1 < N) {
                    const float x1 = src[i + 1];
                    z += ((x0 >= 0.f) != (x1 >= 0.f)) && (x0 == x0) && (x1 == x1);
                }
                if (mb_has(mask, MB_FEAT_BUFFER)) O.buffer[g * N + i] = x0;
            }
            S.energy = block_sum(e, sc.red_d);
            S.zcr = block_sum_int(z, sc.red_i);
        }
        if (want_spectrum) {
            // this CTA's half of BitReverseComplexArray(windowed frame), imag zero
            const int rshift = 32 - log2N;
            for (int pl = tid; pl < H; pl += kThreads) {
                const int i = (int)(__brev((unsigned)(rank * H + pl)) >> rshift);
                xre[xidx(pl)] = __fmul_rn(src[i], __ldg(P.window + i));
                xim[xidx(pl)] = 0.f;
            }
            block_sync();
            exact_stages(xre, xim, P.tw_exact, H, 0, log2N - 1);  // widths 1 .. N/4: inside the half
            cluster.sync();  // both halves are complete and visible cluster-wide
            {                // width N/2: element j of CTA 0 with element j of CTA 1, through DSMEM
                const double2 *__restrict__ tw = P.tw_exact + (H - 1);
                for (int jj = tid; jj < H / 2; jj += kThreads) {
                    const int j = (int)rank * (H / 2) + jj;
                    const double2 f = __ldg(&tw[j]);
                    const double lr = (double)re0[xidx(j)], li = (double)im0[xidx(j)];
                    const double xr = (double)re1[xidx(j)], xi = (double)im1[xidx(j)];
                    const double rr = __dsub_rn(__dmul_rn(f.x, xr), __dmul_rn(f.y, xi));
                    const double ri = __dadd_rn(__dmul_rn(f.y, xr), __dmul_rn(f.x, xi));
                    const float zr = (float)__dmul_rn(SQRT1_2, __dadd_rn(lr, rr));
                    const float zi = (float)__dmul_rn(SQRT1_2, __dadd_rn(li, ri));
                    if (mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM)) {
                        O.complex_real[g * N + j] = zr;
                        O.complex_imag[g * N + j] = zi;
                        O.complex_real[g * N + j + H] = (float)__dmul_rn(SQRT1_2, __dsub_rn(lr, rr));
                        O.complex_imag[g * N + j + H] = (float)__dmul_rn(SQRT1_2, __dsub_rn(li, ri));
                    }
                    const float av = (float)sqrt(__dadd_rn(__dmul_rn((double)zr, (double)zr), __dmul_rn((double)zi, (double)zi)));
                    amp0[j] = av;  // gathered on CTA 0
                    if (mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM)) O.amplitude_spectrum[g * M + j] = av;
                    if (mb_has(mask, MB_FEAT_POWER_SPECTRUM)) O.power_spectrum[g * M + j] = __fmul_rn(av, av);
                }
            }
            cluster.sync();  // CTA 0 holds all N/2 amplitudes
            if (rank == 0) {
                MomentAcc acc;
                if (want_moments)
                    for (int k = tid; k < M; k += kThreads) acc.add(amp[k], k, want_log);
                frame_epilogue<true>(P, O, g, S, acc, amp, sc);
            }
        }
        if (rank == 0 && tid == kScalarThread) mb_store_scalars(P, O, g, S);
        cluster.sync();  // CTA 0 is done with the gathered amplitudes before the next frame overwrites them
    }
}

#endif

#if MB_GENERIC_THREADS == 64 || MB_GENERIC_THREADS == 128 || MB_GENERIC_THREADS == 256 || MB_GENERIC_THREADS == 512
// ---- bufferSize 4096 / 8192 / 16384 / 32768, float32 FFT: one CTA of R = 2 / 4 / 8 / 16 warps per frame ----------
// (written for R = 16, bufferSize 32768 = BASELINE config 5, and described for it; the same code compiled in the
// 64- / 128- / 256-thread instantiations of this header serves the three sizes below it with 8 / 4 / 2 CTAs per SM)
// The packed frame is M = 16384 complex points = 16 x 1024.  Decimation in time by 16: warp r
// transforms z_r[m] = z[16 m + r] with the same 32 x 32 register FFT as the bufferSize-2048 kernel
// (two register FFTs and one transpose through the warp's slot), multiplies X_r[k] by
// exp(+2 pi i r k / M), and a radix-16 register FFT across the sixteen warps' results gives
// C[k + 1024 q].  Shared memory (213 KB): one 136 KB area that is in turn the padded windowed frame,
// the sixteen transpose slots, the twiddled sub-spectra and the natural-order spectrum; 64 KB of
// amplitudes for the band features; 12 KB of twiddle tables.
constexpr int kBigR = kWarps, kBigLogR = kBigR == 16 ? 4 : kBigR == 8 ? 3 : kBigR == 4 ? 2 : 1, kBigK = 32 / kBigR;
constexpr int kBigN = 2048 * kBigR, kBigM = 1024 * kBigR, kBigRow = 33, kBigSlot = 32 * kBigRow;  // slot: 1056 float2
struct BigSmem {
    float2 area[kBigM + kBigM / 16];  // 17408 float2: z padded one per 16; >= 16 slots of 1056; >= C[16384]
    float amp[kBigM];
    float2 tw32[32 * 32];             // exp(+2 pi i b c / 1024) at [c*32 + b]
    float2 tw16[kBigR * 32];          // exp(+2 pi i r d / (32 R)) at [r*32 + d]  (R = 16: / 512)
    float2 tw_step[32];               // exp(+2 pi i j / 64): the split twiddle of bin tid + kThreads j over that of bin tid
    // a frame whose number features and refine decision are still to be made (see "deferred finish" in the kernel)
    MbFrameSums fin_S;
    MbNoiseFrame fin_NF;
    int64_t fin_g;
    int fin_kscale;
};

__global__ void __launch_bounds__(kThreads, 512 / kThreads)
mb_big32768_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                   const float *__restrict__ samples, const __grid_constant__ mb_outputs O) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BigSmem &B = *reinterpret_cast<BigSmem *>(smem_raw);
    __shared__ Scratch sc;
    const int N = kBigN, M = kBigM;
    const uint32_t mask = P.mask;
    const int tid = MB_TID, lane = tid & 31, warp = tid >> 5;
    const bool want_moments =
        mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE));
    const bool want_log = mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS);
    const bool want_time = mask & (MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                                   MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER));
    const uint32_t time_only = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                               MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    const bool want_spectrum = (mask & ~time_only) != 0;
    const bool want_cs = mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM);
    const bool adapt = want_spectrum && T.fix_count != nullptr;  // this launch lists the frames to be redone exactly

    // tables (once per persistent CTA) and this thread's own twiddle exp(+2 pi i warp lane / M)
    for (int i = tid; i < 32 * 32; i += kThreads) {
        const double ang = 2.0 * 3.14159265358979323846 * (double)((i >> 5) * (i & 31)) / 1024.0;
        B.tw32[i] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (int i = tid; i < kBigR * 32; i += kThreads) {
        const double ang = 2.0 * 3.14159265358979323846 * (double)((i >> 5) * (i & 31)) / (32.0 * kBigR);
        B.tw16[i] = make_float2((float)cos(ang), (float)sin(ang));
    }
    if (tid < 32) {
        const double ang = 2.0 * 3.14159265358979323846 * (double)tid / 64.0;
        B.tw_step[tid] = make_float2((float)cos(ang), (float)sin(ang));
    }
    float2 tw_own, tw_split;  // tw_split: exp(+2 pi i tid / N), the real-FFT split twiddle of this thread's first bin
    {
        const double ang = 2.0 * 3.14159265358979323846 * (double)(warp * lane) / (double)M;
        tw_own = make_float2((float)cos(ang), (float)sin(ang));
        const double ang2 = 2.0 * 3.14159265358979323846 * (double)tid / (double)N;
        tw_split = make_float2((float)cos(ang2), (float)sin(ang2));
    }
    block_sync();
    float2 *slot = B.area + warp * kBigSlot;  // this warp's transpose slot / sub-spectrum row

    // Deferred finish.  Turning a frame's sums into its number features and deciding whether the exact kernel must
    // redo it is ~2,000 clocks of ONE thread (float64 roots, powers and divisions); at the end of the frame the other
    // warps wait for it at the next barrier.  With 8 or 16 warps per frame the thread's warp instead does it while
    // the OTHER warps load the next frame -- a phase bound by two L2 round trips, not by issue slots, so that the
    // loaders take over its share for free.  Only when no time-domain output is asked for: those sums (energy, zcr)
    // are per-thread partials whose order -- and so whose last bit -- would otherwise depend on whether a frame is a
    // CTA's first.  (BASELINE config 5: 5.8 -> 6.1 M frames/s.)
    constexpr int kFinWarp = kScalarThread >> 5;
    const bool defer = kBigR >= 8 && !want_time;
    bool pending = false;

    int64_t clip_next = -1;  // the clip of this CTA's next frame, found while its samples are prefetched
    for (int64_t g = blockIdx.x; g < T.total_frames; g += gridDim.x) {
        const int64_t clip = clip_next >= 0 ? clip_next : mb_find_clip_warp(T, g);
        const float *__restrict__ src = samples + T.clip_off[clip] + (g - T.frame_start[clip]) * (int64_t)P.hop;
        MbFrameSums S;
        S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = S.log2sum = S.energy = 0;
        S.zcr = 0;
        S.rolloff_bin = M;

        // ---- 1. the frame, windowed, into the padded area (coalesced 16-byte loads); time-domain sums
        int kscale = 0;
        {
            double e = 0;
            float ef = 0.f;
            int z = 0;
            float mxabs = 0.f;
            const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(src);
            const float4 *__restrict__ win4 = reinterpret_cast<const float4 *>(P.window);
            // a frame off the 16-byte grid (odd clip offsets, hop not a multiple of 4) is read sample by sample:
            // same values, same arithmetic, so a frame's bits do not depend on where it lies
            const bool src_aligned = (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            auto take = [&](const int i, const float4 x) {
                const float4 w = __ldg(win4 + i);
                mxabs = fmaxf(mxabs, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
                if (adapt && !want_time) ef = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, ef))));  // (sigma needs 1 % of it)
                if (want_time) {
                    e += (double)(x.x * x.x + x.y * x.y) + (double)(x.z * x.z + x.w * x.w);
                    const float nx = (4 * i + 4 < N) ? __ldg(src + 4 * i + 4) : x.w;  // last sample has no successor
                    const bool p0 = x.x >= 0.f, p1 = x.y >= 0.f, p2 = x.z >= 0.f, p3 = x.w >= 0.f, p4 = nx >= 0.f;
                    const bool n0 = x.x == x.x, n1 = x.y == x.y, n2 = x.z == x.z, n3 = x.w == x.w, n4 = nx == nx;
                    z += ((p0 != p1) && n0 && n1) + ((p1 != p2) && n1 && n2) + ((p2 != p3) && n2 && n3) +
                         ((p3 != p4) && n3 && n4);
                    if (mb_has(mask, MB_FEAT_BUFFER)) __stcs(reinterpret_cast<float4 *>(O.buffer + g * N) + i, x);
                }
                const int m0 = 2 * i, m1 = 2 * i + 1;
                B.area[m0 + (m0 >> 4)] = mbx2::mul(make_float2(x.x, x.y), make_float2(w.x, w.y));
                B.area[m1 + (m1 >> 4)] = mbx2::mul(make_float2(x.z, x.w), make_float2(w.z, w.w));
            };
            const int i0 = pending ? tid - (warp > kFinWarp ? 32 : 0) : tid, di = pending ? kThreads - 32 : kThreads;
            if (pending && warp == kFinWarp) {  // (deferred finish of the previous frame: this warp loads nothing)
                if (lane == 0) frame_finish(P, T, O, B.fin_g, B.fin_S, B.fin_NF, B.fin_kscale);
            } else if (src_aligned) {
#pragma unroll 4
                for (int i = i0; i < N / 4; i += di) take(i, __ldg(src4 + i));
            } else {
#pragma unroll 2
                for (int i = i0; i < N / 4; i += di)
                    take(i, make_float4(__ldg(src + 4 * i), __ldg(src + 4 * i + 1), __ldg(src + 4 * i + 2), __ldg(src + 4 * i + 3)));
            }
            pending = false;
            if (want_time) {
                S.energy = block_sum(e, sc.red_d);
                S.zcr = block_sum_int(z, sc.red_i);
            }
            if (want_spectrum) {  // frames outside the float32 comfort zone: exact power-of-two rescale (see generic kernel)
                const bool need_e = adapt && !want_time;  // (the frame's energy for mb_adaptive.cuh rides on the same exchange)
                if (need_e) e = (double)mb_warp_sum(ef);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mxabs = fmaxf(mxabs, __shfl_xor_sync(0xffffffffu, mxabs, o));
                block_sync();
                if (lane == 0) {
                    sc.red_f[warp] = mxabs;
                    if (need_e) sc.red_d[warp] = e;
                }
                block_sync();
                if (need_e) S.energy = mb_warp_sum(lane < kWarps ? sc.red_d[lane] : 0.0);
                float mx = lane < kWarps ? sc.red_f[lane] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (mx > 0.f && mx < 3.0e38f && (mx < 0x1p-40f || mx > 0x1p40f)) {
                    int ex;
                    (void)frexpf(mx, &ex);
                    kscale = max(-100, min(100, -ex));
                    const float up = ldexpf(1.f, kscale);
                    for (int m = tid; m < M; m += kThreads) {
                        float2 t = B.area[m + (m >> 4)];
                        B.area[m + (m >> 4)] = make_float2(t.x * up, t.y * up);
                    }
                }
            }
        }
        const float unscale = ldexpf(1.f, -kscale);
        const float noise_sigma = adapt ? mb_noise_sigma((float)S.energy, 1.0f / (float)N) : -1.f;  // mb_adaptive.cuh
        {   // this CTA's next frame on its way into L2 while the current one is transformed (its 128 KB were
            // otherwise fetched with four exposed round trips at the top of the next iteration)
            const int64_t gn = g + gridDim.x;
            clip_next = -1;
            if (gn < T.total_frames) {
                clip_next = mb_find_clip_warp(T, gn);
                const float *nsrc = samples + T.clip_off[clip_next] + (gn - T.frame_start[clip_next]) * (int64_t)P.hop;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + 32 * tid));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nsrc + 32 * (tid + kThreads)));
            }
        }
        block_sync();

        if (want_spectrum) {
            // ---- 2. warp r takes z[16 m + r], m = 32 a + lane
            float2 v[32];
            // z[R m + r], m = 32 a + lane, behind its padding of one per 16: (32 R a + R lane + r) >> 4 = 2 R a + const
            const int z0 = kBigR * lane + warp, zb = z0 + (z0 >> 4);  // (R = 16: 544 a + 17 lane + r)
#pragma unroll
            for (int a = 0; a < 32; a++) v[a] = B.area[34 * kBigR * a + zb];
            block_sync();  // everybody has its samples: the area becomes the sixteen warp slots
            // ---- 3. the 1024-point sub-FFT of this warp: 32 x 32 in registers, one transpose through the slot
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                if (pass == 1) {
#pragma unroll
                    for (int b = 0; b < 32; b++) v[b] = slot[lane * kBigRow + b];
                }
                mbfft::fft_reg<32>(v);
                if (pass == 0) {
#pragma unroll
                    for (int c = 0; c < 32; c++) {
                        float2 y = v[mbfft::brev<5>(c)];
                        if (c > 0) y = cmul(y, B.tw32[c * 32 + lane]);
                        slot[c * kBigRow + lane] = y;
                    }
                }
                __syncwarp();
            }
            // ---- 4. X_r[k] * exp(+2 pi i r k / M), k = lane + 32 d, into the warp's row in natural order
#pragma unroll
            for (int d = 0; d < 32; d++) {
                float2 y = v[mbfft::brev<5>(d)];
                if (warp > 0) y = cmul(y, cmul(tw_own, B.tw16[warp * 32 + d]));
                slot[lane + 32 * d] = y;
            }
            block_sync();
            // ---- 5. radix-R across the warps' rows: C[k + 1024 q] for k = tid + kThreads kk (R = 16: k = tid, tid + 512)
            float2 u[kBigK][kBigR];
#pragma unroll
            for (int r = 0; r < kBigR; r++) {
#pragma unroll
                for (int kk = 0; kk < kBigK; kk++) u[kk][r] = B.area[r * kBigSlot + tid + kThreads * kk];
            }
#pragma unroll
            for (int kk = 0; kk < kBigK; kk++) mbfft::fft_reg<kBigR>(u[kk]);
            block_sync();  // all rows consumed: the area becomes the natural-order spectrum C[0 .. M)
#pragma unroll
            for (int q = 0; q < kBigR; q++) {
#pragma unroll
                for (int kk = 0; kk < kBigK; kk++) B.area[tid + kThreads * kk + 1024 * q] = u[kk][mbfft::brev<kBigLogR>(q)];
            }
            block_sync();
            // ---- 6. real-FFT split, spectra out, amplitude into smem, moment partials
            MomentAcc acc;
            if (adapt && noise_sigma > 0.f && noise_sigma < 1.f / kMbNoiseTheta)
                acc.cf = 0x7EF311C7 + __float_as_int(kMbNoiseTheta * noise_sigma) - 0x3F800000;
            const float sc_n = P.inv_sqrt_N;
            // (the split twiddle exp(+2 pi i k / N) of bin k = tid + kThreads j is this thread's own one times a
            // broadcast table entry, exp(+2 pi i j / 64): the 128 KB global table cost an exposed L2 round trip per
            // four bins here -- 11 % of the kernel's stall samples, profiles/r02_ncu_big32768_source.txt)
            static_assert(kBigM / kThreads == 32, "32 bins per thread");
#pragma unroll 4
            for (int jj = 0; jj < 32; jj++) {
                const int k = tid + kThreads * jj;
                const float2 w = cmul(tw_split, B.tw_step[jj]);
                const float2 a = B.area[k];
                const float2 b = B.area[(M - k) & (M - 1)];
                const float2 Z = split_bin(a, b, w, sc_n);
                const float zr = Z.x, zi = Z.y;
                if (want_cs) {
                    float *re = O.complex_real + g * N, *im = O.complex_imag + g * N;
                    const float zro = zr * unscale, zio = zi * unscale;
                    __stcs(re + k, zro);
                    __stcs(im + k, zio);
                    if (k > 0) {
                        __stcs(re + (N - k), zro);
                        __stcs(im + (N - k), -zio);
                    } else {
                        __stcs(re + M, (a.x - a.y) * sc_n * unscale);  // Nyquist bin
                        __stcs(im + M, (a.x - a.y) * 0.f + 0.f);
                    }
                }
                const float av = sqrt_fast(zr * zr + zi * zi) * unscale;
                B.amp[k] = av;
                if (mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM)) __stcs(O.amplitude_spectrum + g * M + k, av);
                if (mb_has(mask, MB_FEAT_POWER_SPECTRUM)) __stcs(O.power_spectrum + g * M + k, __fmul_rn(av, av));
                if (want_moments) acc.add<true>(av, k, want_log);
            }
            frame_epilogue<false>(P, O, g, S, acc, B.amp, sc, adapt ? noise_sigma : 0.f);
        }
        if (adapt) block_sync();
        MbNoiseFrame NF;
        if (tid == kScalarThread) NF = frame_gather(P, sc, adapt ? noise_sigma : -1.f);
        const int fin_k = (adapt && acc_cf_missing(noise_sigma)) ? 1 : kscale;
        if (defer) {  // parked; made by this thread's warp while the others load the next frame
            if (tid == kScalarThread) {
                B.fin_S = S;
                B.fin_NF = NF;
                B.fin_g = g;
                B.fin_kscale = fin_k;
            }
            pending = true;
        }
        block_sync();  // smem reused by the next frame
        if (!defer && tid == kScalarThread) frame_finish(P, T, O, g, S, NF, fin_k);
    }
    if (pending && tid == kScalarThread) frame_finish(P, T, O, B.fin_g, B.fin_S, B.fin_NF, B.fin_kscale);
}
#endif  // MB_GENERIC_THREADS == 64 .. 512
