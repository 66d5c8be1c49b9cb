// mb_adaptive.cuh -- which frames of a float32-FFT kernel must be redone with the reference's own FFT arithmetic.
//
// Why: x^0.23 (loudness.js:60), ln x (mfcc.js:63) and the k^3 / k^4 weighted sums of src/utils.js:1-11 amplify the
// spectrum's rounding-noise floor without bound.  Where a Bark band, a mel filter or the upper bins of a tonal frame
// hold nothing but FFT rounding noise, the REFERENCE's value is set by its own float32 per-stage rounding
// (lib/jsfft/fft.js:158-161) and only a bit-identical FFT lands within 1e-3 of it.  The float32 kernels therefore
// bound, per frame, how far each requested feature can move under spectral noise of the size that separates a float32
// FFT from the reference's (measured: rms |Z_fast - Z_ref| = 1.5e-7 sqrt(E_windowed / N), both noises together), and
// frames whose bound exceeds half the parity tolerance (1e-3 relative or absolute, BASELINE.json) are appended to a
// list that the exact-FFT kernel then works through, overwriting those frames' spectral outputs.
//
// The noise model (tools/flag_calibrate.py checks it against the reference arithmetic on tonal, noisy, low-passed and
// quantised signals: no violating frame is missed, ordinary audio is not flagged):
//   sigma   = 1e-7 sqrt(E_raw / N)                      rms error of one spectrum bin (E_raw >= E_windowed)
//   a bin   moves by ~sigma where a <~ sigma (its value IS the noise) and by sigma^2 / a above: bias ~ sigma q(a),
//           q(a) = min(1, (32 sigma / a)^2);  Q_p = sum_k q(a_k) k^p (accumulated per 32-bin block, k := block end)
//   a sum   of n bins moves by K sigma sqrt(n) (random signs, K = 16 covers the spikes a tone leaves at its
//           radix-2 aliases k0 + N / 2^s) plus the bias of its floor bins
// Everything here is plan-time constants and a few dozen instructions per frame; the per-bin part is three
// instructions in the kernels' blocked loop (MUFU.RCP, FMNMX, FFMA).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "mb_device.cuh"

constexpr float kMbNoiseRel = 1e-7f;    // sigma / sqrt(E_raw / N)
constexpr float kMbNoiseTheta = 32.f;   // q(a) = min(1, (theta sigma / a)^2)
constexpr float kMbNoiseK = 16.f;       // random-sum safety factor (band / filter sums)
constexpr double kMbNoiseKap = 8.0;     // random-sum safety factor (moment sums over all n bins)
constexpr float kMbNoiseHalfTol = 5e-4f;  // half of the 1e-3 parity tolerance

// sigma of a frame from its raw energy (the caller passes the energy in the units its amplitudes are in)
// (these are bounds with safety factors of 2 .. 32 behind them: one MUFU root, 2^-22 relative, subnormals flushed,
// instead of the ten-instruction IEEE sequence with its slow path)
__device__ __forceinline__ float mb_noise_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float mb_noise_sigma(float energy, float inv_N) { return kMbNoiseRel * mb_noise_sqrt(energy * inv_N); }

// One Bark band (loudness.js:55-63): how far specific[b] = (sum a)^0.23 can move when the sum moves by e = cB sigma,
// cB = 2 n_b + K sqrt(n_b) (plan constant, 0 for an empty band): 0.23 sp (e / sum)(1 + e / sum) x 2 while e < sum / 2
// (the x^0.23 curve's own bound there is 0.23 (e/sum) / (1 - e/sum)^0.77 <= 0.39 (e/sum)), "out" (+inf) beyond, or when it
// alone breaks the tolerance, or on a NaN.  Branch-free: one reciprocal, a handful of multiplies and selects.
__device__ __forceinline__ float mb_noise_band(float bsum, float sp, float cB, float sigma) {
    const float r = __fdividef(cB * sigma, bsum);             // (0 / 0 = NaN for an empty or silent band: nothing to move)
    const float u = 0.46f * sp * r * (1.f + r);
    const bool none = !(cB * sigma > 0.f);
    const bool ok = (r < 0.5f) && (u <= kMbNoiseHalfTol * fmaxf(1.f, sp));
    return none ? 0.f : (ok ? u : INFINITY);
}

// One mel filter (mfcc.js:53-65): how far ln E_f can move.  c1 = 2 K sqrt(W_f / max(W_f, 1)), c2 = 4 W_f with
// W_f the filter's total weight (0: the filter is empty, -inf in the reference too).  Branch-free.
__device__ __forceinline__ float mb_noise_mel(float E, float c1, float c2, float sigma) {
    const float r2 = sigma * __fdividef(sigma, E);  // (E == 0: inf)
    const float d = fmaf(c1, mb_noise_sqrt(r2), c2 * r2);
    const bool none = !(c2 * sigma > 0.f);
    return none ? 0.f : ((d < 0.5f) ? 2.f * d : INFINITY);
}

struct MbNoiseFrame {
    float sigma;    // mb_noise_sigma of the frame (its own units)
    float q0, q4;   // sum q(a_k), sum q(a_k) k_blockend^4
    float sum_u;    // sum of mb_noise_band over the bands (inf: one band alone is out)
    float sum_dl;   // sum of mb_noise_mel over the filters
    float total, sharp;  // loudness.total and the weighted sum of perceptualSharpness.js:6-8 + its constant
};

// The frame-level decision, one frame per thread.  `mask`: requested features; M: what mb_store_scalars derived from
// the sums; P.noise_sqrtT[p] = kap sqrt(sum_k k^(2p)).  First-order error propagation with the exact partial
// derivatives of spectralSkewness.js / spectralKurtosis.js (a handful of multiplications and three divisions).
__device__ __forceinline__ bool mb_noise_needs_exact(const MbDevPlan &P, uint32_t mask, const MbFrameSums &S, const MbMoments &M,
                                                     const MbNoiseFrame &F) {
    if (!(S.energy > 0.0)) return S.energy != 0.0;  // silence: nothing to amplify; NaN: the exact path owns the special values
    const float tol = kMbNoiseHalfTol;
    bool bad = false;
    const uint32_t bark = MB_FEATURE_BIT(MB_FEAT_LOUDNESS) | MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SPREAD) | MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SHARPNESS);
    if (mask & bark) {
        const float su = F.sum_u, tot = F.total;
        bad |= !(su <= tol * fmaxf(1.f, tot));                              // loudness.total (and every specific[b], see mb_noise_band)
        bad |= !(4.f * su <= tol * tot);                                    // perceptualSpread: d(r^2) <= 2 r (max u + sum u) / total
        const float sh = __fdividef(0.11f * F.sharp, tot);
        bad |= !((1.65f + sh) * su <= tol * fmaxf(1.f, sh) * tot);          // perceptualSharpness: weights <= 15 x 0.11
    }
    if (mb_has(mask, MB_FEAT_MFCC)) bad |= !(0.02134f * F.sum_dl <= tol);   // max |dct| / 13 = sqrt(2/26) / 13
    const uint32_t mom = MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                         MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                         MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE);
    if (mask & mom) {
        // float32 throughout (MUFU reciprocals and roots): these are bounds with safety factors of 2 .. 16 behind them,
        // and in the block-per-frame kernels ONE thread evaluates them while the CTA waits.  Where float32 runs out
        // (var cancelling below 1e-7 m1^2) the comparisons fail on a NaN or a negative and the frame is redone.
        const float sg = F.sigma, n = (float)P.M, q0 = F.q0, q4 = F.q4;
        // Q_p <= Q_0^(1 - p/4) Q_4^(p/4) (moments are log-convex in p); sums move by sigma (kap sqrt(T_2p) + Q_p)
        const float g = (q0 > 0.f && q4 > 0.f) ? mb_noise_sqrt(mb_noise_sqrt(__fdividef(q4, q0))) : 0.f;
        const float inv0 = sg * __fdividef(1.f, (float)S.s0);  // sigma / S_0
        const float r0 = ((float)P.noise_sqrtT[0] + q0) * inv0;  // relative motion of S_0
        const float q1 = q0 * g, q2 = q1 * g, q3 = q2 * g;
        const float m1 = (float)M.m1, m2 = (float)M.m2, m3 = (float)M.m3, m4 = (float)M.m4, var = (float)M.var, sd = (float)M.sd;
        // absolute motion of m_i = S_i / S_0: (dS_i + m_i dS_0) / S_0
        const float d1 = ((float)P.noise_sqrtT[1] + q1) * inv0 + m1 * r0;
        const float d2 = ((float)P.noise_sqrtT[2] + q2) * inv0 + m2 * r0;
        const float d3 = ((float)P.noise_sqrtT[3] + q3) * inv0 + m3 * r0;
        const float d4 = ((float)P.noise_sqrtT[4] + q4) * inv0 + m4 * r0;
        if (mb_has(mask, MB_FEAT_SPECTRAL_CENTROID)) bad |= !(d1 <= tol * fmaxf(1.f, m1));
        // spectralSlope.js:17 is linear in the centroid, alpha (c - (n-1)/2), and compared relatively
        if (mb_has(mask, MB_FEAT_SPECTRAL_SLOPE)) bad |= !(d1 <= tol * fabsf(m1 - 0.5f * (n - 1.f)));
        const uint32_t high = MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS);
        if (mask & high) {
            const float dv = d2 + 2.f * m1 * d1;  // motion of var = m2 - m1^2
            bad |= !(dv <= 0.25f * var);          // (first order only holds while the variance keeps its size)
            const float iv = __fdividef(1.f, var), isd = __fdividef(1.f, sd);
            if (mb_has(mask, MB_FEAT_SPECTRAL_SPREAD)) bad |= !(0.5f * dv * isd <= tol * fmaxf(1.f, sd));
            if (mb_has(mask, MB_FEAT_SPECTRAL_SKEWNESS)) {  // A / var^1.5, A = 2 m1^3 - 3 m1 m2 + m3
                const float A = (float)(2.0 * M.m1 * M.m1 * M.m1 - 3.0 * M.m1 * M.m2 + M.m3), c = 1.5f * A * iv;
                const float e = (fabsf(6.f * m1 * m1 - 3.f * m2 + c * 2.f * m1) * d1 + fabsf(-3.f * m1 - c) * d2 + d3) * iv * isd;
                bad |= !(e <= tol * fmaxf(1.f, fabsf(A * iv * isd)));
            }
            if (mb_has(mask, MB_FEAT_SPECTRAL_KURTOSIS)) {  // B / var^2, B = -3 m1^4 + 6 m1 m2 - 4 m1 m3 + m4
                const float B = (float)(-3.0 * M.m1 * M.m1 * M.m1 * M.m1 + 6.0 * M.m1 * M.m2 - 4.0 * M.m1 * M.m3 + M.m4), c = 2.f * B * iv;
                const float e = (fabsf(-12.f * m1 * m1 * m1 + 6.f * m2 - 4.f * m3 + c * 2.f * m1) * d1 + fabsf(6.f * m1 - c) * d2 +
                                 4.f * m1 * d3 + d4) * iv * iv;
                bad |= !(e <= tol * fmaxf(1.f, fabsf(B * iv * iv)));
            }
        }
        if (mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS)) {
            // flatness = exp(mean ln a) n / S0: mean ln a moves by (Q0 + 4 sqrt(Q0) / theta) / n, S0 by r0
            const float dml = __fdividef(q0 + 4.f * mb_noise_sqrt(q0) * (1.f / kMbNoiseTheta), n) + r0;
            bad |= !(dml + dml * dml <= tol);  // RELATIVE motion of the flatness (its magnitude runs from 1e-5 on tones to 1); e^x - 1 <= x + x^2
            bad |= (S.log2sum < -1e30) && (S.s0 > 0.0);  // a bin that is exactly 0 here need not be in the reference
        }
    }
    return bad;
}

// Sums four floats over the warp with five shuffle steps (the first two fold the four values onto lane bits 4 and 3):
// lane l returns the total of value (l >> 3) & 3.
__device__ __forceinline__ float mb_warp_sum4(float a0, float a1, float a2, float a3, int lane) {
    const bool h16 = lane & 16, h8 = lane & 8;
    const float b0 = (h16 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, h16 ? a0 : a2, 16);  // a0 | a2
    const float b1 = (h16 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, h16 ? a1 : a3, 16);  // a1 | a3
    float c = (h8 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, h8 ? b0 : b1, 8);             // a0, a1 | a2, a3
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

// Sums two floats over the warp with five shuffle steps: lanes < 16 return the total of a, the others that of b.
__device__ __forceinline__ float mb_warp_sum2(float a, float b, int lane) {
    const bool h = lane & 16;
    float c = (h ? b : a) + __shfl_xor_sync(0xffffffffu, h ? a : b, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    return c;
}

// A warp appends its flagged frames (need: one frame per lane) to the plan's list.
__device__ __forceinline__ void mb_noise_append(const MbClipTable &T, bool need, int64_t g) {
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (m == 0u || T.fix_count == nullptr) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == 0) base = atomicAdd(T.fix_count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (need) T.fix_list[base + __popc(m & ((1u << lane) - 1u))] = (int)g;
}
