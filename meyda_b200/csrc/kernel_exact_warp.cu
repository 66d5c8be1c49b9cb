// kernel_exact_warp.cu -- the reference's own FFT arithmetic (MB_FLAG_EXACT_FFT, and the second pass of the adaptive
// float32 plans: mb_adaptive.cuh), one WARP per frame, bufferSize 512 / 1024 / 2048.
//
// What is reproduced (lib/jsfft/fft.js:123-171): an N-point radix-2 decimation-in-time transform of the windowed frame
// (imaginary parts zero) after the bit-reversal swap (:185-208); every butterfly in float64 WITHOUT fused multiply-adds
// (JavaScript has none), twiddles from the reference's own recurrence f <- f * del (:162-164, computed on the host in
// the same order), both outputs scaled by SQRT1_2, and a float32 rounding at every stage store (`output` is a
// Float32Array).  complexSpectrum / amplitudeSpectrum / powerSpectrum come out bit for bit.
//
// How it is laid out for the B200 (FP64 runs at half the FP32 rate here: 64 lanes per clock and SM):
//   * a frame is 2 N floats (re, im) in the warp's own shared-memory buffer; no block barrier anywhere;
//   * the log2 N stages are fused into three register passes of 4 + Q1 + Q2 stages (2048: 4 + 4 + 3): a lane holds
//     the 16 (or 8) points of a sub-transform as float64 pairs, so a frame crosses shared memory three times, not 11;
//   * INSIDE a pass the float32 store is emulated on the float64 registers: y = (x + M) - M with
//     M = 1.5 * 2^(e + 29), e = max(exponent(x), -126) -- two DADDs whose round-to-nearest-even lands on the float32
//     grid of x's binade (on the 2^-149 grid below 2^-126, which is the float32 subnormal rounding), the sign of a
//     zero result copied from x.  The block-per-frame kernel converted to float32 and back at every stage: F2F runs
//     on the quarter-rate XU pipe and was that kernel's limiter (42 % of its issue slots busy on XU, 684 M F2F per
//     launch); here F2F only appears where a pass meets shared memory;
//   * pass 1 reads the raw frame straight from global memory in bit-reversed order (lane r takes samples
//     r + (N/16) rev4(j)), its twiddles are the 15 entries of widths 1..8, warp-uniform, read from a __grid_constant__
//     parameter (constant-bank operands); the later passes read theirs from a shared-memory copy of the table;
//   * the buffer is swizzled, position p lives at p ^ ((p >> (log2 N - 4)) & 15), which makes pass 1's scattered
//     stores and every later load / store free of bank conflicts;
//   * the epilogue is the generic kernels' exact one (kernel_generic_impl.cuh compiled for a team of one warp): f64
//     moment sums, the reference's sequential float32 mel running sum (mfcc.js:56-62), f64 Bark sums, rolloff scan.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mb_adaptive.cuh"
#include "mb_device.cuh"
#include "mb_kernels.h"
#include "mb_fft.cuh"

namespace gw {  // the generic epilogue for a team of one warp
#define MB_GENERIC_THREADS 32
#define MB_GENERIC_WARP_LOCAL 1
#include "kernel_generic_impl.cuh"
#undef MB_GENERIC_THREADS
}  // namespace gw

namespace {

// warps (= frames in flight) per CTA: at 2048, 11 x 16 KB of frames + the 32 KB twiddle table + the epilogue scratch fill
// an SM's shared memory; the smaller sizes are bounded by the register file instead
template <int LOG2N>
__host__ __device__ constexpr int warps_per_cta() { return LOG2N == 11 ? 11 : 12; }

struct SmallTw {
    double2 f[15];  // lib/jsfft/fft.js:143-164 for widths 1, 2, 4, 8: entry (w - 1) + j
};

// Float32Array store emulated in float64 registers (see the header).  Exact for every finite x below the float32
// overflow threshold, zeros and subnormals included.
__device__ __forceinline__ double round_f32(double x) {
    const int hi = __double2hiint(x);
    const int e = max(hi & 0x7FF00000, 0x38100000);                 // exponent field, at least that of 2^-126
    const double M = __hiloint2double(e + 0x01D80000, 0);           // 1.5 * 2^(e + 29)
    const double y = __dsub_rn(__dadd_rn(x, M), M);
    return __hiloint2double(__double2hiint(y) | (hi & (int)0x80000000), __double2loint(y));
}

// The same store as a conversion pair (two instructions on the quarter-rate XU pipe instead of seven on the FP64 and
// integer pipes): used for half of the values so that all three pipes carry the roundings.
__device__ __forceinline__ double round_f32_cvt(double x) { return (double)__double2float_rn(x); }

// fft.js:151-161 on float64 registers holding float32 values; results NOT yet rounded.
__device__ __forceinline__ void bfly(double &lr, double &li, double &xr, double &xi, const double fr, const double fi) {
    const double SQRT1_2 = 0.70710678118654752440;
    const double rr = __dsub_rn(__dmul_rn(fr, xr), __dmul_rn(fi, xi));
    const double ri = __dadd_rn(__dmul_rn(fi, xr), __dmul_rn(fr, xi));
    const double a = __dmul_rn(SQRT1_2, __dadd_rn(lr, rr)), b = __dmul_rn(SQRT1_2, __dadd_rn(li, ri));
    const double c = __dmul_rn(SQRT1_2, __dsub_rn(lr, rr)), d = __dmul_rn(SQRT1_2, __dsub_rn(li, ri));
    lr = a; li = b; xr = c; xi = d;
}

template <int LOG2N>
__device__ __forceinline__ int swz(int p) { return p ^ ((p >> (LOG2N - 4)) & 15); }

__host__ __device__ constexpr int rev_bits(int v, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) r |= ((v >> b) & 1) << (bits - 1 - b);
    return r;
}

// Q fused stages of widths W, 2W, .. on the groups {p = (h << (LOG2W + Q)) | (t << LOG2W) | j, t < 2^Q}; lane takes
// group lane + 32 i.  The pass ends in shared memory: its last float32 store is the conversion itself.
template <int LOG2N, int LOG2W, int Q>
__device__ __forceinline__ void pass_smem(float2 *X, const double2 *tw, int lane) {
    constexpr int N = 1 << LOG2N, R = 1 << Q, W = 1 << LOG2W, groups = N >> Q;
#pragma unroll 1
    for (int gi = lane; gi < groups; gi += 32) {
        const int j = gi & (W - 1), base = ((gi >> LOG2W) << (LOG2W + Q)) | j;
        double vr[R], vi[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            const float2 x = X[swz<LOG2N>(base + (t << LOG2W))];
            vr[t] = (double)x.x;
            vi[t] = (double)x.y;
        }
#pragma unroll
        for (int q = 0; q < Q; q++) {
            const int h = 1 << q;  // partner distance in units of W; this stage has width h W
#pragma unroll
            for (int t = 0; t < R; t++) {
                if (t & h) continue;
                const double2 f = tw[(h << LOG2W) - 16 + ((t & (h - 1)) << LOG2W) + j];  // entry (hW - 1) + index, table starts at 15
                bfly(vr[t], vi[t], vr[t + h], vi[t + h], f.x, f.y);
                if (q + 1 < Q) {
                    vr[t] = round_f32(vr[t]); vi[t] = round_f32_cvt(vi[t]);
                    vr[t + h] = round_f32_cvt(vr[t + h]); vi[t + h] = round_f32(vi[t + h]);
                }
            }
        }
#pragma unroll
        for (int t = 0; t < R; t++)
            X[swz<LOG2N>(base + (t << LOG2W))] = make_float2(__double2float_rn(vr[t]), __double2float_rn(vi[t]));
    }
}

template <int LOG2N>
__global__ void __launch_bounds__(warps_per_cta<LOG2N>() * 32, 1)
mb_exact_warp_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                     const float *__restrict__ samples, const __grid_constant__ mb_outputs O,
                     const __grid_constant__ SmallTw TW) {
    constexpr int N = 1 << LOG2N, M = N / 2, kWarpsPerCta = warps_per_cta<LOG2N>();
    constexpr int R1 = LOG2N - 4, Q1 = (R1 + 1) / 2, Q2 = R1 - Q1;  // stages after the first four: two passes
    static_assert(LOG2N >= 9 && LOG2N <= 11 && Q1 <= 4 && Q2 >= 1, "bufferSize 512, 1024 or 2048");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2 *tw = reinterpret_cast<double2 *>(smem_raw);                               // [N - 16]: widths 16 .. N/2
    float2 *Xall = reinterpret_cast<float2 *>(smem_raw + sizeof(double2) * (N - 16));  // [warps][N]
    __shared__ gw::Scratch sc_all[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 *X = Xall + warp * N;
    float *amp = reinterpret_cast<float *>(X) + N;  // N/2 floats over the upper half of the buffer (bins >= N/2), once those are stored
    gw::Scratch &sc = sc_all[warp];
    // (the second pass of an adaptive plan usually finds nothing to do: leave before the 32 KB table is fetched)
    if ((T.sel_list ? (int64_t)*T.sel_count : T.total_frames) <= (int64_t)blockIdx.x * kWarpsPerCta) return;
    for (int i = threadIdx.x; i < N - 16; i += kWarpsPerCta * 32) tw[i] = P.tw_exact[15 + i];
    __syncthreads();
    // Two groups of warps, each in step with itself only (named barriers 1 and 2): while one group sits in a
    // latency-bound phase (the sequential mel sums, the frame loads) the other is in a register pass, and the SM's
    // instruction cache still only has to hold two phases.  The second group starts half a round late.
    constexpr int kGroupA = (kWarpsPerCta + 1) / 2;
    const int grp = warp < kGroupA ? 0 : 1;
    const int grp_threads = 32 * (grp == 0 ? kGroupA : kWarpsPerCta - kGroupA);
    auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(grp_threads) : "memory"); };
    if (grp == 1) __nanosleep(30000);

    const uint32_t mask = P.mask;
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

    // (sel_list: only the frames a float32-FFT kernel flagged, mb_adaptive.cuh)
    const int64_t n_work = T.sel_list ? (int64_t)*T.sel_count : T.total_frames;
    // The warps of a CTA take one frame each per round and move through the phases of a round TOGETHER (a block
    // barrier between phases, nothing shared but the instruction stream): a phase's unrolled register code is
    // 10-25 KB, and with every warp in a phase of its own the SM's instruction cache thrashed (54 % of the stall
    // samples were instruction fetches, issue slots 19 % busy).
    const int64_t r0 = (int64_t)blockIdx.x * kWarpsPerCta, rstride = (int64_t)gridDim.x * kWarpsPerCta;
    for (int64_t round = r0; round < n_work; round += rstride) {
        const int64_t it = round + warp;
        const bool active = it < n_work;
        const int64_t g = !active ? 0 : T.sel_list ? (int64_t)T.sel_list[it] : it;
        const int64_t clip = mb_find_clip_warp(T, g);
        const MbFrameSrc src = mb_frame_src(T, samples, T.clip_off[clip] + (g - T.frame_start[clip]) * (int64_t)P.hop);
        MbFrameSums S;
        S.s0 = S.s1 = S.s2 = S.s3 = S.s4 = S.log2sum = S.energy = 0;
        S.zcr = 0;
        S.rolloff_bin = M;

        if (want_time && active) {  // energy.js, zcr.js, buffer over the raw frame
            double e = 0;
            int z = 0;
            const bool vec = src.format < 0 && ((reinterpret_cast<uintptr_t>(src.base) | reinterpret_cast<uintptr_t>(O.buffer)) & 15) == 0;
            if (vec) {  // float32 frames on the 16-byte grid: four samples per lane and load
                const float4 *s4 = reinterpret_cast<const float4 *>(src.base);
#pragma unroll 4
                for (int i = lane; i < N / 4; i += 32) {
                    const float4 x = __ldg(s4 + i);
                    float nx = __shfl_down_sync(0xffffffffu, x.x, 1);  // the sample after this lane's four
                    if (lane == 31) nx = (i + 1 < N / 4) ? __ldg(reinterpret_cast<const float *>(src.base) + 4 * i + 4) : x.w;
                    e += ((double)x.x * (double)x.x + (double)x.y * (double)x.y) + ((double)x.z * (double)x.z + (double)x.w * (double)x.w);
                    const bool p0 = x.x >= 0.f, p1 = x.y >= 0.f, p2 = x.z >= 0.f, p3 = x.w >= 0.f, p4 = nx >= 0.f;
                    const bool n0 = x.x == x.x, n1 = x.y == x.y, n2 = x.z == x.z, n3 = x.w == x.w, n4 = nx == nx;
                    z += ((p0 != p1) && n0 && n1) + ((p1 != p2) && n1 && n2) + ((p2 != p3) && n2 && n3) +
                         ((p3 != p4) && n3 && n4 && (4 * i + 4 < N));  // (the frame's last sample has no successor)
                    if (mb_has(mask, MB_FEAT_BUFFER)) __stcs(reinterpret_cast<float4 *>(O.buffer + g * N) + i, x);
                }
            } else {
#pragma unroll 4
                for (int i = lane; i < N; i += 32) {
                    const float x0 = src[i];
                    e += (double)x0 * (double)x0;
                    if (i + 1 < N) {
                        const float x1 = src[i + 1];
                        z += ((x0 >= 0.f) != (x1 >= 0.f)) && (x0 == x0) && (x1 == x1);
                    }
                    if (mb_has(mask, MB_FEAT_BUFFER)) O.buffer[g * N + i] = x0;
                }
            }
            S.energy = mb_warp_sum(e);
            S.zcr = mb_warp_sum(z);
        }
        if (want_spectrum) {
            // (the next group's samples are fetched while this one is transformed: with the warps of the CTA in step, a
            // load every lane waits for is a stall of the whole SM)
            float xin[16];
#pragma unroll
            for (int j = 0; j < 16; j++) xin[j] = active ? src[lane + (N / 16) * rev_bits(j, 4)] : 0.f;  // (in flight across the barrier)
            group_sync();  // (also: the previous round's epilogue is done with the buffer)
            if (active) {
            // ---- pass 1: widths 1, 2, 4, 8 on BitReverseComplexArray(windowed frame): group g16 holds positions
            // 16 g16 + j = samples rev(g16) + (N/16) rev4(j); this lane takes r = rev(g16) = lane + 32 i
#pragma unroll 1
            for (int r = lane; r < N / 16; r += 32) {
                double vr[16], vi[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int idx = r + (N / 16) * rev_bits(j, 4);
                    vr[j] = (double)__fmul_rn(xin[j], __ldg(P.window + idx));  // computeWindow src/meyda.js:158-168
                    vi[j] = 0.0;
                }
                if (r + 32 < N / 16) {
#pragma unroll
                    for (int j = 0; j < 16; j++) xin[j] = src[r + 32 + (N / 16) * rev_bits(j, 4)];
                }
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int h = 1 << q;
#pragma unroll
                    for (int t = 0; t < 16; t++) {
                        if (t & h) continue;
                        const double2 f = TW.f[(h - 1) + (t & (h - 1))];
                        bfly(vr[t], vi[t], vr[t + h], vi[t + h], f.x, f.y);
                        if (q < 3) {
                            vr[t] = round_f32(vr[t]); vi[t] = round_f32_cvt(vi[t]);
                            vr[t + h] = round_f32_cvt(vr[t + h]); vi[t + h] = round_f32(vi[t + h]);
                        }
                    }
                }
                const int g16 = (int)(__brev((unsigned)r) >> (32 - (LOG2N - 4)));
#pragma unroll
                for (int j = 0; j < 16; j++)
                    X[swz<LOG2N>(16 * g16 + j)] = make_float2(__double2float_rn(vr[j]), __double2float_rn(vi[j]));
            }
            }
            group_sync();
            // ---- pass 2: widths 16 .. 16 * 2^(Q1 - 1)
            if (active) pass_smem<LOG2N, 4, Q1>(X, tw, lane);
            group_sync();
            // ---- pass 3: the last Q2 stages
            if (active) pass_smem<LOG2N, LOG2N - Q2, Q2>(X, tw, lane);
            group_sync();
            gw::MomentAcc acc;
            if (active) {
            // ---- spectra out; amplitudes (computeAmplitude src/meyda.js:104-114: bins below N/2) into the upper half of
            // the buffer (the bins above N/2 are only ever needed for complexSpectrum, stored right here)
            if (want_cs) {
#pragma unroll 4
                for (int k = M + lane; k < N; k += 32) {
                    const float2 z = X[swz<LOG2N>(k)];
                    __stcs(O.complex_real + g * N + k, z.x);
                    __stcs(O.complex_imag + g * N + k, z.y);
                }
            }
            __syncwarp();  // the upper half is out: its floats become the amplitude array
#pragma unroll 4
            for (int k = lane; k < M; k += 32) {
                const float2 z = X[swz<LOG2N>(k)];
                if (want_cs) {
                    __stcs(O.complex_real + g * N + k, z.x);
                    __stcs(O.complex_imag + g * N + k, z.y);
                }
                const double re = (double)z.x, im = (double)z.y;
                const float a = (float)sqrt(__dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im)));
                amp[k] = a;
                if (mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM)) __stcs(O.amplitude_spectrum + g * M + k, a);
                if (mb_has(mask, MB_FEAT_POWER_SPECTRUM)) __stcs(O.power_spectrum + g * M + k, __fmul_rn(a, a));
                if (want_moments) acc.add(a, k, want_log);
            }
            __syncwarp();
            }
            group_sync();
            if (active) gw::frame_epilogue<true>(P, O, g, S, acc, amp, sc);
        }
        if (active && lane == 0) mb_store_scalars(P, O, g, S);
    }
}

template <int LOG2N>
cudaError_t launch_one(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O, int num_sms,
                       cudaStream_t stream, const SmallTw &TW) {
    constexpr int kWarpsPerCta = warps_per_cta<LOG2N>();
    const size_t smem = mb_exact_warp_smem_bytes(1 << LOG2N);
    cudaError_t e = cudaFuncSetAttribute(mb_exact_warp_kernel<LOG2N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mb_exact_warp_kernel<LOG2N>, kWarpsPerCta * 32, smem);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms * per_sm;
    const int64_t need = (T.total_frames + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > need) grid = need;
    if (grid < 1) return cudaSuccess;
    (void)cudaGetLastError();
    mb_exact_warp_kernel<LOG2N><<<(unsigned)grid, kWarpsPerCta * 32, smem, stream>>>(P, T, samples, O, TW);
    return cudaGetLastError();
}

}  // namespace

size_t mb_exact_warp_smem_bytes(int N) {
    const int warps = N == 2048 ? warps_per_cta<11>() : warps_per_cta<10>();
    return sizeof(double2) * (size_t)(N - 16) + (size_t)warps * N * sizeof(float2);
}

bool mb_exact_warp_supports(int N) { return N == 512 || N == 1024 || N == 2048; }

cudaError_t mb_launch_exact_warp(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                                 int num_sms, cudaStream_t stream, const double *tw_small /* [30]: entries 0..14, (re, im) */) {
    SmallTw TW;
    for (int i = 0; i < 15; i++) TW.f[i] = make_double2(tw_small[2 * i], tw_small[2 * i + 1]);
    if (P.N == 2048) return launch_one<11>(P, T, samples, O, num_sms, stream, TW);
    if (P.N == 1024) return launch_one<10>(P, T, samples, O, num_sms, stream, TW);
    return launch_one<9>(P, T, samples, O, num_sms, stream, TW);
}
