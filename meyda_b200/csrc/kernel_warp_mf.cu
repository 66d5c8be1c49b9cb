// kernel_warp_mf.cu -- warp kernel for bufferSize 256, 512 and 1024 (the reference's own default sizes), float32 FFT.
//
// Same machinery as kernel_warp.cu (bufferSize 2048), with F = 2048 / N frames transformed by a warp at a time:
// the warp's 1024 complex points are F packed frames of M = N/2 = 32 A points each (A = 16, 8 or 4).
//   pass A  lane b holds z_f[32 a + b], a < A, for every frame f: F register FFTs of A points per lane;
//   pass B  after a twiddle and a transpose through the warp's slot, lane f A + p holds row p of frame f and
//           runs ONE 32-point register FFT -> X_f[p + A q], q < 32 (the 2048 kernel's pass 2, unchanged);
//   split   the F spectra lie back to back in the slot (frame stride M + 8 float2 against bank conflicts) and
//           are read as 32 rows of 32 consecutive bins, A rows per frame: from here on "bin c of the warp" is
//           bin c mod M of frame c / M, and the stores, the blocked amplitude layout and the branch-free band
//           pieces are those of the 2048 kernel, with sums that stop at frame boundaries (shuffles inside
//           groups of A lanes) and a band / filter / coefficient finish that walks F frames.
// Frames of a group never exchange data, so a frame's bits do not depend on its neighbours (streaming ==
// batch); a short last group repeats its last frame and masks the stores.
//
// Reference path being replaced: src/meyda.js:69-91,104-114,158-168, lib/jsfft/fft.js:123-208 and the
// extractor files under src/extractors/.
#include <cstddef>
#include <utility>

#include "mb_adaptive.cuh"
#include "mb_device.cuh"
#include "mb_fft.cuh"
#include "mb_kernels.h"
#include "mb_warp_common.cuh"

namespace {

using namespace mbwarp;

#ifdef MB_NO_NOISE_STATS  // A/B builds only (tools/build_variants.sh): the kernels without the mb_adaptive.cuh statistics
constexpr bool kNoise = false;
#else
constexpr bool kNoise = true;
#endif

constexpr int kP = 32;
constexpr int kPts = kP * kP;  // complex points per warp per group: F frames x M
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
#ifndef MB_MF_LOCK_WARPS
#define MB_MF_LOCK_WARPS 4
#endif
constexpr int kLockWarps = MB_MF_LOCK_WARPS;  // warps that step through a group's phases together (see the kernel)
constexpr int kRow = kP + 1;
constexpr int kSlotFloats = 2176;  // per warp: >= the 32 x 33 float2 transpose, F (M + 8) float2 spectra, amplitudes + pieces
constexpr int kAmpStride = 36;
constexpr int kPieceOff = 1152;
constexpr int kMaxPieces = MB_MF_MAX_PIECES;
constexpr int kStashRows = 22;  // (18 .. 21: mb_adaptive.cuh statistics)
// pieces follow the blocked amplitudes: {sum a, sum p, sum w p} as a float4 (one 16-byte store per flush), or packed
// as three floats at bufferSize 256, where eight frames need 336 of them
static_assert(kSlotFloats >= 2 * kP * kRow && kPieceOff + 4 * 256 <= kSlotFloats && kPieceOff + 3 * kMaxPieces <= kSlotFloats,
              "slot layout");

struct Smem {
    float2 tw32[kP * kP];  // [c'][b]: exp(+2 pi i p b / M), p = c' mod A
    float2 twN[kPts];      // h exp(+2 pi i k / N), k = c mod M (F copies)
    float window[2 * kPts];  // F copies of the N-point window
    float dct[MB_NUM_MFCC * MB_NUM_MEL_FILTERS];
    float mel_inv[MB_NUM_MEL_FILTERS + 2];
    int mel_edge[MB_NUM_MEL_FILTERS + 2];
    short piece_edge[kMaxPieces];
    float noise_c[3][32];  // mb_adaptive.cuh: band_c[b], mel_c1[f], mel_c2[f]
    float sig[kWarps][8];  // rms rounding error of one spectrum bin, per frame of the warp's current group
    short seg_ptr[MB_MF_MAX_SEGMENTS + 1];
    unsigned short seg_items[MB_MF_MAX_ITEMS];
    unsigned long long bar[kWarps];
    float stash[kWarps][kStashRows][kChunk];
    alignas(128) float slot[kWarps][kSlotFloats];
};
static_assert(offsetof(Smem, slot) % 128 == 0 && (kSlotFloats * 4) % 16 == 0, "warp slots must stay 16-byte aligned");

__host__ __device__ constexpr int brev5(int k) { return mbfft::brev<5>(k); }
// sums inside aligned groups of G lanes
template <int G>
__device__ __forceinline__ double group_sum_d(double v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum_f(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// FFT of the A points v[BASE .. BASE + A): X[k] left in v[BASE + brev(k)]
template <int A, int BASE>
__device__ __forceinline__ void fft_at(float2 (&v)[32]) {
    float2 t[A];
#pragma unroll
    for (int i = 0; i < A; i++) t[i] = v[BASE + i];
    mbfft::fft_reg<A>(t);
#pragma unroll
    for (int i = 0; i < A; i++) v[BASE + i] = t[i];
}
template <int A, int... F>
__device__ __forceinline__ void fft_frames(float2 (&v)[32], std::integer_sequence<int, F...>) {
    (fft_at<A, F * A>(v), ...);
}

// kA: rows of 32 bins per frame (bufferSize = 64 kA).  kPcm: 16-bit PCM input converted in the pass-A load
// (x = s / 32768, exact), as in kernel_warp.cu: bit-identical to the float path on the converted samples.
// kMask: compile-time feature set (0 = the plan's, at run time), as in kernel_warp.cu.
template <int kA, bool kPcm, uint32_t kMask>
__global__ void __launch_bounds__(kThreads, 1)
mb_warpmf_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                 const float *__restrict__ samples, const __grid_constant__ mb_outputs O, const int64_t total_chunks) {
    constexpr int kF = 32 / kA;        // frames per group
    constexpr int kM = 32 * kA;        // complex points = amplitude bins per frame
    constexpr int kN = 2 * kM;         // bufferSize
    constexpr int kMs = kM + 8;        // float2 stride between the frames' spectra in the slot
    constexpr int kABits = kA == 16 ? 4 : kA == 8 ? 3 : 2;
    constexpr int kPf = kA == 4 ? 3 : 4;  // floats per piece
    static_assert(kA == 16 || kA == 8 || kA == 4, "bufferSize 1024, 512 or 256");
    static_assert(2 * (kF * kMs) <= kSlotFloats, "spectra must fit the slot");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    if (smem_u32(smem_raw) & 127u) __trap();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = kMask ? kMask : P.mask;
    const MbWarpMfTables *__restrict__ WT = P.warp_mf_tables;

    for (int i = tid; i < kP * kP; i += kThreads) S.tw32[i] = WT->tw32[i];
    {
        const float h = 0.5f * P.inv_sqrt_N;
        for (int i = tid; i < kPts; i += kThreads) S.twN[i] = make_float2(P.twN[i % kM].x * h, P.twN[i % kM].y * h);
    }
    for (int i = tid; i < 2 * kPts; i += kThreads) S.window[i] = P.window[i % kN];
    for (int i = tid; i < MB_NUM_MFCC * MB_NUM_MEL_FILTERS; i += kThreads) S.dct[i] = P.dct[i];
    if (tid < MB_NUM_MEL_FILTERS + 2) {
        S.mel_edge[tid] = P.mel[tid];
        S.mel_inv[tid] = tid < MB_NUM_MEL_FILTERS + 1 ? P.mel_inv_width[tid] : 0.f;
    }
    for (int i = tid; i < kMaxPieces; i += kThreads) S.piece_edge[i] = WT->piece_edge[i];
    if (tid < 32) {
        S.noise_c[0][tid] = tid < MB_NUM_BARK_BANDS ? P.noise->band_c[tid] : 0.f;
        S.noise_c[1][tid] = tid < MB_NUM_MEL_FILTERS ? P.noise->mel_c1[tid] : 0.f;
        S.noise_c[2][tid] = tid < MB_NUM_MEL_FILTERS ? P.noise->mel_c2[tid] : 0.f;
    }
    for (int i = tid; i <= MB_MF_MAX_SEGMENTS; i += kThreads) S.seg_ptr[i] = WT->seg_ptr[i];
    for (int i = tid; i < MB_MF_MAX_ITEMS; i += kThreads) S.seg_items[i] = WT->seg_items[i];
    if (lane == 0) mbar_init(&S.bar[warp], 1);
    __syncthreads();

    const uint32_t bmask = WT->lane_bmask[lane];
    const int slot_base = WT->lane_slot_base[lane];

    const bool want_buffer = mb_has(mask, MB_FEAT_BUFFER);
    const bool want_time = (mask & (MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) | MB_FEATURE_BIT(MB_FEAT_ZCR)));
    const bool want_cs = mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM);
    const bool want_amp_out = mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM);
    const bool want_pow_out = mb_has(mask, MB_FEAT_POWER_SPECTRUM);
    const bool want_moments =
        (mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                 MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                 MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE)));
    // sum k^2 a .. sum k^4 a are only needed by spread / skewness / kurtosis (centroid, flatness and slope stop at k a)
    // (decided at compile time only: the run-time-mask instantiation keeps all five sums, a branch there cost 3 %)
    constexpr bool want_high = kMask == 0 || (kMask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) |
                                                       MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS))) != 0;
    const bool want_log = mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS);
    const bool want_rolloff = mb_has(mask, MB_FEAT_SPECTRAL_ROLLOFF);
    const bool want_bark = (mask & (MB_FEATURE_BIT(MB_FEAT_LOUDNESS) | MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SPREAD) |
                                    MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SHARPNESS)));
    const bool want_mfcc = mb_has(mask, MB_FEAT_MFCC);
    const bool want_pieces = want_bark || want_mfcc;
    const bool want_blocked = want_rolloff || want_pieces || want_moments;
    const uint32_t time_only = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                               MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    const bool want_spectrum = (mask & ~time_only) != 0;

    float *slot = S.slot[warp];
    float2 *slot2 = reinterpret_cast<float2 *>(slot);
    float(*stash)[kChunk] = S.stash[warp];
    unsigned long long *bar = &S.bar[warp];
    uint32_t parity = 0;
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();
    const int lf = lane / kA, lp = lane % kA;  // this lane's frame and row in pass B / the blocked layout

    int64_t clip = 0, clip_f0 = 0, clip_f1 = 0;
    if (T.n_clips > 0) clip_f1 = T.frame_start[1];

    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kWarps;

    // Lock-step phases.  A group's code is ~8000 straight-line instructions (~125 KB): sixteen warps each somewhere
    // else in it cycle through more than the SM's instruction cache holds, and the kernel then starves on instruction
    // fetch (ncu: 28 % of the stall samples "no instruction" once the adaptive statistics had added 6 % of code;
    // profiles/README.md, round 2).  Warps therefore march in groups of kLockWarps -- one warp per scheduler -- that
    // meet at a named barrier after every phase (load + time domain, pass A, pass B, split, blocked sums; the band
    // loops share the next group's first phase), so that at most 16 / kLockWarps phases are resident at a time.  The
    // groups drift against each other as before and keep hiding each other's memory and barrier waits.  Trip counts
    // are uniform per CTA; a warp without work just keeps the barrier count.
    const int n_phase_bars = 1 + (want_spectrum ? 3 + (want_blocked ? 1 : 0) : 0);
    // (a feature set without band features -- no pieces, no Bark / mel loops: the config-1 instantiation -- is a third of
    // the code and fits as it is: there the barriers only cost)
    constexpr bool kLock = kMask == 0 || (kMask & (MB_FEATURE_BIT(MB_FEAT_LOUDNESS) | MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SPREAD) |
                                                   MB_FEATURE_BIT(MB_FEAT_PERCEPTUAL_SHARPNESS) | MB_FEATURE_BIT(MB_FEAT_MFCC))) != 0;
    auto phase_sync = [&]() {
        __syncwarp();
        if constexpr (kLock) {
            if (kMask != 0 || want_pieces)  // (run-time mask: the same rule, decided per launch)
                asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / kLockWarps), "n"(32 * kLockWarps) : "memory");
        }
    };
    (void)warp_global;
    for (int64_t base = (int64_t)blockIdx.x * kWarps; base < total_chunks; base += warp_stride) {
        const int64_t ch = base + warp;
        const int64_t g0 = ch * kChunk;
        const int nfc = ch < total_chunks ? (int)min((int64_t)kChunk, T.total_frames - g0) : 0;

        for (int j = 0; j < kChunk; j += kF) {
            if (j >= nfc) {
                for (int b = 0; b < n_phase_bars; b++) phase_sync();
                continue;
            }
            const int64_t g = g0 + j;             // first frame of the group
            const int nfg = min(kF, nfc - j);     // valid frames in it
            // ---- where the frames start (a short group repeats its last frame)
            const float *src[kF];        // float32 input
            const int16_t *src16[kF];    // 16-bit PCM input
            bool all_aligned = true;
#pragma unroll
            for (int f = 0; f < kF; f++) {
                const int64_t gf = g + min(f, nfg - 1);
                if (gf >= clip_f1 || gf < clip_f0) {
                    if (gf >= clip_f1 && clip + 2 <= T.n_clips && gf < T.frame_start[clip + 2]) clip += 1;
                    else clip = mb_find_clip(T, gf);
                    clip_f0 = T.frame_start[clip];
                    clip_f1 = T.frame_start[clip + 1];
                }
                const int64_t first = T.clip_off[clip] + (gf - clip_f0) * (int64_t)P.hop;
                src[f] = samples + first;
                src16[f] = reinterpret_cast<const int16_t *>(samples) + first * T.pcm_channels + T.pcm_channel;
                all_aligned = all_aligned && (kPcm ? (T.pcm_channels == 1 && (reinterpret_cast<uintptr_t>(src16[f]) & 15) == 0)
                                                   : (reinterpret_cast<uintptr_t>(src[f]) & 15) == 0);
            }
            constexpr uint32_t kFrameBytes = kPcm ? kN * 2 : kN * 4;

            // ---- 1. the F frames into the warp's slot, `buffer` out of it
            __syncwarp();
            if (!all_aligned) {
                if (lane == 0) bulk_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int f = 0; f < kF; f++) {
                    if (kPcm) {  // also the channel pick of interleaved PCM
                        const int st = T.pcm_channels;
                        for (int i = lane; i < kN; i += 32)
                            reinterpret_cast<int16_t *>(slot)[f * kN + i] = __ldg(src16[f] + (int64_t)i * st);
                    } else {
                        for (int i = lane; i < kN; i += 32) slot[f * kN + i] = __ldg(src[f] + i);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            } else {
                if (lane == 0) {
                    bulk_store_wait_read();
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(bar, kF * kFrameBytes);
#pragma unroll
                    for (int f = 0; f < kF; f++)
                        bulk_load(reinterpret_cast<char *>(slot) + f * kFrameBytes, kPcm ? (const void *)src16[f] : (const void *)src[f],
                                  kFrameBytes, bar, pol_keep);
                }
                __syncwarp();
                mbar_wait(bar, parity);
                parity ^= 1;
            }
            if (!kPcm && want_buffer && lane == 0) {
                if (nfg == kF) {
                    bulk_store(O.buffer + g * kN, slot, kF * kN * 4, pol_stream);  // F consecutive rows
                } else {
                    for (int f = 0; f < nfg; f++) bulk_store(O.buffer + (g + f) * kN, slot + f * kN, kN * 4, pol_stream);
                }
            }

            // ---- 2. pass A load: window, time-domain sums (per frame)
            float2 v[32];
            float esum[kF];
            float2 esum2[kF];  // (even samples, odd samples): one packed FFMA2 per sample pair
            uint32_t sgn_e, sgn_o;  // bit a': sample 2(32 a' + lane) (+1) of the warp is >= 0
            auto pass1 = [&]() {
#pragma unroll
                for (int f = 0; f < kF; f++) esum2[f] = make_float2(0.f, 0.f);
                sgn_e = sgn_o = 0;
#pragma unroll
                for (int a = 0; a < 32; a++) {
                    float2 x;
                    if (kPcm) {
                        const uint32_t raw = reinterpret_cast<const uint32_t *>(slot)[32 * a + lane];  // two samples
                        x = make_float2((float)(int16_t)(raw & 0xffffu) * (1.0f / 32768.0f),
                                        (float)(int16_t)(raw >> 16) * (1.0f / 32768.0f));
                        if (want_buffer && a / kA < nfg)  // `buffer` leaves from registers: the group's rows are contiguous
                            __stcs(reinterpret_cast<float2 *>(O.buffer + g * kN) + 32 * a + lane, x);
                    } else {
                        x = slot2[32 * a + lane];
                    }
                    const float2 w = reinterpret_cast<const float2 *>(S.window)[32 * a + lane];
                    esum2[a / kA] = mbx2::fma(x, x, esum2[a / kA]);
                    if (want_time) {
                        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}"
                            : "+r"(sgn_e) : "f"(x.x), "r"(1u << a));
                        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}"
                            : "+r"(sgn_o) : "f"(x.y), "r"(1u << a));
                    }
                    v[a] = mbx2::mul(x, w);
                }
#pragma unroll
                for (int f = 0; f < kF; f++) esum[f] = esum2[f].x + esum2[f].y;
            };
            pass1();
            float energy[kF];
            bool odd = false;
#pragma unroll
            for (int f = 0; f < kF; f++) {
                energy[f] = mb_warp_sum(esum[f]);
                odd = odd || !(energy[f] >= 0x1p-60f && energy[f] <= 0x1p70f);
            }
            // Frames outside the float32 comfort zone (see kernel_warp.cu): exact power-of-two rescale, per frame.
            // The scale of every frame of the group is kept in the stash (row 17).
            bool any_scaled = false;
            if (lane < kF) stash[17][min(j + lane, kChunk - 1)] = __int_as_float(0);
            __syncwarp();
            if (!kPcm && want_spectrum && odd) {  // (PCM: |x| is 0 or >= 2^-15)
#pragma unroll
                for (int f = 0; f < kF; f++) {
                    if (energy[f] >= 0x1p-60f && energy[f] <= 0x1p70f) continue;
                    float mx = 0.f;
                    for (int i = lane; i < kN; i += 32) mx = fmaxf(mx, fabsf(slot[f * kN + i]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (mx > 0.f && mx < 3.0e38f && (mx < 0x1p-40f || mx > 0x1p40f)) {
                        int e;
                        (void)frexpf(mx, &e);
                        const int ks = max(-100, min(100, -e));
                        if (want_buffer && lane == 0) bulk_store_wait_read();
                        __syncwarp();
                        const float up = ldexpf(1.f, ks);
                        for (int i = lane; i < kN; i += 32) slot[f * kN + i] *= up;
                        if (lane == 0 && j + f < kChunk) stash[17][j + f] = __int_as_float(ks);
                        any_scaled = true;
                    }
                }
                if (any_scaled) {
                    __syncwarp();
                    pass1();
#pragma unroll
                    for (int f = 0; f < kF; f++) energy[f] = mb_warp_sum(esum[f]);
                }
            }
            if (want_time) {
                const uint32_t nxt_dn = __shfl_down_sync(0xffffffffu, sgn_e, 1);
                const uint32_t e0 = __shfl_sync(0xffffffffu, sgn_e, 0);
                uint32_t nxt = nxt_dn, valid = 0xffffffffu;
                if (lane == 31) {
                    nxt = e0 >> 1;
                    // the last sample of every frame has no successor (zcr.js: signal[N] undefined)
                    uint32_t ends = 0;
#pragma unroll
                    for (int f = 0; f < kF; f++) ends |= 1u << (f * kA + kA - 1);
                    valid = ~ends;
                }
                const uint32_t cross = (sgn_e ^ sgn_o), cross2 = (sgn_o ^ nxt) & valid;
#pragma unroll
                for (int f = 0; f < kF; f++) {
                    const uint32_t fm = ((1u << kA) - 1u) << (f * kA);
                    int zc = mb_warp_sum(__popc(cross & fm) + __popc(cross2 & fm));
                    if (!kPcm && !(energy[f] == energy[f])) {  // a NaN sample: recount exactly as zcr.js compares
                        int z = 0;
                        for (int i = lane; i < kN - 1; i += 32) {
                            const float p = slot[f * kN + i], q = slot[f * kN + i + 1];
                            z += ((p >= 0.f && q < 0.f) || (p < 0.f && q >= 0.f)) ? 1 : 0;
                        }
                        zc = mb_warp_sum(z);
                    }
                    if (lane == 0 && f < nfg) stash[1][j + f] = __int_as_float(zc);
                }
            }
#pragma unroll
            for (int f = 0; f < kF; f++)
                if (lane == 0 && f < nfg) stash[0][j + f] = energy[f];  // (of the rescaled samples where a frame was rescaled)

            __syncwarp();
            // mb_adaptive.cuh: rms rounding error of one spectrum bin of frame f of the group, in the frame's own units
            if (kNoise && lane < kF) {
                const int col = min(j + min(lane, nfg - 1), kChunk - 1);
                S.sig[warp][lane] = mb_noise_sigma(stash[0][col], 1.0f / (float)kN) * ldexpf(1.f, -__float_as_int(stash[17][min(j + lane, kChunk - 1)]));
            }
            __syncwarp();
            auto sigma_of = [&](int f) { return S.sig[warp][f]; };
            phase_sync();
            if (want_spectrum) {
                // ---- 3. pass A: F FFTs of A points per lane; twiddle; transpose
                fft_frames<kA>(v, std::make_integer_sequence<int, kF>{});
                __syncwarp();
                if (want_buffer && lane == 0) bulk_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int c = 0; c < 32; c++) {
                    const int f = c / kA, p = c % kA;
                    float2 y = v[f * kA + mbfft::brev<kABits>(p)];
                    if (p > 0) {
                        const float2 t = S.tw32[c * 32 + lane];
                        y = mbx2::cmul(y, t);
                    }
                    slot2[c * kRow + lane] = y;
                }
                phase_sync();
                // ---- pass B: one 32-point FFT per lane (row lp of frame lf) -> X[lp + A q] in v[brev5(q)]
#pragma unroll
                for (int b = 0; b < 32; b++) v[b] = slot2[lane * kRow + b];
                mbfft::fft_reg<32>(v);
                __syncwarp();
                {
                    float2 *xf = slot2 + lf * kMs + lp;
#pragma unroll
                    for (int q = 0; q < 32; q++) xf[kA * q] = v[brev5(q)];
                    if (lp == 0) slot2[lf * kMs + kM] = v[0];  // X[M] := X[0]
                }
                phase_sync();

                // ---- 4. real-FFT split over 32 rows of 32 bins (A rows per frame); spectra out
                float av[32];
                float lgs[kF];
#pragma unroll
                for (int f = 0; f < kF; f++) lgs[f] = 0.f;
                const float sc = P.inv_sqrt_N, hsc = 0.5f * sc;
                const bool l0 = lane == 0;
                float *up_re = O.complex_real + g * kN + lane, *up_im = O.complex_imag + g * kN + lane;
                float *dn_re = O.complex_real + g * kN + (kN - lane) + (l0 ? 0 : 32);
                float *dn_im = O.complex_imag + g * kN + (kN - lane) + (l0 ? 0 : 32);
                float *out_amp = O.amplitude_spectrum + g * kM + lane, *out_pow = O.power_spectrum + g * kM + lane;
                const float2 *xa = slot2 + lane, *xb = slot2 + (kM - lane);
                const float2 *twp = S.twN + lane;
                float pzr = 0.f, pzi = 0.f;
#pragma unroll
                for (int d = 0; d < 32; d++) {
                    const int f = d / kA, r = d % kA;
                    const bool fv = f < nfg;
                    const float2 a = xa[f * kMs + 32 * r];   // X_f[k], k = 32 r + lane
                    const float2 b = xb[f * kMs - 32 * r];   // X_f[M - k]
                    const float2 w = twp[32 * d];
                    // E = a + conj(b) = (sx, dy), F = a - conj(b) = (dx, sy); Z = hsc E + sy w + dx (w.y, -w.x), packed (kernel_warp.cu)
                    const float2 cb = make_float2(b.x, -b.y);
                    const float2 E = mbx2::add(a, cb), F = mbx2::sub(a, cb);
                    const float2 Z = mbx2::fma(E, mbx2::bc(hsc), mbx2::fma(w, mbx2::bc(F.y), mbx2::mul(make_float2(w.y, -w.x), mbx2::bc(F.x))));
                    const float zr = Z.x, zi = Z.y;
                    if (want_cs && fv) {
                        st_stream(up_re + f * kN + 32 * r, zr);
                        st_stream(up_im + f * kN + 32 * r, zi);
                        if (r == 0) {
                            if (l0) {
                                st_stream(O.complex_real + (g + f) * kN + kM, (a.x - a.y) * sc);         // Nyquist bin
                                st_stream(O.complex_imag + (g + f) * kN + kM, (a.x - a.y) * 0.f + 0.f);  // +0 (NaN with the frame)
                            }
                        } else {
                            st_stream(dn_re + f * kN - 32 * r, l0 ? zr : pzr);
                            st_stream(dn_im + f * kN - 32 * r, -(l0 ? zi : pzi));
                        }
                        if (r == kA - 1 && !l0) {  // the frame's last row completes the line its Nyquist bin opened
                            st_stream(dn_re + f * kN - 32 * kA, zr);
                            st_stream(dn_im + f * kN - 32 * kA, -zi);
                        }
                    }
                    pzr = zr;
                    pzi = zi;
                    const float amp = sqrt_approx(fmaf(zr, zr, zi * zi));
                    av[d] = amp;
                    if (want_amp_out && fv) st_stream(out_amp + 32 * d, amp);
                    if (want_pow_out && fv) st_stream(out_pow + 32 * d, __fmul_rn(amp, amp));
                    if (want_log) lgs[f] += log2_approx(amp);
                }
                if (want_log) {
#pragma unroll
                    for (int f = 0; f < kF; f++) {
                        const float t = mb_warp_sum(lgs[f]);
                        if (l0 && f < nfg) stash[12][j + f] = t;
                    }
                }
                if (any_scaled) {
                    // Rare: bring the rescaled frames' stored spectra and amplitudes back to their own units.
                    __syncwarp();
#pragma unroll
                    for (int d = 0; d < 32; d++) {
                        const int f = d / kA;
                        av[d] *= ldexpf(1.f, -__float_as_int(stash[17][min(j + f, kChunk - 1)]));
                    }
                    for (int f = 0; f < nfg; f++) {
                        const int ks = __float_as_int(stash[17][j + f]);
                        if (ks == 0) continue;
                        const float u = ldexpf(1.f, -ks);
                        float *re = O.complex_real + (g + f) * kN, *im = O.complex_imag + (g + f) * kN;
                        float *am = O.amplitude_spectrum + (g + f) * kM, *pw = O.power_spectrum + (g + f) * kM;
                        for (int r = 0; r < kA; r++) {
                            const int k = 32 * r + lane;
                            if (want_cs) {
                                re[k] *= u;
                                im[k] *= u;
                                if (k > 0) { re[kN - k] *= u; im[kN - k] *= u; }
                                else re[kM] *= u;
                            }
                            if (want_amp_out) {
                                const float t = am[k] * u;
                                am[k] = t;
                                if (want_pow_out) pw[k] = __fmul_rn(t, t);
                            } else if (want_pow_out) {
                                pw[k] = pw[k] * u * u;
                            }
                        }
                    }
                }

                phase_sync();
                if (want_blocked) {
                    // ---- 5. blocked layout: lane L = (frame lf, row lp) owns bins [32 lp, 32 lp + 32) of its frame
                    __syncwarp();
#pragma unroll
                    for (int d = 0; d < 32; d++) slot[lane + kAmpStride * d] = av[d];
                    __syncwarp();
                    float ab[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 t = *reinterpret_cast<const float4 *>(slot + kAmpStride * lane + 4 * q);
                        ab[4 * q] = t.x; ab[4 * q + 1] = t.y; ab[4 * q + 2] = t.z; ab[4 * q + 3] = t.w;
                    }
                    float *piece = slot + kPieceOff;  // piece id -> piece + kPf * id: {sum a, sum p, sum w p}
                    float ra = 0.f, rp = 0.f, rr = 0.f;
                    // k-weights count from the piece's own first bin: a strong bin that opens a mel segment then
                    // weighs exactly 0 there (counted from the lane start it left a rounding residue of
                    // 6e-8 x 17 x its power in a filter that may hold a billion times less)
                    float wk = -1.f;
                    double ta = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
                    const float sigma_l = sigma_of(lf);  // this lane's frame
                    // as_float(cf - as_int(a)) ~ theta sigma / a (valid while theta sigma < 1: rescaled frames are flagged outright)
                    const int cf = 0x7EF311C7 + __float_as_int(kMbNoiseTheta * sigma_l) - 0x3F800000;
                    float qn = 0.f;  // sum min(1 / a, 1 / (theta sigma))^2 over the lane's bins (mb_adaptive.cuh)
                    uint32_t paddr = smem_u32(piece + kPf * (slot_base + lane));
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        if (want_moments || want_rolloff) {
                            const double ad = (double)ab[i];
                            ta += ad;
                            if (want_moments) {
                                if (kNoise && ((i + (i >> 2)) & 3) == 0) {  // bins 0, 7, 10, 13, 16, 23, 26, 29 of the block: a quarter, off every comb
                                    // theta sigma / a from the exponent trick (ONE integer subtraction, within ~20 %: this feeds a
                                    // bound, and MUFU.RCP with its range fix-up would cost six more instructions per bin), squared
                                    // and clipped to 1 by the multiplier's .sat; 0 and denormals come out huge: a floor bin counts 1
                                    float u_;
                                    asm("mul.sat.f32 %0, %1, %1;" : "=f"(u_) : "f"(__int_as_float(cf - __float_as_int(ab[i]))));
                                    qn += u_;
                                }
                                t1 = fma(ad, (double)i, t1);
                                if (want_high) {
                                    t2 = fma(ad, (double)(i * i), t2);
                                    t3 = fma(ad, (double)(i * i * i), t3);
                                    t4 = fma(ad, (double)(i * i * i * i), t4);
                                }
                            }
                        }
                        if (!want_pieces) continue;
                        float keep;
                        if constexpr (kPf == 4) {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                                "and.b32 t, %6, %7;\n\t"
                                "setp.ne.u32 p, t, 0;\n\t"
                                "@p st.shared.v4.f32 [%1], {%2, %3, %4, %5};\n\t"
                                "@p add.u32 %1, %1, 16;\n\t"
                                "selp.f32 %0, 0f00000000, 0f3F800000, p;\n\t}"
                                : "=f"(keep), "+r"(paddr)
                                : "f"(ra), "f"(rp), "f"(rr), "f"(0.f), "r"(bmask), "r"(1u << i)
                                : "memory");
                        } else {
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
                                "and.b32 t, %5, %6;\n\t"
                                "setp.ne.u32 p, t, 0;\n\t"
                                "@p st.shared.f32 [%1], %2;\n\t"
                                "@p st.shared.f32 [%1+4], %3;\n\t"
                                "@p st.shared.f32 [%1+8], %4;\n\t"
                                "@p add.u32 %1, %1, 12;\n\t"
                                "selp.f32 %0, 0f00000000, 0f3F800000, p;\n\t}"
                                : "=f"(keep), "+r"(paddr)
                                : "f"(ra), "f"(rp), "f"(rr), "r"(bmask), "r"(1u << i)
                                : "memory");
                        }
                        const float pf = __fmul_rn(ab[i], ab[i]);
                        wk = fmaf(wk, keep, keep);  // bins since the piece began: 0 at a boundary, else one more
                        ra = fmaf(ra, keep, ab[i]);
                        rp = fmaf(rp, keep, pf);
                        rr = fmaf(wk, pf, rr * keep);
                    }
                    if (want_pieces) {
                        if constexpr (kPf == 4)
                            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(paddr), "f"(ra), "f"(rp), "f"(rr), "f"(0.f)
                                         : "memory");
                        else
                            asm volatile("st.shared.f32 [%0], %1;\n\tst.shared.f32 [%0+4], %2;\n\tst.shared.f32 [%0+8], %3;" ::"r"(paddr),
                                         "f"(ra), "f"(rp), "f"(rr)
                                         : "memory");
                    }
                    if (want_moments) {
                        const double c = (double)(32 * lp), c2 = c * c;  // bins counted inside the lane's own frame
                        const double s1 = fma(c, ta, t1);
                        const double r0 = group_sum_d<kA>(ta), r1 = group_sum_d<kA>(s1);
                        // Q_0 and Q_4 of mb_adaptive.cuh (k := the last bin of the lane's block)
                        const float ql = 6.4f * qn,  /* (4: every fourth bin was looked at; 1.6: the exponent trick's worst case, squared) */ k4 = (float)(32 * lp + 31) * (float)(32 * lp + 31);
                        const float q0 = kNoise ? group_sum_f<kA>(ql) : 0.f, q4 = kNoise ? group_sum_f<kA>(ql * (k4 * k4)) : 0.f;
                        if (lp == 0 && lf < nfg) {
                            if (!want_rolloff) stash_put_d(stash, 2, j + lf, r0);
                            stash_put_d(stash, 4, j + lf, r1);
                            stash[18][j + lf] = q0;
                            stash[19][j + lf] = q4;
                        }
                        if (want_high) {
                            const double s2 = fma(c2, ta, fma(2.0 * c, t1, t2));
                            const double s3 = fma(c2 * c, ta, fma(3.0 * c2, t1, fma(3.0 * c, t2, t3)));
                            const double s4 = fma(c2 * c2, ta, fma(4.0 * c2 * c, t1, fma(6.0 * c2, t2, fma(4.0 * c, t3, t4))));
                            const double r2 = group_sum_d<kA>(s2), r3 = group_sum_d<kA>(s3), r4 = group_sum_d<kA>(s4);
                            if (lp == 0 && lf < nfg) {
                                stash_put_d(stash, 6, j + lf, r2);
                                stash_put_d(stash, 8, j + lf, r3);
                                stash_put_d(stash, 10, j + lf, r4);
                            }
                        }
                    }
                    if (want_rolloff) {
                        // per frame: lane totals scanned in double inside the frame's A lanes, then the one lane the
                        // 0.99 threshold falls into is scanned bin by bin by the whole warp
                        double ia = ta;
#pragma unroll
                        for (int o = 1; o < kA; o <<= 1) {
                            const double ya = __shfl_up_sync(0xffffffffu, ia, o, kA);
                            if (lp >= o) ia += ya;
                        }
                        const double total_a = __shfl_sync(0xffffffffu, ia, kA - 1, kA);
                        const double thr = 0.99 * total_a;
                        const uint32_t under = __ballot_sync(0xffffffffu, (ia - ta) <= thr);
#pragma unroll
                        for (int f = 0; f < kF; f++) {
                            const double tot_f = __shfl_sync(0xffffffffu, total_a, f * kA);
                            const double thr_f = 0.99 * tot_f;
                            const uint32_t bits = (under >> (f * kA)) & ((1u << kA) - 1u);
                            int rbin = kM;
                            if (tot_f > thr_f && bits != 0u) {
                                const int lc = f * kA + (31 - __clz(bits));
                                const double base = __shfl_sync(0xffffffffu, ia - ta, lc);
                                double x = (double)slot[kAmpStride * lc + lane];
                                double incl = x;
#pragma unroll
                                for (int o = 1; o < 32; o <<= 1) {
                                    const double y = __shfl_up_sync(0xffffffffu, incl, o);
                                    if (lane >= o) incl += y;
                                }
                                const uint32_t ok = __ballot_sync(0xffffffffu, base + (incl - x) <= thr_f);
                                rbin = 32 * (lc - f * kA) + (31 - __clz(ok));
                            }
                            if (l0 && f < nfg) {
                                stash[13][j + f] = __int_as_float(rbin);
                                if (want_moments) stash_put_d(stash, 2, j + f, tot_f);
                            }
                        }
                    }
                    phase_sync();  // pieces visible; the blocked amplitudes are spent and become scratch
                    float *sp_s = slot;                                  // [24 F] specific loudness
                    float *rise_s = slot + 24 * kF;                      // [27 F]
                    float *fall_s = rise_s + 27 * kF;                    // [27 F]
                    float *lge_s = fall_s + 27 * kF;                     // [26 F]
                    // Lane (lf, lp) works on its own frame, as in the blocked layout: bands / segments lp, lp + A, ... .
                    // The per-frame sums (loudness total, its maximum, the sharpness sum, the noise bounds of
                    // mb_adaptive.cuh) then stay in registers and close with log2(A) shuffles inside the frame's lanes.
                    if (want_bark) {
                        float total = 0.f, mx = 0.f, sharp = 0.f, nu = 0.f;
                        for (int bnd = lp; bnd < MB_NUM_BARK_BANDS; bnd += kA) {
                            const int ts = lf * MB_WARP_SEGMENTS + bnd;
                            float bsum = 0.f;
                            for (int it = S.seg_ptr[ts]; it < S.seg_ptr[ts + 1]; it++) bsum += piece[kPf * S.seg_items[it]];
                            const float sp = pow023_approx(bsum);
                            sp_s[lf * MB_NUM_BARK_BANDS + bnd] = sp;
                            total += sp;
                            mx = (sp > mx) ? sp : mx;  // NaN never compares greater (perceptualSpread.js:6)
                            if (bnd >= 1 && bnd <= 15) sharp = fmaf((float)bnd, sp, sharp);
                            if (kNoise) nu += mb_noise_band(bsum, sp, S.noise_c[0][bnd], sigma_l);
                        }
#pragma unroll
                        for (int o = kA / 2; o > 0; o >>= 1) {
                            total += __shfl_xor_sync(0xffffffffu, total, o);
                            sharp += __shfl_xor_sync(0xffffffffu, sharp, o);
                            nu += __shfl_xor_sync(0xffffffffu, nu, o);
                            const float other = __shfl_xor_sync(0xffffffffu, mx, o);
                            mx = (other > mx) ? other : mx;
                        }
                        if (lp == 0 && lf < nfg) {
                            stash[14][j + lf] = total;
                            stash[15][j + lf] = mx;
                            stash[16][j + lf] = sharp;
                            if (kNoise) stash[20][j + lf] = nu;
                        }
                        if (mb_has(mask, MB_FEAT_LOUDNESS)) {
                            __syncwarp();
                            for (int idx = lane; idx < MB_NUM_BARK_BANDS * nfg; idx += 32) O.loudness_specific[g * MB_NUM_BARK_BANDS + idx] = sp_s[idx];
                        }
                    }
                    if (want_mfcc) {
                        for (int sg = lp; sg < MB_NUM_MEL_FILTERS + 1; sg += kA) {
                            const int ts = lf * MB_WARP_SEGMENTS + MB_NUM_BARK_BANDS + sg, e0 = S.mel_edge[sg];
                            const float inv = S.mel_inv[sg];
                            float rise = 0.f, fall = 0.f;
                            for (int it = S.seg_ptr[ts]; it < S.seg_ptr[ts + 1]; it++) {
                                const int pc = S.seg_items[it];
                                const float pp = piece[kPf * pc + 1], pz = piece[kPf * pc + 2];
                                float up = fmaf((float)((int)S.piece_edge[pc] - e0), pp, pz) * inv;
                                up = (up < 0.f) ? 0.f : up;
                                rise += up;
                                fall += fmaxf(pp - up, 0.f);
                            }
                            rise_s[lf * (MB_NUM_MEL_FILTERS + 1) + sg] = rise;
                            fall_s[lf * (MB_NUM_MEL_FILTERS + 1) + sg] = fall;
                        }
                        __syncwarp();
                        float ndl = 0.f;
                        for (int m = lp; m < MB_NUM_MEL_FILTERS; m += kA) {
                            const float melE = rise_s[lf * (MB_NUM_MEL_FILTERS + 1) + m] + fall_s[lf * (MB_NUM_MEL_FILTERS + 1) + m + 1];
                            lge_s[lf * MB_NUM_MEL_FILTERS + m] = ln_approx(melE);
                            if (kNoise) ndl += mb_noise_mel(melE, S.noise_c[1][m], S.noise_c[2][m], sigma_l);
                        }
                        if (kNoise) {
#pragma unroll
                            for (int o = kA / 2; o > 0; o >>= 1) ndl += __shfl_xor_sync(0xffffffffu, ndl, o);
                            if (lp == 0 && lf < nfg) stash[21][j + lf] = ndl;
                        }
                        __syncwarp();
                        for (int it0 = 0; it0 < MB_NUM_MFCC * kF; it0 += 32) {
                            const int idx = it0 + lane;
                            if (idx < MB_NUM_MFCC * kF) {
                                const int f = idx / MB_NUM_MFCC, k = idx % MB_NUM_MFCC;
                                float acc = 0.f;
#pragma unroll 2
                                for (int n = 0; n < MB_NUM_MEL_FILTERS; n++)
                                    acc = fmaf(S.dct[k + n * MB_NUM_MFCC], lge_s[f * MB_NUM_MEL_FILTERS + n], acc);
                                if (f < nfg) O.mfcc[g * MB_NUM_MFCC + idx] = acc * (1.0f / (float)MB_NUM_MFCC);
                            }
                        }
                    }
                }
            }
        }  // groups of the chunk

        // ---- 6. one frame per lane: the "number" features of the chunk, coalesced
        __syncwarp();
        bool need_exact = false;
        if (lane < nfc) {
            MbFrameSums F;
            const int ks = want_spectrum ? __float_as_int(stash[17][lane]) : 0;
            F.energy = ldexp((double)stash[0][lane], -2 * ks);
            F.zcr = __float_as_int(stash[1][lane]);
            F.s0 = stash_get_d(stash, 2, lane);
            F.s1 = stash_get_d(stash, 4, lane);
            F.s2 = stash_get_d(stash, 6, lane);
            F.s3 = stash_get_d(stash, 8, lane);
            F.s4 = stash_get_d(stash, 10, lane);
            F.log2sum = (double)stash[12][lane] - (double)(kM * ks);
            F.rolloff_bin = __float_as_int(stash[13][lane]);
            const int64_t gg = g0 + lane;
            MbMoments MO;
            mb_store_scalars(P, O, gg, F, &MO);
            if (want_bark) {
                const double total = (double)stash[14][lane], mx = (double)stash[15][lane];
                const double sharp = (double)stash[16][lane] + P.sharp_const;
                if (mb_has(mask, MB_FEAT_LOUDNESS)) O.loudness_total[gg] = (float)total;
                if (mb_has(mask, MB_FEAT_PERCEPTUAL_SPREAD)) {
                    const double r = (total - mx) / total;
                    O.perceptual_spread[gg] = (float)(r * r);
                }
                if (mb_has(mask, MB_FEAT_PERCEPTUAL_SHARPNESS)) O.perceptual_sharpness[gg] = (float)(sharp * (0.11 / total));
            }
            if (kNoise && want_spectrum && T.fix_count != nullptr) {
                MbNoiseFrame NF;
                NF.sigma = mb_noise_sigma((float)F.energy, 1.0f / (float)kN);  // (frames too small for this float were rescaled: flagged below)
                NF.q0 = stash[18][lane];
                NF.q4 = stash[19][lane];
                NF.sum_u = stash[20][lane];
                NF.sum_dl = stash[21][lane];
                NF.total = stash[14][lane];
                NF.sharp = stash[16][lane] + (float)P.sharp_const;
                need_exact = ks != 0 || mb_noise_needs_exact(P, mask, F, MO, NF);
            }
        }
        mb_noise_append(T, need_exact, g0 + lane);  // frames of this chunk that the exact-FFT kernel redoes
        __syncwarp();
    }
    if (lane == 0) bulk_store_wait_all();
}

}  // namespace

size_t mb_warpmf_smem_bytes() { return sizeof(Smem) + 128; }

cudaError_t mb_launch_warpmf(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                             int num_sms, cudaStream_t stream) {
    const size_t smem = mb_warpmf_smem_bytes();
    const bool pcm = T.pcm_channels > 0;
    // fixed feature sets: all, and all but the big arrays (the config-3 set of kernel_warp.cu trips a ptxas
    // register-allocation failure in this kernel -- "register count of 7", the predicate file -- and runs
    // through the run-time-mask instantiation instead)
    const uint32_t m = P.mask & MB_ALL_FEATURES;
    constexpr uint32_t kNoArrays = MB_ALL_FEATURES & ~(MB_FEATURE_BIT(MB_FEAT_BUFFER) | MB_FEATURE_BIT(MB_FEAT_COMPLEX_SPECTRUM) |
                                                       MB_FEATURE_BIT(MB_FEAT_AMPLITUDE_SPECTRUM) | MB_FEATURE_BIT(MB_FEAT_POWER_SPECTRUM));
#define MB_PICK2(A, MASK) (pcm ? mb_warpmf_kernel<A, true, MASK> : mb_warpmf_kernel<A, false, MASK>)
#define MB_PICK(A) (m == MB_ALL_FEATURES ? MB_PICK2(A, MB_ALL_FEATURES) : m == kNoArrays ? MB_PICK2(A, kNoArrays) : MB_PICK2(A, 0u))
    // BASELINE config 1 (the feature list of the reference's own demo run) at the reference's default bufferSize
    constexpr uint32_t kConfig1 = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) | MB_FEATURE_BIT(MB_FEAT_ZCR) |
                                  MB_FEATURE_BIT(MB_FEAT_AMPLITUDE_SPECTRUM) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID);
    auto kernel = P.N == 1024 ? MB_PICK(16) : P.N == 512 ? (m == kConfig1 ? MB_PICK2(8, kConfig1) : MB_PICK(8))
                  : (m == MB_ALL_FEATURES ? MB_PICK2(4, MB_ALL_FEATURES) : MB_PICK2(4, 0u));
#undef MB_PICK
#undef MB_PICK2
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t chunks = (T.total_frames + kChunk - 1) / kChunk;
    int64_t grid = (chunks + kWarps - 1) / kWarps;
    if (grid > num_sms) grid = num_sms;
    if (grid < 1) return cudaSuccess;
    (void)cudaGetLastError();
    kernel<<<(unsigned)grid, kThreads, smem, stream>>>(P, T, samples, O, chunks);
    return cudaGetLastError();
}
