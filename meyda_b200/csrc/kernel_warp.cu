// kernel_warp.cu -- warp-per-frame kernel for bufferSize 2048 (the headline
// configuration), float32 FFT.
//
// One warp owns one frame at a time and never synchronises with another warp:
//   1. lane 0 pulls the frame's 2048 raw samples into the warp's shared-memory
//      slot with one TMA bulk copy (cp.async.bulk + mbarrier); `buffer` leaves
//      the same slot with a TMA bulk store, untouched by registers;
//   2. the windowed frame is read as 1024 packed complex values, 32 per lane,
//      and transformed as 32 x 32: a register radix-2^5 FFT per lane, a
//      twiddle, a transpose through the warp's slot, a second register FFT;
//   3. the real-FFT split produces Z[k] for k = lane + 32 d; complex /
//      amplitude / power spectra go straight from registers to HBM in 128-byte
//      rows, and the spectral moments, log sum, and time-domain sums are
//      accumulated on the way;
//   4. the amplitudes take one more trip through the slot into a blocked
//      layout (32 consecutive bins per lane) for the prefix sums that give
//      rolloff, the 24 Bark bands and the 26 mel filters, which the lanes then
//      finish (one band / filter / cepstral coefficient per lane);
//   5. per-frame scalars are parked in a 32-frame stash and turned into the
//      twelve "number" features once per 32 frames with one frame per lane,
//      so the float64 divisions/sqrt/exp and the scalar stores run at full
//      lane efficiency and leave as coalesced 128-byte rows.
// A warp's work unit is a chunk of 32 consecutive output frames.
//
// Reference path being replaced: src/meyda.js:69-91,104-114,158-168,
// lib/jsfft/fft.js:123-208 and the extractor files under src/extractors/.
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

constexpr int kP = 32;             // points per lane per pass
constexpr int kM = kP * kP;        // 1024 complex points
constexpr int kN = 2 * kM;         // bufferSize 2048
#ifndef MB_WARPS
#define MB_WARPS 16
#endif
constexpr int kWarps = MB_WARPS;  // warps per persistent CTA: 128 registers each at 16
constexpr int kThreads = kWarps * 32;
#ifndef MB_LOCK_WARPS
#define MB_LOCK_WARPS 2
#endif
constexpr int kLockWarps = MB_LOCK_WARPS;  // warps that step through a frame's phases together (0: free-running); see the kernel
constexpr int kRow = kP + 1;                  // float2 stride of a transpose row
constexpr int kSlotFloats = 2 * kP * kRow;    // 2112 floats = 8448 B per warp
constexpr int kAmpStride = 36;                // floats per lane in the blocked amplitude layout
constexpr int kPieceOff = 1152;               // float offset of the piece area (after 32*36 amps)
constexpr int kPieces = MB_WARP_PIECES;       // 96: 64 boundary pieces + 32 lane heads
constexpr int kStashRows = 22;                // floats per frame in the scalar stash (18 .. 21: mb_adaptive.cuh statistics)

static_assert(kPieceOff * 4 + kPieces * 16 <= kSlotFloats * 4, "band pieces must fit the warp slot");

// ---- shared memory carve-up (dynamic)
struct Smem {
    float2 tw32[kP * kP];        // exp(+2 pi i b c / 1024) at [c*32 + b]
    float2 twN[kM];              // exp(+2 pi i k / 2048)
    float window[kN];
    float dct2[MB_NUM_MFCC * 32];  // [n][lane]: lane = k + 13 h holds dct[k + 13 (n + 13 h)] (mfcc.js:72-83), 0 for lanes >= 26
    float mel_inv[MB_NUM_MEL_FILTERS + 2];
    int mel_edge[MB_NUM_MEL_FILTERS + 2];
    int piece_edge[kPieces];
    float noise_c[3][32];        // mb_adaptive.cuh: band_c[lane], mel_c1[lane], mel_c2[lane]
    int seg_ptr[MB_WARP_SEGMENTS + 1];
    unsigned char seg_items[MB_WARP_MAX_ITEMS];
    unsigned long long bar[kWarps];
    float stash[kWarps][kStashRows][kChunk];
    alignas(128) float slot[kWarps][kSlotFloats];  // TMA destination / float4 reads: 16-byte alignment required
};
static_assert(offsetof(Smem, slot) % 128 == 0 && (kSlotFloats * 4) % 16 == 0, "warp slots must stay 16-byte aligned");
static_assert(offsetof(Smem, twN) % 8 == 0 && offsetof(Smem, window) % 8 == 0, "float2 tables");

// ---- 32-point FFT in registers (mb_fft.cuh): natural order in, X[k] left in v[brev5(k)].
__device__ __forceinline__ void fft32(float2 (&v)[32]) { mbfft::fft_reg<32>(v); }
__host__ __device__ constexpr int brev5(int k) { return mbfft::brev<5>(k); }

__device__ __forceinline__ double warp_sum_d(double v) { return mb_warp_sum(v); }

// Sums four doubles across the warp with 6 shuffle steps instead of 20: the first two steps fold the four
// values onto lane bits 4 and 3 (a lane keeps one value of a pair and hands over the other), the last three
// are plain.  Lane l returns the total of value (l >> 3) & 3.
__device__ __forceinline__ double warp_sum4_d(double a0, double a1, double a2, double a3, int lane) {
    const bool h16 = lane & 16, h8 = lane & 8;
    const double b0 = (h16 ? a2 : a0) + __shfl_xor_sync(0xffffffffu, h16 ? a0 : a2, 16);  // a0 | a2
    const double b1 = (h16 ? a3 : a1) + __shfl_xor_sync(0xffffffffu, h16 ? a1 : a3, 16);  // a1 | a3
    double c = (h8 ? b1 : b0) + __shfl_xor_sync(0xffffffffu, h8 ? b0 : b1, 8);            // a0, a1 | a2, a3
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
    return c;
}

// kMask: a compile-time feature set (0 = take the plan's at run time).  With the set known the per-bin feature
// tests fold away; instantiated for every feature (the headline configuration), for BASELINE config 3 (mfcc +
// the four spectral moments) and for "everything but the four big arrays".
// kPcm: `samples` holds 16-bit PCM (MbClipTable::pcm_channels interleaved channels); a frame arrives as 4 KB
// instead of 8 KB and is converted in pass 1 (x = s / 32768, exact), after which nothing differs -- the results
// are bit-identical to the float32 path on the converted samples.  `buffer` then leaves from registers.
template <uint32_t kMask, bool kPcm>
__global__ void __launch_bounds__(kThreads, 1)
mb_warp2048_kernel(const __grid_constant__ MbDevPlan P, const __grid_constant__ MbClipTable T,
                   const float *__restrict__ samples, const __grid_constant__ mb_outputs O, const int64_t total_chunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // Keep the shared address space visible to the compiler (LDS/STS, not generic LD/ST): no integer round trip.
    // The kernel has no static shared memory, so the dynamic window starts at the (1 KB aligned) window base.
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    if (smem_u32(smem_raw) & 127u) __trap();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mask = kMask ? kMask : P.mask;
    const MbWarpTables *__restrict__ WT = P.warp_tables;

    // ---- CTA-wide tables into shared memory (once per persistent CTA)
    for (int i = tid; i < kP * kP; i += kThreads) S.tw32[i] = WT->tw32[i];
    {  // split twiddles pre-multiplied by 0.5 / sqrt(N): Z = h (E-part) + (h w) (O-part)
        const float h = 0.5f * P.inv_sqrt_N;
        for (int i = tid; i < kM; i += kThreads) S.twN[i] = make_float2(P.twN[i].x * h, P.twN[i].y * h);
    }
    for (int i = tid; i < kN; i += kThreads) S.window[i] = P.window[i];
    for (int i = tid; i < MB_NUM_MFCC * 32; i += kThreads) {
        const int n = i >> 5, l = i & 31, k = l % MB_NUM_MFCC, h = l / MB_NUM_MFCC;
        S.dct2[i] = l < 2 * MB_NUM_MFCC ? P.dct[k + MB_NUM_MFCC * (n + MB_NUM_MFCC * h)] : 0.f;
    }
    if (tid < MB_NUM_MEL_FILTERS + 2) {
        S.mel_edge[tid] = P.mel[tid];
        S.mel_inv[tid] = tid < MB_NUM_MEL_FILTERS + 1 ? P.mel_inv_width[tid] : 0.f;
    }
    if (tid < kPieces) S.piece_edge[tid] = WT->piece_edge[tid];
    if (tid < 32) {
        S.noise_c[0][tid] = tid < MB_NUM_BARK_BANDS ? P.noise->band_c[tid] : 0.f;
        S.noise_c[1][tid] = tid < MB_NUM_MEL_FILTERS ? P.noise->mel_c1[tid] : 0.f;
        S.noise_c[2][tid] = tid < MB_NUM_MEL_FILTERS ? P.noise->mel_c2[tid] : 0.f;
    }
    if (tid <= MB_WARP_SEGMENTS) S.seg_ptr[tid] = WT->seg_ptr[tid];
    if (tid < MB_WARP_MAX_ITEMS) S.seg_items[tid] = WT->seg_items[tid];
    if (lane == 0) mbar_init(&S.bar[warp], 1);
    __syncthreads();

    const uint32_t bmask = WT->lane_bmask[lane];     // boundaries inside this lane's 32 blocked bins
    const int slot_base = WT->lane_slot_base[lane];  // how many boundaries lie below this lane's first bin

    const bool want_buffer = mb_has(mask, MB_FEAT_BUFFER);
    const bool want_time = (mask & (MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) | MB_FEATURE_BIT(MB_FEAT_ZCR)));
    const bool want_cs = mb_has(mask, MB_FEAT_COMPLEX_SPECTRUM);
    const bool want_amp_out = mb_has(mask, MB_FEAT_AMPLITUDE_SPECTRUM);
    const bool want_pow_out = mb_has(mask, MB_FEAT_POWER_SPECTRUM);
    const bool want_moments = (mask & (MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS) |
                MB_FEATURE_BIT(MB_FEAT_SPECTRAL_FLATNESS) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SLOPE)));
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

    int64_t clip = 0, clip_f0 = 0, clip_f1 = 0;  // cached clip of the previous frame: frames [clip_f0, clip_f1)
    if (T.n_clips > 0) clip_f1 = T.frame_start[1];

    const int64_t warp_global = (int64_t)blockIdx.x * kWarps + warp;
    const int64_t warp_stride = (int64_t)gridDim.x * kWarps;

    // Lock-step phases (as in kernel_warp_mf.cu): warps march in groups of kLockWarps that meet at a named barrier
    // after every phase (load + pass 1, the two FFT passes, the split, the blocked sums), so that fewer distinct
    // stretches of the ~70 KB of straight-line code compete for the instruction cache.  PAIRS here: this kernel waits
    // on HBM as well, and warps that wait for their frames together lose what the shared instruction stream gains --
    // one box, round 2: free-running 129.3 M frames/s (full set) / 238.5 M (mfcc + moments), pairs 129.6 / 243.5,
    // groups of four 125.6, groups of eight 119 / 213.  0 = free-running.
    const int n_phase_bars = 1 + (want_spectrum ? 2 + (want_blocked ? 2 : 0) : 0);
    auto phase_sync = [&]() {
        __syncwarp();
        if constexpr (kLockWarps > 0)
            asm volatile("bar.sync %0, %1;" ::"r"(1 + warp / (kLockWarps > 0 ? kLockWarps : 1)), "n"(32 * (kLockWarps > 0 ? kLockWarps : 1)) : "memory");
    };
    (void)warp_global;
    for (int64_t base = (int64_t)blockIdx.x * kWarps; base < total_chunks; base += warp_stride) {
        const int64_t ch = base + warp;
        const int64_t g0 = ch * kChunk;
        const int nfc = ch < total_chunks ? (int)min((int64_t)kChunk, T.total_frames - g0) : 0;

        for (int j = 0; j < (kLockWarps > 0 ? kChunk : nfc); j++) {
            if (j >= nfc) {
                for (int b = 0; b < n_phase_bars; b++) phase_sync();
                continue;
            }
            const int64_t g = g0 + j;
            // ---- which clip (frames ascend, so mostly the cached one or its successor)
            if (g >= clip_f1 || g < clip_f0) {
                if (g >= clip_f1 && clip + 2 <= T.n_clips && g < T.frame_start[clip + 2]) clip += 1;
                else clip = mb_find_clip(T, g);
                clip_f0 = T.frame_start[clip];
                clip_f1 = T.frame_start[clip + 1];
            }
            const int64_t first = T.clip_off[clip] + (g - clip_f0) * (int64_t)P.hop;
            const float *src = samples + first;  // (float32 input)
            const int16_t *src16 = reinterpret_cast<const int16_t *>(samples) + first * T.pcm_channels + T.pcm_channel;
            constexpr uint32_t kFrameBytes = kPcm ? kN * 2 : kN * 4;

            // ---- 1. frame into the warp's slot (TMA), buffer out of it (TMA)
            __syncwarp();  // every lane is done with the slot's previous contents
            // TMA needs a 16-byte aligned source: a frame that starts elsewhere (odd clip offsets, hop not a
            // multiple of 4) is copied in by the lanes instead; everything after that is the same.
            const bool src_aligned = kPcm ? (T.pcm_channels == 1 && (reinterpret_cast<uintptr_t>(src16) & 15) == 0)
                                          : (reinterpret_cast<uintptr_t>(src) & 15) == 0;
            if (!src_aligned) {
                if (lane == 0) bulk_store_wait_read();
                __syncwarp();
                if (kPcm) {  // also the channel pick of interleaved PCM
                    const int st = T.pcm_channels;
                    for (int i = lane; i < kN; i += 32) reinterpret_cast<int16_t *>(slot)[i] = __ldg(src16 + (int64_t)i * st);
                } else {
                    for (int i = lane; i < kN; i += 32) slot[i] = __ldg(src + i);
                }
                // the `buffer` bulk store below reads the slot through the async proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
            } else {
#ifdef MB_EXP_NOWAIT  // timing experiment only (results are garbage): how much does the frame load + wait cost?
            if (j == 0) {
#endif
            if (lane == 0) {
                bulk_store_wait_read();  // an earlier frame's `buffer` store has finished reading the slot
                // generic-proxy accesses to the slot are ordered before the async-proxy write that follows
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, kFrameBytes);
                bulk_load(slot, kPcm ? (const void *)src16 : (const void *)src, kFrameBytes, bar, pol_keep);
                // the samples the next frame adds, towards L2 while this frame is being worked on
                if (j + 1 < nfc && g + 1 < clip_f1)
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kPcm ? (const void *)(src16 + kN)
                                                                                           : (const void *)(src + kN)),
                                 "r"((uint32_t)(P.hop * (kPcm ? 2 : 4)))
                                 : "memory");
            }
            __syncwarp();
            mbar_wait(bar, parity);
            parity ^= 1;
#ifdef MB_EXP_NOWAIT
            }
#endif
            }
#ifdef MB_EXP_NOSTORE  // timing experiment only: all the arithmetic, none of the spectra/buffer traffic
            const bool exp_store = (total_chunks < 0);
#else
            const bool exp_store = true;
#endif
            if (!kPcm && want_buffer && exp_store && lane == 0) bulk_store(O.buffer + g * kN, slot, kN * 4, pol_stream);

            // ---- 2. pass 1: window, time-domain sums, FFT32 over a for b = lane
            float2 v[32];
            float esum;
            float2 esum2;  // (even samples, odd samples): one packed FFMA2 per sample pair
            uint32_t sgn_e, sgn_o;  // bit a: sample 2(32a+lane) (+1) is >= 0
            auto pass1 = [&]() {
                esum2 = make_float2(0.f, 0.f);
                sgn_e = sgn_o = 0;
#pragma unroll
                for (int a = 0; a < 32; a++) {
                    float2 x;
                    if (kPcm) {
                        const uint32_t raw = reinterpret_cast<const uint32_t *>(slot)[32 * a + lane];  // two samples
                        x = make_float2((float)(int16_t)(raw & 0xffffu) * (1.0f / 32768.0f),
                                        (float)(int16_t)(raw >> 16) * (1.0f / 32768.0f));
                        if (want_buffer && exp_store)
                            __stcs(reinterpret_cast<float2 *>(O.buffer + g * kN) + 32 * a + lane, x);
                    } else {
                        x = slot2[32 * a + lane];
                    }
                    const float2 w = reinterpret_cast<const float2 *>(S.window)[32 * a + lane];
                    esum2 = mbx2::fma(x, x, esum2);
                    if (want_time) {  // (compare + predicated OR: two instructions per sample)
                        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}"
                            : "+r"(sgn_e) : "f"(x.x), "r"(1u << a));
                        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}"
                            : "+r"(sgn_o) : "f"(x.y), "r"(1u << a));
                    }
                    v[a] = mbx2::mul(x, w);
                }
                esum = esum2.x + esum2.y;
            };
            pass1();
            float energy = mb_warp_sum(esum);
            // Frames whose samples all sit below 2^-40 (or reach above 2^40) would under/overflow the
            // float32 squares in |Z|; the reference squares in float64.  Such a frame (rare: decayed
            // tails) is rescaled by an exact power of two in place and redone; Z and |Z| are scaled back
            // on the way out.  The frame's energy is the cheap trigger (all |x| < 2^-40 puts it under 2^-69,
            // one |x| > 2^40 above 2^80; zero and NaN take the look too), the largest sample decides.
            int kscale = 0;
            {
                if (!kPcm && want_spectrum && !(energy >= 0x1p-60f && energy <= 0x1p70f)) {  // (PCM: |x| is 0 or >= 2^-15)
                    float mx = 0.f;
                    for (int i = lane; i < kN; i += 32) mx = fmaxf(mx, fabsf(slot[i]));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                    if (mx > 0.f && mx < 3.0e38f && (mx < 0x1p-40f || mx > 0x1p40f)) {
                        int e;
                        (void)frexpf(mx, &e);
                        kscale = max(-100, min(100, -e));
                        if (want_buffer && lane == 0) bulk_store_wait_read();  // the slot is about to change
                        __syncwarp();
                        const float up = ldexpf(1.f, kscale);
                        for (int i = lane; i < kN; i += 32) slot[i] *= up;
                        __syncwarp();
                        pass1();
                        energy = mb_warp_sum(esum);
                    }
                }
            }
            const float unscale = ldexpf(1.f, -kscale);
            // mb_adaptive.cuh: the rms rounding error of one spectrum bin, in the units of the amplitudes used below
            const float sigma = mb_noise_sigma(energy, 1.0f / (float)kN) * unscale;
            // as_float(cf - as_int(a)) ~ theta sigma / a (valid while theta sigma < 1: rescaled frames are flagged outright)
            const int cf = 0x7EF311C7 + __float_as_int(kMbNoiseTheta * sigma) - 0x3F800000;
            int zcr = 0;
            if (want_time) {
                // crossings inside a pair (2m, 2m+1), and between 2m+1 and 2m+2 (= lane+1, or lane 0 of the next row)
                int z = __popc(sgn_e ^ sgn_o);
                uint32_t nxt = __shfl_down_sync(0xffffffffu, sgn_e, 1);
                const uint32_t e0 = __shfl_sync(0xffffffffu, sgn_e, 0);
                uint32_t valid = 0xffffffffu;
                if (lane == 31) {
                    nxt = e0 >> 1;
                    valid = 0x7fffffffu;  // sample 2047 has no successor in the frame (zcr.js: signal[N] undefined)
                }
                z += __popc((sgn_o ^ nxt) & valid);
                zcr = mb_warp_sum(z);
                if (!kPcm && !(energy == energy)) {
                    // a NaN sample: zcr.js compares are all false around it; recount exactly
                    z = 0;
                    for (int i = lane; i < kN - 1; i += 32) {
                        const float p = slot[i], q = slot[i + 1];
                        z += ((p >= 0.f && q < 0.f) || (p < 0.f && q >= 0.f)) ? 1 : 0;
                    }
                    zcr = mb_warp_sum(z);
                }
            }
            if (lane == j) {
                stash[0][j] = energy;  // of the rescaled samples when kscale != 0
                stash[1][j] = __int_as_float(zcr);
                stash[17][j] = __int_as_float(kscale);
            }
            phase_sync();

            if (want_spectrum) {
                // Both FFT passes run through ONE copy of the 32-point register FFT (a two-trip loop, not
                // unrolled): the straight-line code of a frame is ~70 KB and instruction fetch matters.
#pragma unroll 1
                for (int pass = 0; pass < 2; pass++) {
                    if (pass == 1) {
                        // ---- pass 2: FFT32 over b for c = lane -> X[lane + 32 d] in v[brev5(d)]
#pragma unroll
                        for (int b = 0; b < 32; b++) v[b] = slot2[lane * kRow + b];
                    }
                    fft32(v);
                    __syncwarp();  // pass 0: all lanes have read their raw samples, the slot becomes the transpose buffer
                    if (pass == 0) {
                        if (!kPcm && want_buffer && lane == 0) bulk_store_wait_read();
                        __syncwarp();
                        float2 t_nx = S.tw32[32 + lane];
#pragma unroll
                        for (int c = 0; c < 32; c++) {
                            float2 y = v[brev5(c)];
                            if (c > 0) {
                                const float2 t = t_nx;
                                if (c + 1 < 32) t_nx = S.tw32[(c + 1) * 32 + lane];
                                y = mbx2::cmul(y, t);
                            }
                            slot2[c * kRow + lane] = y;
                        }
                    } else {
                        // ---- 3. real-FFT split.  X in natural order through the slot so that lane can fetch X[M-k].
#pragma unroll
                        for (int d = 0; d < 32; d++) slot2[lane + 32 * d] = v[brev5(d)];
                        if (lane == 0) slot2[kM] = v[0];  // X[M] := X[0], so that X[M-k] is slot2[kM - k] for every k
                    }
                    phase_sync();
                }

                float av[32];
                const float2 *xm = slot2 + (kM - lane);
                const float2 *twp = S.twN + lane;
                float2 b_nx = xm[0], w_nx = twp[0];  // operands of bin d are fetched during bin d-1
                float lg = 0.f;
                const float sc = P.inv_sqrt_N, hsc = 0.5f * sc;
                // per-lane row pointers: every store below is base + a compile-time offset (32 d floats)
                float *out_re = O.complex_real + g * kN + lane, *out_im = O.complex_imag + g * kN + lane;
                float *mir_re = O.complex_real + g * kN + (kN - lane), *mir_im = O.complex_imag + g * kN + (kN - lane);
                float *out_amp = O.amplitude_spectrum + g * kM + lane, *out_pow = O.power_spectrum + g * kM + lane;
                // The mirrored half Z[N-k] = conj(Z[k]): row d of k = lane + 32 d lands on [N-32d-31, N-32d], which
                // straddles two 128-byte lines.  Lanes >= 1 therefore hold their value back by one row, and
                // store it next to lane 0's current one: [N-32d, N-32d+31] is one aligned line.
                float *mirq_re = mir_re + (lane == 0 ? 0 : 32), *mirq_im = mir_im + (lane == 0 ? 0 : 32);
                float pzr = 0.f, pzi = 0.f;
#pragma unroll
                for (int d = 0; d < 32; d++) {
                    const float2 a = v[brev5(d)];
                    const float2 b = b_nx;  // X[M-k]
                    const float2 w = w_nx;  // h * exp(+2 pi i k / N), h = 0.5 / sqrt(N)
                    if (d + 1 < 32) {
                        b_nx = xm[-32 * (d + 1)];
                        w_nx = twp[32 * (d + 1)];
                    }
                    // E = a + conj(b) = (sx, dy), F = a - conj(b) = (dx, sy); Z = hsc E + sy w + dx (w.y, -w.x): five packed
                    // instructions, each half rounded as in zr = fma(hsc, sx, fma(w.x, sy, w.y dx)), zi = fma(hsc, dy, fma(w.y, sy, -(w.x dx)))
                    const float2 cb = make_float2(b.x, -b.y);
                    const float2 E = mbx2::add(a, cb), F = mbx2::sub(a, cb);
                    const float2 Z = mbx2::fma(E, mbx2::bc(hsc), mbx2::fma(w, mbx2::bc(F.y), mbx2::mul(make_float2(w.y, -w.x), mbx2::bc(F.x))));
                    const float zr = Z.x, zi = Z.y;
                    if (want_cs && exp_store) {
                        // (a rescaled frame, kscale != 0, is brought back to its own units by the fix-up below)
                        st_stream(out_re + 32 * d, zr);
                        st_stream(out_im + 32 * d, zi);
                        if (d == 0) {
                            if (lane == 0) {
                                st_stream(out_re + kM, (a.x - a.y) * sc);         // Nyquist bin (E[0] - O[0]) / sqrt(N)
                                st_stream(out_im + kM, (a.x - a.y) * 0.f + 0.f);  // +0, or NaN when the frame holds a NaN
                            }
                        } else {
                            st_stream(mirq_re - 32 * d, lane == 0 ? zr : pzr);
                            st_stream(mirq_im - 32 * d, -(lane == 0 ? zi : pzi));
                        }
                        pzr = zr;
                        pzi = zi;
                        if (d == 31 && lane != 0) {  // row 31 of lanes >= 1 completes the line the Nyquist bin opened
                            st_stream(mirq_re - 32 * 32, zr);
                            st_stream(mirq_im - 32 * 32, -zi);
                        }
                    }
                    const float amp = sqrt_approx(fmaf(zr, zr, zi * zi));
                    av[d] = amp;
                    if (want_amp_out && exp_store) st_stream(out_amp + 32 * d, amp);
                    if (want_pow_out && exp_store) st_stream(out_pow + 32 * d, __fmul_rn(amp, amp));
                    if (want_log) lg += log2_approx(amp);  // (of the rescaled amplitude: see the chunk finalize)
                }
                if (want_log) {
                    lg = mb_warp_sum(lg);
                    if (lane == j) stash[12][j] = lg;
                }

                if (kscale != 0) {
                    // Rare: the frame was rescaled by 2^kscale.  Bring the stored spectra (each lane re-reads what
                    // it wrote) and the amplitudes used by the band features back to the frame's own units.
#pragma unroll
                    for (int d = 0; d < 32; d++) {
                        av[d] *= unscale;
                        if (want_cs && exp_store) {
                            out_re[32 * d] *= unscale;
                            out_im[32 * d] *= unscale;
                            if (d == 0 && lane == 0) out_re[kM] *= unscale;
                            else { mir_re[-32 * d] *= unscale; mir_im[-32 * d] *= unscale; }
                        }
                        if (want_amp_out && exp_store) out_amp[32 * d] = av[d];
                        if (want_pow_out && exp_store) out_pow[32 * d] = __fmul_rn(av[d], av[d]);
                    }
                }

                if (want_blocked) {
                    // ---- 4. amplitudes into the blocked layout: lane L gets bins [32 L, 32 L + 32)
                    phase_sync();  // every lane has fetched its X[M-k]
#pragma unroll
                    for (int d = 0; d < 32; d++) slot[lane + kAmpStride * d] = av[d];
                    __syncwarp();
                    float ab[32];
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const float4 t = *reinterpret_cast<const float4 *>(slot + kAmpStride * lane + 4 * q);
                        ab[4 * q] = t.x; ab[4 * q + 1] = t.y; ab[4 * q + 2] = t.z; ab[4 * q + 3] = t.w;
                    }
                    // One sequential pass per lane over its 32 bins: running sums of a, p = a^2 and w p (w = bins
                    // since the piece began) since the last boundary, flushed into a piece at every boundary.
                    // Only additions of non-negative terms: a silent band next to a loud bin keeps its value.
                    // (float32 is enough: at most 32 non-negative terms per piece.)  The lane's pieces have
                    // consecutive ids (head first), and the pass is branch-free: boundaries sit at different
                    // positions in every lane, so a branch per boundary would run the flush ~30 times per warp.
                    // A flush is a predicated store, and the restart is folded into the next accumulation
                    // (r * keep + x with keep 0 or 1; r is finite unless the whole frame is NaN anyway).
                    float4 *piece = reinterpret_cast<float4 *>(slot + kPieceOff);  // [kPieces]: {sum a, sum p, sum w p, -}
                    float ra = 0.f, rp = 0.f, rr = 0.f;
                    // k-weights count from the piece's own first bin: a strong bin that opens a mel segment then
                    // weighs exactly 0 there (counted from the lane start it left a rounding residue of
                    // 6e-8 x 17 x its power in a filter that may hold a billion times less)
                    float wk = -1.f;
                    // Float64 on the otherwise idle DP pipe: sum i^q a over the lane's bins, i = 0..31 (exact
                    // products, compile-time weights).  q = 0 is the lane total the rolloff scan needs (a
                    // discrete output); q = 1..4 become the spectral moments (src/utils.js:1-11) after the
                    // shift to k = 32 lane + i below.
                    double ta = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
                    float qn = 0.f, q_lane = 0.f, q4_lane = 0.f;  // sum of q(a) over the lane's looked-at bins (mb_adaptive.cuh)
                    uint32_t paddr = smem_u32(piece + (slot_base + lane));
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
                                t2 = fma(ad, (double)(i * i), t2);
                                t3 = fma(ad, (double)(i * i * i), t3);
                                t4 = fma(ad, (double)(i * i * i * i), t4);
                            }
                        }
                        if (!want_pieces) continue;
                        float keep;  // 0 at a boundary (flush, step to the next piece, restart), else 1
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
                        const float pf = __fmul_rn(ab[i], ab[i]);
                        wk = fmaf(wk, keep, keep);  // bins since the piece began: 0 at a boundary, else one more
                        ra = fmaf(ra, keep, ab[i]);
                        rp = fmaf(rp, keep, pf);
                        rr = fmaf(wk, pf, rr * keep);
                    }
                    if (want_pieces)
                        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(paddr), "f"(ra), "f"(rp), "f"(rr), "f"(0.f)
                                     : "memory");
                    if (want_moments) {
                        // sum (c + i)^p a with c = 32 lane: binomial shift, every term non-negative
                        const double c = (double)(32 * lane), c2 = c * c;
                        const double s1 = fma(c, ta, t1);
                        const double s2 = fma(c2, ta, fma(2.0 * c, t1, t2));
                        const double s3 = fma(c2 * c, ta, fma(3.0 * c2, t1, fma(3.0 * c, t2, t3)));
                        const double s4 = fma(c2 * c2, ta, fma(4.0 * c2 * c, t1, fma(6.0 * c2, t2, fma(4.0 * c, t3, t4))));
                        // four sums reduced together: after the two folding steps every lane carries one of them
                        const double rq = warp_sum4_d(s1, s2, s3, s4, lane);
                        if ((lane & 7) == 0) stash_put_d(stash, 4 + 2 * (lane >> 3), j, rq);  // lanes 0, 8, 16, 24: s1 .. s4
                        if (!want_rolloff) {  // (the rolloff scan below yields the same total)
                            const double r0 = warp_sum_d(ta);
                            if (lane == j) stash_put_d(stash, 2, j, r0);
                        }
                        // Q_0 and Q_4 of mb_adaptive.cuh (k := the last bin of the lane's block)
                        const float ql = 6.4f * qn,  /* (4: every fourth bin was looked at; 1.6: the exponent trick's worst case, squared) */ k4 = (float)(32 * lane + 31) * (float)(32 * lane + 31);
                        q_lane = ql;
                        q4_lane = ql * (k4 * k4);
                    }
                    if (want_rolloff) {
                        // spectralRolloff.js: the largest m with sum_{k<m} a[k] <= 0.99 sum a.  Lane totals are
                        // scanned in double; the one lane the threshold falls into is then scanned bin by
                        // bin by the whole warp (its 32 amplitudes, one per lane).
                        double ia = ta;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const double ya = __shfl_up_sync(0xffffffffu, ia, o);
                            if (lane >= o) ia += ya;
                        }
                        const double total_a = __shfl_sync(0xffffffffu, ia, 31);
                        const double thr = 0.99 * total_a;
                        // lanes whose first bin is still at or under the threshold form a prefix of the warp
                        const uint32_t under = __ballot_sync(0xffffffffu, (ia - ta) <= thr);
                        int rbin = kM;
                        if (total_a > thr && under != 0u) {  // spectralRolloff.js:11-15: the loop runs only while ec > threshold
                            const int lc = 31 - __clz(under);  // last lane starting at or under the threshold
                            const double base = __shfl_sync(0xffffffffu, ia - ta, lc);
                            double x = (double)slot[kAmpStride * lc + lane];  // bin 32 lc + lane
                            double incl = x;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const double y = __shfl_up_sync(0xffffffffu, incl, o);
                                if (lane >= o) incl += y;
                            }
                            const uint32_t ok = __ballot_sync(0xffffffffu, base + (incl - x) <= thr);  // P[32 lc + lane] <= thr
                            rbin = 32 * lc + (31 - __clz(ok));
                        }
                        if (lane == j) {
                            stash[13][j] = __int_as_float(rbin);
                            if (want_moments) stash_put_d(stash, 2, j, total_a);
                        }
                    }
                    phase_sync();
                    float nu_lane = 0.f, ndl_lane = 0.f;  // this lane's share of the noise bounds (mb_adaptive.cuh)
                    // ---- lanes finish the bands: loudness.js:55-63, perceptual*.js
                    if (want_bark) {
                        float sp = 0.f, nu = 0.f;
                        if (lane < MB_NUM_BARK_BANDS) {
                            float bsum = 0.f;
                            for (int it = S.seg_ptr[lane]; it < S.seg_ptr[lane + 1]; it++) bsum += piece[S.seg_items[it]].x;
                            sp = pow023_approx(bsum);
                            if (mb_has(mask, MB_FEAT_LOUDNESS)) O.loudness_specific[g * MB_NUM_BARK_BANDS + lane] = sp;
                            if (kNoise) nu = mb_noise_band(bsum, sp, S.noise_c[0][lane], sigma);
                        }
                        const float total = mb_warp_sum(sp);
                        nu_lane = nu;
                        float mx = (sp > 0.f) ? sp : 0.f;  // NaN never compares greater (perceptualSpread.js:6)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                        const float wsh = (lane >= 1 && lane <= 15) ? (float)lane : 0.f;  // (i+1) * spec[i+1], i < 15
                        const float sharp = mb_warp_sum(wsh * sp);
                        if (lane == j) {
                            stash[14][j] = total;
                            stash[15][j] = mx;
                            stash[16][j] = sharp;
                        }
                    }
                    // ---- mel energies from the pieces, log, DCT (mfcc.js:40-93).  Lane s owns mel segment
                    // [mel[s], mel[s+1]): rising weights (k - mel[s]) / width feed filter s, the
                    // complement feeds filter s - 1.
                    if (want_mfcc) {
                        float rise = 0.f, fall = 0.f;
                        if (lane <= MB_NUM_MEL_FILTERS) {
                            const int seg = MB_NUM_BARK_BANDS + lane, e0 = S.mel_edge[lane];
                            const float inv = S.mel_inv[lane];
                            for (int it = S.seg_ptr[seg]; it < S.seg_ptr[seg + 1]; it++) {
                                const int pc = S.seg_items[it];
                                const float4 pv = piece[pc];
                                const float pp = pv.y;
                                // sum (k - e0) p over the piece (pv.z counts k from the piece's own first bin, at or after e0: non-negative terms only), then its
                                // complement (e1 - k) p: both non-negative up to rounding (NaN stays NaN)
                                float up = fmaf((float)(S.piece_edge[pc] - e0), pp, pv.z) * inv;
                                up = (up < 0.f) ? 0.f : up;
                                rise += up;
                                fall += fmaxf(pp - up, 0.f);
                            }
                        }
                        const float fall_next = __shfl_down_sync(0xffffffffu, fall, 1);
                        const float melE = rise + fall_next;
                        const float lgE = ln_approx(melE);  // lanes >= 26 are not used below
                        if (kNoise && lane < MB_NUM_MEL_FILTERS) ndl_lane = mb_noise_mel(melE, S.noise_c[1][lane], S.noise_c[2][lane], sigma);
                        // 13 x 26 DCT on 26 lanes: lane k + 13 h sums filters 13 h .. 13 h + 12 of coefficient k
                        float acc = 0.f;
                        const int half = lane >= MB_NUM_MFCC ? MB_NUM_MFCC : 0;
#pragma unroll
                        for (int n = 0; n < MB_NUM_MFCC; n++) {
                            const float lf = __shfl_sync(0xffffffffu, lgE, n + half);
                            acc = fmaf(S.dct2[n * 32 + lane], lf, acc);
                        }
                        acc += __shfl_down_sync(0xffffffffu, acc, MB_NUM_MFCC);
                        if (lane < MB_NUM_MFCC) O.mfcc[g * MB_NUM_MFCC + lane] = acc * (1.0f / (float)MB_NUM_MFCC);
                    }
                    if (kNoise) {  // Q_0, Q_4, the bands' bound, the filters' bound: one reduction, lanes 0 / 8 / 16 / 24 hold them
                        const float t4 = mb_warp_sum4(q_lane, q4_lane, nu_lane, ndl_lane, lane);
                        if ((lane & 7) == 0) stash[18 + (lane >> 3)][j] = t4;
                    }
                }
            }
        }  // frames of the chunk

        // ---- 5. one frame per lane: the "number" features of the chunk, coalesced
        __syncwarp();
        bool need_exact = false;
        if (lane < nfc) {
            MbFrameSums F;
            F.energy = ldexp((double)stash[0][lane], -2 * __float_as_int(stash[17][lane]));
            F.zcr = __float_as_int(stash[1][lane]);
            F.s0 = stash_get_d(stash, 2, lane);
            F.s1 = stash_get_d(stash, 4, lane);
            F.s2 = stash_get_d(stash, 6, lane);
            F.s3 = stash_get_d(stash, 8, lane);
            F.s4 = stash_get_d(stash, 10, lane);
            // the log sum was taken on the rescaled amplitudes (a 2^kscale), the moment sums on the frame's own
            F.log2sum = (double)stash[12][lane] - (double)(kM * __float_as_int(stash[17][lane]));
            F.rolloff_bin = __float_as_int(stash[13][lane]);
            const int64_t g = g0 + lane;
            MbMoments MO;
            mb_store_scalars(P, O, g, F, &MO);
            if (want_bark) {
                const double total = (double)stash[14][lane], mx = (double)stash[15][lane];
                const double sharp = (double)stash[16][lane] + P.sharp_const;
                if (mb_has(mask, MB_FEAT_LOUDNESS)) O.loudness_total[g] = (float)total;
                if (mb_has(mask, MB_FEAT_PERCEPTUAL_SPREAD)) {
                    const double r = (total - mx) / total;
                    O.perceptual_spread[g] = (float)(r * r);
                }
                if (mb_has(mask, MB_FEAT_PERCEPTUAL_SHARPNESS)) O.perceptual_sharpness[g] = (float)(sharp * (0.11 / total));
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
                need_exact = __float_as_int(stash[17][lane]) != 0 || mb_noise_needs_exact(P, mask, F, MO, NF);
            }
        }
        mb_noise_append(T, need_exact, g0 + lane);  // frames of this chunk that the exact-FFT kernel redoes
    }
    if (lane == 0) bulk_store_wait_all();  // outstanding `buffer` stores complete before the CTA retires
}

}  // namespace

size_t mb_warp2048_smem_bytes() { return sizeof(Smem) + 128; }

cudaError_t mb_launch_warp2048(const MbDevPlan &P, const MbClipTable &T, const float *samples, const mb_outputs &O,
                               int num_sms, cudaStream_t stream) {
    const size_t smem = mb_warp2048_smem_bytes();
    const bool pcm = T.pcm_channels > 0;
    const uint32_t m = P.mask & MB_ALL_FEATURES;
    constexpr uint32_t kC3 = MB_FEATURE_BIT(MB_FEAT_MFCC) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_CENTROID) |
                             MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SPREAD) | MB_FEATURE_BIT(MB_FEAT_SPECTRAL_SKEWNESS) |
                             MB_FEATURE_BIT(MB_FEAT_SPECTRAL_KURTOSIS);
    constexpr uint32_t kNoArrays = MB_ALL_FEATURES & ~(MB_FEATURE_BIT(MB_FEAT_BUFFER) | MB_FEATURE_BIT(MB_FEAT_COMPLEX_SPECTRUM) |
                                                       MB_FEATURE_BIT(MB_FEAT_AMPLITUDE_SPECTRUM) | MB_FEATURE_BIT(MB_FEAT_POWER_SPECTRUM));
#define MB_PICK(MASK) (pcm ? mb_warp2048_kernel<MASK, true> : mb_warp2048_kernel<MASK, false>)
    auto kernel = m == MB_ALL_FEATURES ? MB_PICK(MB_ALL_FEATURES) : m == kC3 ? MB_PICK(kC3) : m == kNoArrays ? MB_PICK(kNoArrays)
                                                                                                             : MB_PICK(0u);
#undef MB_PICK
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
