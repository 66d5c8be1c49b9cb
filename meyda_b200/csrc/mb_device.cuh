// mb_device.cuh -- device-side plan/clip structures and helpers shared by the
// kernels of the Meyda frame path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/meyda_b200.h"

// Tables of the warp-per-frame kernel (bufferSize 2048), one device copy per plan.
//
// Band sums are pure additions of "pieces".  The union of Bark limits and mel
// edges cuts the N/2 bins into runs; in the blocked layout (lane L owns bins
// [32 L, 32 L + 32)) a lane produces one piece per boundary inside it (bins
// from that boundary to the next one or the lane's end) plus a head piece (its
// bins before its first boundary).  A lane's pieces have consecutive ids, head
// first: head of lane L -> lane_slot_base[L] + L, boundary s (in lane L) ->
// s + L + 1.  A Bark band or a mel segment is a short list of pieces.
#define MB_WARP_MAX_SLOTS 64
#define MB_WARP_HEAD 64
#define MB_WARP_PIECES (MB_WARP_HEAD + 32)
#define MB_WARP_SEGMENTS (MB_NUM_BARK_BANDS + MB_NUM_MEL_FILTERS + 1)  // 24 bands + 27 mel segments
#define MB_WARP_MAX_ITEMS 256
struct MbWarpTables {
    float2 tw32[32 * 32];          // exp(+2 pi i b c / 1024) at [c*32 + b]
    uint32_t lane_bmask[32];       // bit i: bin 32*lane + i is a Bark limit or a mel edge
    int lane_slot_base[32];        // number of such boundaries below bin 32*lane
    int lane_seg_start[32];        // largest boundary <= 32*lane
    int piece_edge[MB_WARP_PIECES];            // the bin a piece's k-weights are measured from (its own first bin)
    int seg_ptr[MB_WARP_SEGMENTS + 1];         // CSR: pieces of band b (0..23) / mel segment s (24 + s)
    unsigned char seg_items[MB_WARP_MAX_ITEMS];
    int n_slots;                   // boundaries below M
};

// Tables of the multi-frame warp kernel (bufferSize 256 / 512 / 1024: F = 2048 / N frames per warp at a time).  Same
// idea as MbWarpTables with the warp's 1024 bins being F frames back to back: lane L = (frame L / A, row L % A)
// owns bins [32 (L % A), +32) of its frame; segments are numbered frame-major (f * MB_WARP_SEGMENTS + s).
#define MB_MF_MAX_PIECES 340  // (256 when a piece is a float4: bufferSize 512 / 1024; three floats at 256)
#define MB_MF_MAX_SEGMENTS (8 * MB_WARP_SEGMENTS)
#define MB_MF_MAX_ITEMS 1024
struct MbWarpMfTables {
    float2 tw32[32 * 32];          // [c'][b]: exp(+2 pi i (c' mod A) b / M)
    uint32_t lane_bmask[32];       // bit i: bin 32 (lane % A) + i of the frame is a Bark limit or a mel edge
    int lane_slot_base[32];        // boundaries in the lanes below (all frames below included)
    short piece_edge[MB_MF_MAX_PIECES];  // the frame-local bin a piece's k-weights are measured from
    short seg_ptr[MB_MF_MAX_SEGMENTS + 1];
    unsigned short seg_items[MB_MF_MAX_ITEMS];
    int n_pieces;                  // > MB_MF_MAX_PIECES: does not fit, the plan keeps the generic kernel
};

// Plan constants of the noise bounds (mb_adaptive.cuh), one device copy per plan.
struct MbNoiseTables {
    float band_c[MB_MAX_BARK_BANDS];    // 2 n_b + K sqrt(n_b), n_b = bins of Bark band b (0: empty band)
    float mel_c1[MB_MAX_MEL_FILTERS];   // 2 K sqrt(W_f / max(W_f, 1)), W_f = total weight of mel filter f
    float mel_c2[MB_MAX_MEL_FILTERS];   // 4 W_f (0: empty filter)
};

// Per-plan constants handed to kernels by value (__grid_constant__).  Tables
// are what `new Meyda(...)` precomputes (src/meyda.js:44-48) plus the ones
// mfcc.js rebuilds per call (src/extractors/mfcc.js:15-83).
struct MbDevPlan {
    int N;              // bufferSize
    int M;              // N/2: complex FFT length == ampSpectrum.length
    int log2M;
    int hop;
    uint32_t mask;      // MB_FEATURE_BIT(...) set
    float inv_sqrt_N;   // unitary scaling: jsfft multiplies by SQRT1_2 per stage (lib/jsfft/fft.js:158-161)
    double sr;
    double slope_freq_sum;      // sum_k f_k         (spectralSlope.js:14)
    double slope_pow_freq_sum;  // sum_k f_k^2       (spectralSlope.js:13)
    double rolloff_bin_hz;      // sr / (2 (n-1))    (spectralRolloff.js:4)
    double sharp_const;         // sum_{i=15..23} 0.066 exp(0.171 (i+1))  (perceptualSharpness.js:10)
    const float *window;        // [N] hanning or hamming (src/meyda.js:116-138)
    const float2 *twM;          // [M/2] exp(+2 pi i j / M)
    const float2 *twN;          // [M]   exp(+2 pi i k / N)
    const float *dct;           // [nc*nf] (13*26) idx = i + j*nc (mfcc.js:72-83)
    const float *mel_inv_width; // [nf+1] (27) 1 / (mel[s+1] - mel[s]) (0 if empty)
    const double2 *tw_exact;    // [N-1] jsfft recurrence twiddles, stage of width w at [w-1, 2w-1) (exact mode)
    int exact;                  // MB_FLAG_EXACT_FFT
    const double *mel_w_exact;  // exact mode: filter f's weights for bins mel[f] .. mel[f+2]-1, (i-lo)/(hi-lo) as doubles
    int mel_w_off[MB_MAX_MEL_FILTERS + 1];  // offsets of each filter's run in mel_w_exact
    const MbWarpTables *warp_tables;  // bufferSize 2048 only, else NULL
    const MbWarpMfTables *warp_mf_tables;  // bufferSize 256 / 512 / 1024 only, else NULL
    int bb[MB_MAX_BARK_BANDS + 1];     // loudness.js:24-45: nb + 1 limits
    int mel[MB_MAX_MEL_FILTERS + 2];   // mfcc.js:31-38: nf + 2 bins
    // what the reference keeps as constants (mb_plan_create_ex); the warp kernels only ever see 24 / 26 / 13 / 0.99
    // mb_adaptive.cuh: kap sqrt(sum_k k^(2p)), p = 0..4, and the per-band / per-filter constants of the noise bounds
    double noise_sqrtT[5];
    const struct MbNoiseTables *noise;
    int nb;               // Bark bands   (loudness.js:14)
    int nf;               // mel filters  (mfcc.js:15)
    int nc;               // coefficients (mfcc.js:71)
    double rolloff_frac;  // spectralRolloff.js:9
};

// Clip list of one extract call (device arrays).
struct MbClipTable {
    const int64_t *clip_off;     // [n_clips] first sample (per channel) of each clip
    const int64_t *frame_start;  // [n_clips + 1] exclusive prefix of frames per clip
    int64_t n_clips;
    int64_t total_frames;
    // 0: `samples` is float32 mono.  > 0: `samples` is what a WAV data chunk holds, this many interleaved
    // channels of `pcm_format` samples, of which channel `pcm_channel` is taken and converted on load as
    // decodeAudioData does (int16 / 32768, int24 / 8388608, float32 as is): the step lib/bufferLoader.js:13-44 +
    // getChannelData(0), src/meyda.js:72, perform in the reference.
    int pcm_channels;
    int pcm_channel;
    int pcm_format;  // MB_SAMPLE_S16 / MB_SAMPLE_S24 / MB_SAMPLE_F32 (meaningful when pcm_channels > 0)
    // Adaptive exactness (mb_adaptive.cuh).  A float32-FFT kernel appends the frames whose features sit in the
    // reference's own rounding noise to fix_list (fix_count: how many; NULL: do not).  An exact-FFT kernel given
    // sel_list works through frames sel_list[0 .. *sel_count) instead of [0, total_frames).
    int *fix_count;
    int *fix_list;
    const int *sel_count;
    const int *sel_list;
};

// One frame's samples, whatever the storage.
struct MbFrameSrc {
    const unsigned char *base;  // first sample of the frame (of the chosen channel)
    int format;                 // -1: float32 mono
    int stride;                 // bytes between consecutive samples of the channel
    __device__ __forceinline__ float operator[](int i) const {
        const unsigned char *p = base + (int64_t)i * stride;
        if (format < 0) return __ldg(reinterpret_cast<const float *>(base) + i);
        if (format == MB_SAMPLE_S16) return (float)__ldg(reinterpret_cast<const int16_t *>(p)) * (1.0f / 32768.0f);
        if (format == MB_SAMPLE_S24) {  // little-endian packed 3 bytes, sign in the last one
            const int v = (int)__ldg(p) | ((int)__ldg(p + 1) << 8) | ((int)(signed char)__ldg(p + 2) << 16);
            return (float)v * (1.0f / 8388608.0f);
        }
        return __ldg(reinterpret_cast<const float *>(p));
    }
};
__host__ __device__ __forceinline__ int mb_sample_bytes(int format) {
    return format == MB_SAMPLE_S16 ? 2 : format == MB_SAMPLE_S24 ? 3 : 4;
}
__device__ __forceinline__ MbFrameSrc mb_frame_src(const MbClipTable &T, const float *samples, int64_t first) {
    MbFrameSrc r;
    if (T.pcm_channels > 0) {
        const int sb = mb_sample_bytes(T.pcm_format);
        r.base = reinterpret_cast<const unsigned char *>(samples) + (first * T.pcm_channels + T.pcm_channel) * sb;
        r.format = T.pcm_format;
        r.stride = T.pcm_channels * sb;
    } else {
        r.base = reinterpret_cast<const unsigned char *>(samples + first);
        r.format = -1;
        r.stride = 4;
    }
    return r;
}

__host__ __device__ __forceinline__ bool mb_has(uint32_t mask, int f) { return (mask >> f) & 1u; }

// Largest c with frame_start[c] <= g (g < total_frames).
__device__ __forceinline__ int64_t mb_find_clip(const MbClipTable &T, int64_t g) {
    int64_t lo = 0, hi = T.n_clips;  // invariant: frame_start[lo] <= g < frame_start[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if (T.frame_start[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// The same search done by a whole (converged) warp for a warp-uniform g: 32 probes per round trip, so two rounds for
// 1,024 clips and three for 20,000 where the scalar search walks 10-15 dependent loads (3 % of the stall samples of
// the bufferSize-32768 kernel sat on that chain).
__device__ __forceinline__ int64_t mb_find_clip_warp(const MbClipTable &T, int64_t g) {
    const int lane = threadIdx.x & 31;
    int64_t lo = 0, hi = T.n_clips;  // invariant: frame_start[lo] <= g < frame_start[hi]
    while (hi - lo > 1) {
        const int64_t step = (hi - lo + 31) >> 5;
        const int64_t idx = lo + (int64_t)lane * step;  // lane 0 probes lo itself: always true
        const bool le = idx < hi && __ldg(T.frame_start + idx) <= g;
        const unsigned m = __ballot_sync(0xffffffffu, le) | 1u;  // monotone: a prefix of the lanes
        const int top = 31 - __clz(m);
        lo += (int64_t)top * step;
        hi = min(hi, lo + step);
    }
    return lo;
}

__device__ __forceinline__ double mb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float mb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int mb_warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Scalar features from the per-frame sums; shared by every kernel variant so
// that the formulas (and their quirks) live in one place.
struct MbFrameSums {
    double s0, s1, s2, s3, s4;  // sum_k k^p a[k]            (src/utils.js:1-11)
    double log2sum;             // sum_k log2 a[k]           (spectralFlatness.js:6)
    double energy;              // sum x^2 over the raw frame (energy.js, rms.js)
    int zcr;                    // zcr.js
    int rolloff_bin;            // (n + 1) of spectralRolloff.js:15
};

// What mb_store_scalars derived on the way (mb_adaptive.cuh bounds its error terms from these).
struct MbMoments {
    double m1, m2, m3, m4;  // S_i / S_0
    double var, sd;         // m2 - m1^2 and its root
    double flatness;
};

__device__ __forceinline__ void mb_store_scalars(const MbDevPlan &P, const mb_outputs &O, int64_t g,
                                                 const MbFrameSums &S, MbMoments *mo = nullptr) {
    const uint32_t mask = P.mask;
    const double n = (double)P.M;
    if (mb_has(mask, MB_FEAT_RMS)) O.rms[g] = (float)sqrt(S.energy / (double)P.N);
    if (mb_has(mask, MB_FEAT_ENERGY)) O.energy[g] = (float)S.energy;
    if (mb_has(mask, MB_FEAT_ZCR)) O.zcr[g] = S.zcr;
    const double m1 = S.s1 / S.s0, m2 = S.s2 / S.s0, m3 = S.s3 / S.s0, m4 = S.s4 / S.s0;
    if (mb_has(mask, MB_FEAT_SPECTRAL_CENTROID)) O.spectral_centroid[g] = (float)m1;
    const double var = m2 - m1 * m1;
    const double sd = sqrt(var);
    if (mb_has(mask, MB_FEAT_SPECTRAL_SPREAD)) O.spectral_spread[g] = (float)sd;
    if (mb_has(mask, MB_FEAT_SPECTRAL_SKEWNESS))  // spectralSkewness.js:6-8
        O.spectral_skewness[g] = (float)((2 * m1 * m1 * m1 - 3 * m1 * m2 + m3) / (sd * sd * sd));
    if (mb_has(mask, MB_FEAT_SPECTRAL_KURTOSIS))  // spectralKurtosis.js:7-9: 6*m1*m2, not 6*m1^2*m2
        O.spectral_kurtosis[g] =
            (float)((-3 * m1 * m1 * m1 * m1 + 6 * m1 * m2 - 4 * m1 * m3 + m4) / (sd * sd * sd * sd));
    double flat = 0.0;
    if (mb_has(mask, MB_FEAT_SPECTRAL_FLATNESS)) {  // (exp(mean ln a) * n) / sum a
        flat = exp(S.log2sum * 0.6931471805599453 / n) * n / S.s0;
        O.spectral_flatness[g] = (float)flat;
    }
    if (mo) {
        mo->m1 = m1; mo->m2 = m2; mo->m3 = m3; mo->m4 = m4;
        mo->var = var; mo->sd = sd;
        mo->flatness = flat;
    }
    if (mb_has(mask, MB_FEAT_SPECTRAL_SLOPE)) {   // spectralSlope.js:17; sum f a = (sr/N) s1
        const double amp_freq_sum = S.s1 * (P.sr / (double)P.N);
        O.spectral_slope[g] = (float)((n * amp_freq_sum - P.slope_freq_sum * S.s0) /
                                      (S.s0 * (P.slope_pow_freq_sum - P.slope_freq_sum * P.slope_freq_sum)));
    }
    if (mb_has(mask, MB_FEAT_SPECTRAL_ROLLOFF)) O.spectral_rolloff[g] = (float)((double)S.rolloff_bin * P.rolloff_bin_hz);
}
