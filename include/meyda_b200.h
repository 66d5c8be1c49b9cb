/*
 * meyda_b200.h -- C ABI of the B200-native Meyda frame-feature path.
 *
 * This is the drop-in boundary: a Node.js N-API addon (js/addon.cc), the
 * Python ctypes binding (meyda_b200/_capi.py) and any other FFI bind exactly
 * these entry points.  Plain pointers and sizes only; nothing throws across
 * the boundary; every function returns an mb_status and mb_last_error() holds
 * the message of the calling thread's last failure.
 *
 * What each entry point replaces in the reference (kirbysayshi/meyda v1.1.0;
 * the reference has no FFI of its own, so these are the seams its JavaScript
 * would call through):
 *
 *   mb_plan_create   <- `new Meyda(audioContext, src, bufSize, callback)`
 *                       src/meyda.js:17-65: power-of-two check (:20-22), bark
 *                       scale (:44,170-182), hanning/hamming tables (:47-48,
 *                       116-138), Loudness bark-band limits
 *                       (src/extractors/loudness.js:24-45), and the mel/DCT
 *                       tables mfcc rebuilds on every call
 *                       (src/extractors/mfcc.js:15-83).
 *   mb_extract       <- the per-buffer pipeline `onaudioprocess`
 *                       src/meyda.js:69-91 (window :158-168, FFT
 *                       lib/jsfft/fft.js:123-208, amplitude :104-114) followed
 *                       by `get([...features])` src/meyda.js:244-261 over
 *                       the extractor files under src/extractors/, for every frame of every clip.
 *   mb_query_output  <- the implicit result shapes of src/feature-info.js:3-64.
 *   mb_stream_*      <- the stateful buffer-by-buffer use of the same
 *                       pipeline (src/meyda.js:69-91, start/stop :233-241).
 *   mb_extract_pcm16 <- the same pipeline fed with the 16-bit PCM a WAV file
 *   mb_wav_parse        holds instead of decoded float32: replaces
 *                       lib/bufferLoader.js:13-44 (XHR + decodeAudioData, which
 *                       turns int16 s into s / 32768) and the channel pick
 *                       `getChannelData(0)` of src/meyda.js:72; the conversion
 *                       happens inside the framing load of the kernels.
 *
 * Framing rule (the reference has none: ScriptProcessor hands over back-to-back
 * buffers, i.e. hop == bufferSize): frame f of a clip covers samples
 * [f*hop, f*hop + bufferSize); frames = len < bufferSize ? 0 :
 * (len - bufferSize) / hop + 1; no padding, trailing partial buffer dropped.
 *
 * Output layout: struct-of-arrays per feature, frame-major, clips concatenated
 * in input order.  "number" features are float32[totalFrames] (zcr is
 * int32[totalFrames]); array features are float32[totalFrames][len].
 */
#ifndef MEYDA_B200_H
#define MEYDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_VERSION 100 /* 0.1.0 */

typedef int mb_status;
enum {
    MB_OK = 0,
    MB_ERR_INVALID_ARG = 1,      /* NULL pointer, negative size, unknown enum */
    MB_ERR_NOT_POWER_OF_TWO = 2, /* "Buffer size is not a power of two: Meyda will not run." */
    MB_ERR_UNSUPPORTED = 3,      /* bufferSize outside [MB_MIN_BUFFER_SIZE, MB_MAX_BUFFER_SIZE] */
    MB_ERR_CUDA = 4,             /* a CUDA runtime call failed; message has the CUDA error */
    MB_ERR_NO_DEVICE = 5,        /* no CUDA device / device index out of range */
    MB_ERR_MISSING_OUTPUT = 6,   /* a requested feature's output pointer is NULL */
    MB_ERR_OUT_OF_RANGE = 7      /* clip offset/length outside the sample array */
};

#define MB_MIN_BUFFER_SIZE 16
#define MB_MAX_BUFFER_SIZE 32768

/* Feature bits, in the key order of src/feature-info.js:3-64. */
enum {
    MB_FEAT_BUFFER = 0,
    MB_FEAT_RMS = 1,
    MB_FEAT_ENERGY = 2,
    MB_FEAT_ZCR = 3,
    MB_FEAT_COMPLEX_SPECTRUM = 4,
    MB_FEAT_AMPLITUDE_SPECTRUM = 5,
    MB_FEAT_POWER_SPECTRUM = 6,
    MB_FEAT_SPECTRAL_CENTROID = 7,
    MB_FEAT_SPECTRAL_FLATNESS = 8,
    MB_FEAT_SPECTRAL_SLOPE = 9,
    MB_FEAT_SPECTRAL_ROLLOFF = 10,
    MB_FEAT_SPECTRAL_SPREAD = 11,
    MB_FEAT_SPECTRAL_SKEWNESS = 12,
    MB_FEAT_SPECTRAL_KURTOSIS = 13,
    MB_FEAT_LOUDNESS = 14,
    MB_FEAT_PERCEPTUAL_SPREAD = 15,
    MB_FEAT_PERCEPTUAL_SHARPNESS = 16,
    MB_FEAT_MFCC = 17,
    MB_NUM_FEATURES = 18
};
#define MB_FEATURE_BIT(f) (1u << (f))
#define MB_ALL_FEATURES ((1u << MB_NUM_FEATURES) - 1u)

#define MB_NUM_BARK_BANDS 24 /* src/meyda.js:214 */
#define MB_NUM_MEL_FILTERS 26 /* src/extractors/mfcc.js:15 */
#define MB_NUM_MFCC 13        /* src/extractors/mfcc.js:71 */

/* The constants above are what the reference hard-codes; a plan may be created with others (mb_plan_create_ex).
 * NUM_BARK_BANDS is an option of the reference's Loudness constructor (src/extractors/loudness.js:14); the mel
 * filter count (mfcc.js:15), the coefficient count (mfcc.js:71) and the rolloff fraction (spectralRolloff.js:9)
 * are local constants there.  A field left 0 takes the reference's value.  Plans with non-reference values run
 * on the generic kernels (any bufferSize, float32 or exact FFT); output rows then hold num_bark_bands /
 * num_mfcc floats instead of 24 / 13. */
#define MB_MAX_BARK_BANDS 64
#define MB_MAX_MEL_FILTERS 128
typedef struct mb_params {
    int32_t num_bark_bands;   /* 1 .. MB_MAX_BARK_BANDS; 0 = 24 */
    int32_t num_mel_filters;  /* 1 .. MB_MAX_MEL_FILTERS; 0 = 26 */
    int32_t num_mfcc;         /* 1 .. num_mel_filters; 0 = 13 */
    int32_t reserved;         /* must be 0 */
    double rolloff_fraction;  /* in (0, 1]; 0 = 0.99 */
} mb_params;

/* `windowingFunction`, src/meyda.js:41, docs.md:5-11.  Blackman is the window the reference leaves commented
 * out as unfinished (src/meyda.js:140-156); its stated formula, 0.42 - 0.5 cos(2 pi i/(N-1)) + 0.08 cos(4 pi i/(N-1)). */
enum { MB_WINDOW_HANNING = 0, MB_WINDOW_HAMMING = 1, MB_WINDOW_BLACKMAN = 2 };

/* Where the caller's sample and output pointers live. */
enum { MB_MEM_HOST = 0, MB_MEM_DEVICE = 1 };

/* Plan flags. */
enum {
    MB_FLAG_DEFAULT = 0,
    /* Force the generic block-per-frame kernel even where a tuned one exists
     * (testing / A-B comparison). */
    MB_FLAG_GENERIC_KERNEL = 1u << 0,
    /* Reproduce the reference FFT's arithmetic exactly (lib/jsfft/fft.js:123-171:
     * radix-2, float64 butterflies with the recurrence twiddles, float32 store
     * per stage): complexSpectrum / amplitudeSpectrum / powerSpectrum come out
     * bit for bit and every derived feature follows.  Slower than the default
     * float32 FFT.  Above 16384 samples the N-point complex frame no longer fits
     * one CTA and a 2-CTA thread-block cluster holds it (DSMEM exchange). */
    MB_FLAG_EXACT_FFT = 1u << 1,
    /* With MB_FLAG_EXACT_FFT: use the 2-CTA cluster kernel at every bufferSize >= 64
     * (it is automatic above 16384). */
    MB_FLAG_CLUSTER_FFT = 1u << 2,
    /* Float32 FFT only.  By default a float32-FFT plan is ADAPTIVE: the kernels bound, per frame, how far each
     * requested feature can move under FFT rounding noise, and the frames whose values the reference's own
     * per-stage float32 rounding decides (near-pure tones, silent bands: x^0.23, ln x and k^3 / k^4 weights
     * amplify the noise floor without bound) are redone with the MB_FLAG_EXACT_FFT arithmetic, so that every
     * feature lands within 1e-3 of the reference.  Ordinary (noisy) audio flags nothing.  This flag turns the
     * second pass off (A/B measurements). */
    MB_FLAG_NO_REFINE = 1u << 3
};
#define MB_MAX_EXACT_BUFFER_SIZE 32768

/* A plan is NOT thread-safe and runs on ONE stream at a time: its clip table, flagged-frame list and staging
 * buffers are ordered by that stream alone.  Use one plan per thread; mb_plan_set_stream synchronizes the stream it
 * leaves before switching.  (mb_extract_multi drives one plan per device from its own threads.) */
typedef struct mb_plan mb_plan;     /* opaque: tables, stream, scratch of one (device, bufferSize, hop, ...) */
typedef struct mb_stream mb_stream; /* opaque: stateful buffer-by-buffer extractor */

/* Caller-allocated outputs.  A pointer may be NULL iff its feature is not in
 * the plan's mask.  Element counts per frame are given by mb_layout. */
typedef struct mb_outputs {
    float *buffer;               /* [frames][N]   raw signal (docs.md:19-21) */
    float *rms;                  /* [frames]      src/extractors/rms.js */
    float *energy;               /* [frames]      energy.js */
    int32_t *zcr;                /* [frames]      zcr.js (integer, bit-exact) */
    float *complex_real;         /* [frames][N]   complexSpectrum.js -> .real */
    float *complex_imag;         /* [frames][N]                      -> .imag */
    float *amplitude_spectrum;   /* [frames][N/2] amplitudeSpectrum.js */
    float *power_spectrum;       /* [frames][N/2] powerSpectrum.js */
    float *spectral_centroid;    /* [frames]      spectralCentroid.js (bins) */
    float *spectral_flatness;    /* [frames]      spectralFlatness.js */
    float *spectral_slope;       /* [frames]      spectralSlope.js */
    float *spectral_rolloff;     /* [frames]      spectralRolloff.js (Hz) */
    float *spectral_spread;      /* [frames]      spectralSpread.js */
    float *spectral_skewness;    /* [frames]      spectralSkewness.js */
    float *spectral_kurtosis;    /* [frames]      spectralKurtosis.js */
    float *loudness_specific;    /* [frames][24]  loudness.js -> .specific (24 = the plan's num_bark_bands) */
    float *loudness_total;       /* [frames]                  -> .total */
    float *perceptual_spread;    /* [frames]      perceptualSpread.js */
    float *perceptual_sharpness; /* [frames]      perceptualSharpness.js */
    float *mfcc;                 /* [frames][13]  mfcc.js (13 = the plan's num_mfcc) */
} mb_outputs;

/* Result shapes for a given clip list (what `get([...])` would have produced). */
typedef struct mb_layout {
    int64_t total_frames;
    int32_t buffer_size;      /* N */
    int32_t spectrum_size;    /* N/2 */
    uint32_t feature_mask;
    int32_t reserved;
    int64_t output_bytes;     /* sum over requested outputs, all frames */
    int64_t bytes_per_frame;  /* requested output bytes per frame */
    int32_t num_bark_bands;   /* floats per frame of loudness_specific (24 unless mb_plan_create_ex said otherwise) */
    int32_t num_mfcc;         /* floats per frame of mfcc (13 ...) */
} mb_layout;

int mb_version(void);
const char *mb_last_error(void);
const char *mb_feature_name(int feature);   /* "rms", "spectralCentroid", ...; NULL if out of range */
int mb_feature_from_name(const char *name); /* -1 if unknown */
mb_status mb_device_count(int *count);

/* Frames of one clip under the framing rule above. */
int64_t mb_num_frames(int64_t clip_len, int buffer_size, int hop);

mb_status mb_plan_create(mb_plan **plan, int device, int buffer_size, int hop, double sample_rate,
                         int window, uint32_t feature_mask, uint32_t flags);
/* The same with the parameters the reference keeps as constants; params == NULL is mb_plan_create. */
mb_status mb_plan_create_ex(mb_plan **plan, int device, int buffer_size, int hop, double sample_rate,
                            int window, uint32_t feature_mask, uint32_t flags, const mb_params *params);
/* The parameters a plan runs with (defaults filled in). */
mb_status mb_plan_get_params(const mb_plan *plan, mb_params *params);
void mb_plan_destroy(mb_plan *plan);

/* Launch on this CUDA stream (a cudaStream_t) instead of the plan's own;
 * NULL restores the plan's stream.  Lets a caller time with its own events. */
mb_status mb_plan_set_stream(mb_plan *plan, void *cuda_stream);

/* Read back plan tables (host copies) for inspection/tests.  Any pointer may
 * be NULL.  window: N floats; bb_limits: num_bark_bands + 1 ints (25); mel_bins: num_mel_filters + 2 ints (28). */
mb_status mb_plan_tables(const mb_plan *plan, float *window, int32_t *bb_limits, int32_t *mel_bins);

mb_status mb_query_output(const mb_plan *plan, int64_t n_clips, const int64_t *clip_len,
                          int64_t *frames_per_clip /* [n_clips] or NULL */, mb_layout *layout);

/*
 * Extract every frame of every clip.  clip c is samples[clip_offset[c] ..
 * clip_offset[c] + clip_len[c]).  clip_offset/clip_len are HOST arrays in
 * both memory kinds.  Blocking: returns after the results are in `out`.
 *   MB_MEM_HOST:   samples/out are host pointers (pinned memory from
 *                  mb_host_alloc makes the copies asynchronous and overlapped).
 *                  The rows of `buffer` (the caller's own samples, framed) are
 *                  filled on the host while the device works; they never
 *                  cross PCIe.
 *   MB_MEM_DEVICE: samples/out are device pointers on the plan's device.
 * n_samples is the length of `samples` in floats (bounds check).
 */
mb_status mb_extract(mb_plan *plan, const float *samples, int64_t n_samples, const int64_t *clip_offset,
                     const int64_t *clip_len, int64_t n_clips, const mb_outputs *out, int mem_kind);

/* Same, MB_MEM_DEVICE only, without the final stream synchronize (the caller
 * synchronizes or records events on the stream it set). */
mb_status mb_extract_async(mb_plan *plan, const float *samples, int64_t n_samples, const int64_t *clip_offset,
                           const int64_t *clip_len, int64_t n_clips, const mb_outputs *out);
mb_status mb_plan_synchronize(mb_plan *plan);

/* Clip-sharded extraction over several devices from HOST memory: contiguous
 * clip ranges balanced on frame count, one plan per device (all created with
 * the same parameters), no inter-GPU traffic.  Results land in `out` exactly
 * as a single-device call would put them. */
mb_status mb_extract_multi(mb_plan *const *plans, int n_plans, const float *samples, int64_t n_samples,
                           const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                           const mb_outputs *out);

/*
 * 16-bit PCM input (what the reference's fixtures audio/sound1-3.wav hold).  `pcm` is
 * n_sample_frames x channels interleaved int16; channel `channel` is used and
 * every sample becomes (float)s / 32768 on load (Web Audio decodeAudioData),
 * so the results equal mb_extract on the converted samples bit for bit while
 * the device reads half the bytes.  clip_offset / clip_len count sample frames.
 * mem_kind as in mb_extract.  The _async form is MB_MEM_DEVICE only.
 */
enum { MB_SAMPLE_S16 = 1, MB_SAMPLE_S24 = 2, MB_SAMPLE_F32 = 3 };
/* The general form: `data` holds n_sample_frames x channels interleaved samples of `sample_format` (16-bit,
 * packed little-endian 24-bit, or 32-bit float -- the payloads of WAV formats 1 and 3); int24 becomes s / 8388608.
 * mb_extract_pcm16 is this call with MB_SAMPLE_S16 (the only format with a tuned bufferSize-2048 path). */
mb_status mb_extract_pcm(mb_plan *plan, const void *data, int sample_format, int64_t n_sample_frames, int channels,
                         int channel, const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                         const mb_outputs *out, int mem_kind);
mb_status mb_extract_pcm16(mb_plan *plan, const int16_t *pcm, int64_t n_sample_frames, int channels, int channel,
                           const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                           const mb_outputs *out, int mem_kind);
mb_status mb_extract_pcm16_async(mb_plan *plan, const int16_t *pcm, int64_t n_sample_frames, int channels,
                                 int channel, const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                                 const mb_outputs *out);

/* RIFF/WAVE header of an in-memory file: where the sample data lies and how to read it. */
typedef struct mb_wav_info {
    int32_t format;          /* 1 = integer PCM, 3 = IEEE float */
    int32_t channels;
    int32_t sample_rate;
    int32_t bits_per_sample;
    int64_t data_offset;     /* byte offset of the first sample frame */
    int64_t n_sample_frames; /* per channel */
} mb_wav_info;
mb_status mb_wav_parse(const void *file_bytes, int64_t n_bytes, mb_wav_info *info);

/* Measured non-tensor arithmetic peaks of a device, TFLOP/s (an FFMA / DFMA micro-kernel, a few milliseconds): the
 * denominators of the FP32 / FP64 rooflines bench.py reports for the compute-bound configurations.  Either pointer
 * may be NULL. */
mb_status mb_measure_peaks(int device, double *fp32_ffma_tflops, double *fp64_dfma_tflops);

/* Host threads that a host-memory call (MB_MEM_HOST) uses to write the rows the device does not produce: `buffer`
 * (the caller's own samples, framed) and powerSpectrum (amplitudeSpectrum squared).  Process-wide; 0 (the default)
 * picks min(8, cores / 2).  Calls already running keep their count. */
mb_status mb_set_host_threads(int n);

/* Which rows a host-memory call produces on the host while the device works, instead of copying them back:
 *   1  `buffer` (the caller's own samples, framed) and powerSpectrum (amplitudeSpectrum squared): 12 KB of 33 KB per
 *      frame of PCIe traffic at bufferSize 2048 -- 2.02 vs 1.49 M frames/s end to end with one device per host,
 *      3.33 vs 2.82 M with eight (profiles/README.md);
 *   2  also the upper half of complexSpectrum, Z[N-k] = conj(Z[k]) (8 KB more per frame; 2.55 vs 2.12 M frames/s with
 *      one device on a 16-core host): the rows of frames that the exact-FFT kernel redid (whose upper half is computed,
 *      as the reference's is) still come from the device, exact-FFT plans and pageable output arrays copy everything;
 *   0  nothing: the device produces every row and all are copied back (hosts short of cores);
 *  -1  (the default) 2 where one device is visible and the host has twelve or more cores, else 1: mode 2 loads the
 *      host's memory system, which several devices on one host already saturate in mode 1 (two B200s: 3.84 M frames/s
 *      in mode 1, 2.80 M in mode 2).
 * The results are bit-identical in every mode.  Process-wide. */
mb_status mb_set_host_rows(int mode);
/* The mode in force: 0, 1 or 2 (the automatic choice resolved). */
int mb_get_host_rows(void);

/* Number of kernel launches issued by this plan so far (bench evidence). */
int64_t mb_plan_launch_count(const mb_plan *plan);
/* How many frames of the plan's last extract call (or stream push) were redone with the exact-FFT arithmetic
 * (adaptive plans; 0 otherwise).  Synchronizes the plan's stream after a device-memory call. */
mb_status mb_plan_refined_frames(mb_plan *plan, int64_t *frames);
/* Name of the kernel variant the plan dispatches to ("generic", "warp2048", ...). */
const char *mb_plan_kernel_name(const mb_plan *plan);

/* Pinned host memory for samples/outputs of MB_MEM_HOST calls. */
mb_status mb_host_alloc(void **ptr, size_t bytes);
void mb_host_free(void *ptr);

/*
 * Buffer-by-buffer (streaming) use, mirroring onaudioprocess: push any number
 * of new samples, get back the features of every frame completed by them
 * (0 or more).  The hop-overlap tail stays on the device between calls.
 * `out` arrays must have room for mb_stream_frames_after(n_new) frames.
 */
mb_status mb_stream_create(mb_stream **stream, mb_plan *plan);
void mb_stream_destroy(mb_stream *stream);
int64_t mb_stream_frames_after(const mb_stream *stream, int64_t n_new_samples);
mb_status mb_stream_push(mb_stream *stream, const float *new_samples, int64_t n_new_samples,
                         const mb_outputs *out, int mem_kind, int64_t *frames_done);
mb_status mb_stream_reset(mb_stream *stream);
/* The same for 16-bit PCM blocks (what a capture device or a WAV reader hands over): the stream keeps int16 sample
 * frames (`channels` interleaved) on the device and the kernels convert on load, as mb_extract_pcm16 does.
 * n_new_sample_frames counts frames of `channels` samples. */
mb_status mb_stream_create_pcm16(mb_stream **stream, mb_plan *plan, int channels, int channel);
mb_status mb_stream_push_pcm16(mb_stream *stream, const int16_t *new_sample_frames, int64_t n_new_sample_frames,
                               const mb_outputs *out, int mem_kind, int64_t *frames_done);
/* How many pushes were replayed as a CUDA graph so far (MB_MEM_HOST pushes of a repeating shape). */
int64_t mb_stream_graph_launches(const mb_stream *stream);

#ifdef __cplusplus
}
#endif
#endif /* MEYDA_B200_H */
