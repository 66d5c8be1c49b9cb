// capi.cu -- the C ABI of include/meyda_b200.h: plan construction (the tables
// `new Meyda(...)` precomputes), clip/frame bookkeeping, host<->device staging
// and kernel dispatch.  No CPU compute fallback exists: every feature value
// comes out of a CUDA kernel or the call fails.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "mb_adaptive.cuh"
#include "mb_kernels.h"

namespace {

thread_local std::string g_last_error;

mb_status fail(mb_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define MB_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) return fail(MB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
    } while (0)

const char *const kFeatureNames[MB_NUM_FEATURES] = {
    "buffer", "rms", "energy", "zcr", "complexSpectrum", "amplitudeSpectrum", "powerSpectrum",
    "spectralCentroid", "spectralFlatness", "spectralSlope", "spectralRolloff", "spectralSpread",
    "spectralSkewness", "spectralKurtosis", "loudness", "perceptualSpread", "perceptualSharpness", "mfcc"};

// One output array of mb_outputs: which feature owns it and its floats per frame.
struct OutField {
    size_t offset;  // byte offset of the pointer inside mb_outputs
    int feature;
    int kind;  // 0: 1, 1: N, 2: N/2, 3: the plan's Bark bands (24), 4: its mfcc coefficients (13)
};
#define MB_FIELD(name, feat, kind) {offsetof(mb_outputs, name), feat, kind}
const OutField kFields[] = {
    MB_FIELD(buffer, MB_FEAT_BUFFER, 1),
    MB_FIELD(rms, MB_FEAT_RMS, 0),
    MB_FIELD(energy, MB_FEAT_ENERGY, 0),
    MB_FIELD(zcr, MB_FEAT_ZCR, 0),
    MB_FIELD(complex_real, MB_FEAT_COMPLEX_SPECTRUM, 1),
    MB_FIELD(complex_imag, MB_FEAT_COMPLEX_SPECTRUM, 1),
    MB_FIELD(amplitude_spectrum, MB_FEAT_AMPLITUDE_SPECTRUM, 2),
    MB_FIELD(power_spectrum, MB_FEAT_POWER_SPECTRUM, 2),
    MB_FIELD(spectral_centroid, MB_FEAT_SPECTRAL_CENTROID, 0),
    MB_FIELD(spectral_flatness, MB_FEAT_SPECTRAL_FLATNESS, 0),
    MB_FIELD(spectral_slope, MB_FEAT_SPECTRAL_SLOPE, 0),
    MB_FIELD(spectral_rolloff, MB_FEAT_SPECTRAL_ROLLOFF, 0),
    MB_FIELD(spectral_spread, MB_FEAT_SPECTRAL_SPREAD, 0),
    MB_FIELD(spectral_skewness, MB_FEAT_SPECTRAL_SKEWNESS, 0),
    MB_FIELD(spectral_kurtosis, MB_FEAT_SPECTRAL_KURTOSIS, 0),
    MB_FIELD(loudness_specific, MB_FEAT_LOUDNESS, 3),
    MB_FIELD(loudness_total, MB_FEAT_LOUDNESS, 0),
    MB_FIELD(perceptual_spread, MB_FEAT_PERCEPTUAL_SPREAD, 0),
    MB_FIELD(perceptual_sharpness, MB_FEAT_PERCEPTUAL_SHARPNESS, 0),
    MB_FIELD(mfcc, MB_FEAT_MFCC, 4),
};
constexpr int kNumFields = sizeof(kFields) / sizeof(kFields[0]);

int field_elems(const OutField &f, const MbDevPlan &D) {
    switch (f.kind) {
        case 0: return 1;
        case 1: return D.N;
        case 2: return D.N / 2;
        case 3: return D.nb;
        default: return D.nc;
    }
}
void *&field_ptr(mb_outputs &o, const OutField &f) { return *reinterpret_cast<void **>(reinterpret_cast<char *>(&o) + f.offset); }
void *field_ptr(const mb_outputs &o, const OutField &f) {
    return *reinterpret_cast<void *const *>(reinterpret_cast<const char *>(&o) + f.offset);
}

bool is_power_of_two(int n) { return n > 0 && (n & (n - 1)) == 0; }

struct Slot {  // one pipeline stage of a host-memory extract
    cudaStream_t stream = nullptr;
    float *d_samples = nullptr;
    size_t samples_cap = 0;  // bytes
    char *d_out = nullptr;
    size_t out_cap = 0;  // bytes
    int64_t *d_tab = nullptr;
    int64_t *h_tab = nullptr;  // pinned
    size_t tab_cap = 0;        // int64 entries
    int *d_fix = nullptr;      // [0]: how many frames the float32 kernel flagged, [1 ..]: which (mb_adaptive.cuh)
    size_t fix_cap = 0;        // ints
    int *h_fix = nullptr;      // pinned: the counts of this slot's chunks, one after the other
    size_t h_fix_cap = 0, h_fix_used = 0;
    // The per-frame numbers and the short arrays (loudness.specific, mfcc) of a chunk leave the device as ONE copy into
    // this pinned staging area and are handed out to the caller's arrays once the slot's stream has drained: fifteen
    // copies of 12 .. 300 KB per chunk cost more in launch gaps than in bytes.
    char *h_small = nullptr;
    size_t small_cap = 0;
    struct SmallPart { int field; size_t off; };  // off: byte offset inside the staged region
    std::vector<SmallPart> small_parts;           // the pending chunk's small fields (empty: nothing pending)
    int64_t small_g0 = 0, small_frames = 0;
    // mb_set_host_rows(2): the pending chunk whose complexSpectrum upper halves the host threads mirror; the frames the
    // exact kernel redid (not conjugate-symmetric to the last bit, like the reference's) get theirs from the device
    int64_t mir_frames = 0;
    const int *mir_list = nullptr;                // this chunk's copy of d_fix: [0] how many frames were redone, [1 ..] which
    // the redone frames' upper halves, gathered on the device ([i][re | im][N/2 - 1], in list order), copied as one
    // piece when the slot drains and handed out at its next drain
    // (two areas, used in turn: the copy of one chunk's rows runs on the plan's auxiliary stream -- on the slot's own
    // stream it would hold the next chunk's kernels back behind the other slot's copies -- while the next chunk's
    // rows are gathered into the other)
    float *d_gather[2] = {nullptr, nullptr}, *h_gather[2] = {nullptr, nullptr};
    size_t gather_cap = 0;                        // frames
    int gat_t = 0;                                // the area the slot's current chunk gathers into
    cudaEvent_t gat_ev = nullptr;                 // the pending copy has arrived
    const int *gat_list = nullptr;                // pending hand-out: list, how many, first frame of the chunk, area
    int gat_cnt = 0, gat_from = 0;
    int64_t gat_g0 = 0;
    const float *mir_d_re = nullptr, *mir_d_im = nullptr;
};

}  // namespace

struct mb_plan {
    int device = 0, N = 0, hop = 0, window = 0;
    double sr = 0;
    uint32_t mask = 0, flags = 0;
    int num_sms = 0;
    MbDevPlan dev{};
    std::vector<float> h_window;
    float *d_window = nullptr, *d_dct = nullptr, *d_mel_inv = nullptr;
    float2 *d_twM = nullptr, *d_twN = nullptr;
    double2 *d_tw_exact = nullptr;
    double *d_mel_w_exact = nullptr;
    MbWarpTables *d_warp_tables = nullptr;
    MbWarpMfTables *d_warp_mf_tables = nullptr;
    bool has_warp_kernel = false;
    bool has_mf_kernel = false;
    bool use_cluster = false;
    bool has_big_kernel = false;
    bool has_exact_warp = false;  // exact-FFT arithmetic on the warp-per-frame kernel (kernel_exact_warp.cu)
    double tw_small[30] = {0};    // the recurrence twiddles of widths 1, 2, 4, 8 (its first register pass)
    int64_t launches_warp = 0, launches_generic = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // device-memory calls: clip tables staged through pinned memory
    int64_t *d_tab = nullptr, *h_tab = nullptr;
    size_t tab_cap = 0;
    cudaEvent_t tab_event = nullptr;
    bool tab_event_pending = false;
    Slot slots[2];
    std::vector<cudaEvent_t> landed;  // host-memory calls: chunk i's amplitude rows have arrived (kept across calls)
    std::vector<cudaEvent_t> landed_c;  // ... chunk i's complexSpectrum lower halves have arrived (mb_set_host_rows(2))
    cudaStream_t aux_stream = nullptr;  // host-memory calls: the redone frames' gathered upper halves travel on it
    int *h_fixlist = nullptr;           // pinned: every chunk's list of redone frames, one region per chunk of the call
    size_t fixlist_cap = 0;             // ints
    // adaptive exactness (mb_adaptive.cuh): frames the float32 kernels flag are redone by the exact-FFT kernel
    bool adaptive = false;
    MbDevPlan dev_fix{};       // the plan as the exact kernel sees it: spectral features only
    MbNoiseTables *d_noise = nullptr;
    int *d_fix = nullptr;      // device-memory calls and streams (one stream at a time)
    size_t fix_cap = 0;
    int64_t refined_frames = 0;  // host-memory calls: frames redone in the last call
    bool refined_on_device = false;  // the last call was a device-memory one: its count still lies in d_fix[0]
    int64_t launches = 0;
    int64_t bytes_per_frame = 0;
    const char *kernel_name = "generic";
};

struct StreamGraph {  // one captured push: same buffer parity, same fill level, same block size
    int cur = -1;
    int64_t filled = -1, n_new = -1;
    int seen = 0;
    cudaGraphExec_t exec = nullptr;
};

struct mb_stream {
    mb_plan *plan = nullptr;
    char *d_buf[2] = {nullptr, nullptr};
    size_t cap = 0;  // sample frames per buffer
    // what a sample frame is: one float32 (4 bytes), or `pcm_channels` interleaved int16 (mb_stream_create_pcm16)
    int pcm_channels = 0, pcm_channel = 0;
    size_t frame_bytes = 4;
    int cur = 0;
    int64_t filled = 0;
    int64_t skip = 0;  // samples still to drop before the next frame starts (hop > bufferSize)
    // host-memory pushes: pinned staging both ways, one device arena for all outputs, a private clip table
    char *h_in = nullptr;
    size_t h_in_cap = 0;  // sample frames
    char *d_out = nullptr, *h_out = nullptr;
    size_t out_cap = 0;  // bytes
    int64_t *d_tab = nullptr, *h_tab = nullptr;  // {clip_off = 0, frame_start = 0, nf}
    int64_t tab_nf = -1;
    StreamGraph graphs[4];
    int64_t graph_launches = 0;
};

namespace {

// Row producers of the host threads: streaming (non-temporal) stores, so that rows nobody reads again soon neither
// cost a read-for-ownership nor push the caller's data out of the caches.
#if defined(__SSE2__) || defined(__x86_64__)
#include <emmintrin.h>
inline void copy_row_stream(float *dst, const float *src, int64_t n) {
    if (((uintptr_t)dst & 15) == 0 && (n & 3) == 0) {
        for (int64_t i = 0; i < n; i += 4) _mm_stream_ps(dst + i, _mm_loadu_ps(src + i));
    } else {
        memcpy(dst, src, sizeof(float) * (size_t)n);
    }
}
inline void square_rows_stream(float *dst, const float *src, int64_t n) {
    int64_t i = 0;
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        for (; i + 4 <= n; i += 4) {
            const __m128 a = _mm_load_ps(src + i);
            _mm_stream_ps(dst + i, _mm_mul_ps(a, a));  // (one IEEE float32 multiply per element: what __fmul_rn does)
        }
    }
    for (; i < n; i++) dst[i] = src[i] * src[i];
}
// The upper half of a complexSpectrum row from its lower half: Z[N-k] = conj(Z[k]), k = 1 .. N/2-1, as every float32
// kernel stores it: the real part copied, the imaginary part negated BY AN ARITHMETIC INSTRUCTION -- so a NaN comes
// out as the canonical 0x7fffffff there, whatever its sign was (the kernels' own bits in every host-rows mode).
inline float mirror_neg(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u & 0x7fffffffu) > 0x7f800000u ? 0x7fffffffu : (u ^ 0x80000000u);
    memcpy(&x, &u, 4);
    return x;
}
inline void mirror_row_stream(float *row, int N, bool negate) {
    const int M = N / 2;
    int j = M + 1;
    if (((uintptr_t)row & 15) == 0 && M >= 8) {
        const __m128 sg = _mm_castsi128_ps(_mm_set1_epi32(negate ? (int)0x80000000u : 0));
        const __m128 qnan = _mm_castsi128_ps(_mm_set1_epi32(0x7fffffff));
        for (; j < M + 4; j++) row[j] = negate ? mirror_neg(row[N - j]) : row[N - j];
        for (; j + 4 <= N; j += 4) {  // dst[j .. j+3] = src[N-j], src[N-j-1], src[N-j-2], src[N-j-3]
            const __m128 a = _mm_loadu_ps(row + (N - j - 3));
            __m128 r = _mm_xor_ps(_mm_shuffle_ps(a, a, _MM_SHUFFLE(0, 1, 2, 3)), sg);
            if (negate) {
                const __m128 un = _mm_cmpunord_ps(r, r);
                r = _mm_or_ps(_mm_andnot_ps(un, r), _mm_and_ps(un, qnan));
            }
            _mm_stream_ps(row + j, r);
        }
    }
    for (; j < N; j++) row[j] = negate ? mirror_neg(row[N - j]) : row[N - j];
}
inline void stream_fence() { _mm_sfence(); }
#else
inline float mirror_neg(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u = (u & 0x7fffffffu) > 0x7f800000u ? 0x7fffffffu : (u ^ 0x80000000u);
    memcpy(&x, &u, 4);
    return x;
}
inline void mirror_row_stream(float *row, int N, bool negate) {
    for (int j = N / 2 + 1; j < N; j++) row[j] = negate ? mirror_neg(row[N - j]) : row[N - j];
}
inline void stream_fence() {}
#endif

// A few host threads that finish what needs no device: the `buffer` rows (the caller's own samples, framed) and the
// powerSpectrum rows (amplitude squared, float32: the same single rounding the kernels apply) of a host-memory call.
// Work may be gated on a CUDA event (rows that a device-to-host copy is still delivering): a worker takes a gated piece
// as soon as its event has completed, fills the wait with ungated pieces, and only blocks on an event when nothing
// else is left.
constexpr unsigned kDefaultHostThreads = 8;
std::atomic<int> g_host_threads{0};  // mb_set_host_threads (0: min(kDefaultHostThreads, cores / 2))
std::atomic<int> g_host_rows{-1};    // mb_set_host_rows: 1 buffer + powerSpectrum rows, 2 also the mirrored half of complexSpectrum, 0 off, -1 automatic (default)
// Automatic: the mirrored half as well where ONE device is visible and the host has twelve or more cores (one B200 on a
// 16-core host: 2.55 vs 2.12 M frames/s end to end with the full set at bufferSize 2048).  Mode 2 trades PCIe bytes for
// host memory traffic (54.7 against 46.7 KB per frame) and runs into the host's memory system at ~2.6 M frames/s however
// many devices share it: two B200s on one host reach 3.84 M in mode 1 and 2.80 M in mode 2, so hosts with several
// devices stay with mode 1.
int auto_host_rows() {
    static const int mode = []() {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) ndev = 1;
        (void)cudaGetLastError();
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return (ndev == 1 && hw >= 12u) ? 2 : 1;
    }();
    return mode;
}

class HostWorkers {
public:
    explicit HostWorkers(int n) {
        for (int i = 0; i < n; i++) threads_.emplace_back([this]() { run(); });
    }
    ~HostWorkers() {
        {
            std::lock_guard<std::mutex> lk(m_);
            done_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();  // (drains the queue first)
    }
    // Cut [g0, g1) into pieces and post f(piece begin, piece end) for each; `after` (optional): the event the pieces
    // wait for.  Gated pieces are posted in the order their events complete (one stream after the other, chunk by chunk).
    void post_range(int64_t g0, int64_t g1, int64_t grain, std::function<void(int64_t, int64_t)> f, cudaEvent_t after = nullptr) {
        {
            std::lock_guard<std::mutex> lk(m_);
            for (int64_t a = g0; a < g1; a += grain) {
                const int64_t b = std::min(g1, a + grain);
                if (after) gated_.push_back({after, [f, a, b]() { f(a, b); }});
                else q_.push_back([f, a, b]() { f(a, b); });
            }
        }
        cv_.notify_all();
    }

private:
    struct Gated {
        cudaEvent_t after;
        std::function<void()> f;
    };
    void run() {
        for (;;) {
            std::function<void()> f;
            cudaEvent_t wait_for = nullptr;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this]() { return done_ || !q_.empty() || !gated_.empty(); });
                if (!gated_.empty() && (q_.empty() || cudaEventQuery(gated_.front().after) != cudaErrorNotReady)) {
                    // (ready, or nothing else to do: then block on it below, outside the lock; an error is left to the
                    // stream synchronize of the calling thread to report)
                    if (q_.empty()) wait_for = gated_.front().after;
                    f = std::move(gated_.front().f);
                    gated_.pop_front();
                } else if (!q_.empty()) {
                    f = std::move(q_.front());
                    q_.pop_front();
                } else {
                    return;  // done_, both queues drained
                }
            }
            if (wait_for) (void)cudaEventSynchronize(wait_for);
            f();
        }
    }
    std::vector<std::thread> threads_;
    std::deque<std::function<void()>> q_;
    std::deque<Gated> gated_;
    std::mutex m_;
    std::condition_variable cv_;
    bool done_ = false;
};

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- plan tables (host, double precision, same formulas and evaluation
// order as the reference so that every discrete table is identical).

void build_window(std::vector<float> &w, int N, int which) {
    w.resize(N);
    for (int i = 0; i < N; i++) {
        if (which == MB_WINDOW_BLACKMAN)  // src/meyda.js:140-156 (commented out there: the formula it states)
            w[i] = (float)(0.42 - 0.5 * cos(2 * M_PI * i / (N - 1)) + 0.08 * cos(4 * M_PI * i / (N - 1)));
        else if (which == MB_WINDOW_HAMMING)  // src/meyda.js:116-126
            w[i] = (float)(0.54 - 0.46 * cos(2 * M_PI * ((double)i / N - 1)));
        else  // src/meyda.js:128-138
            w[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / (N - 1)));
    }
}

void build_bark_limits(int *bb, int N, double sr, int nb) {
    // src/meyda.js:170-182 then src/extractors/loudness.js:24-45
    const int n = N / 2;
    std::vector<float> bark(N);
    for (int i = 0; i < N; i++) {
        const float hz = (float)(i * sr / N);
        bark[i] = (float)(13 * atan((double)hz / 1315.8) + 3.5 * atan(pow((double)hz / 7518, 2)));
    }
    const double last = bark[n - 1];
    double band_end = last / nb;
    int band = 1;
    for (int i = 0; i <= nb; i++) bb[i] = 0;
    for (int i = 0; i < n; i++) {
        while ((double)bark[i] > band_end) {
            if (band <= nb) bb[band] = i;
            band++;
            band_end = band * last / nb;
        }
    }
    bb[nb] = n - 1;
}

void build_mel_bins(int *mel, int N, double sr, int nf) {
    // src/extractors/mfcc.js:7-38
    const double lower = 1125 * log(1 + 0.0 / 700), upper = 1125 * log(1 + (sr / 2) / 700);
    const double step = (upper - lower) / (nf + 1);
    for (int i = 0; i < nf + 2; i++) {
        const float m = (float)(i * step);
        const float hz = (float)(700 * (exp((double)m / 1125) - 1));
        int b = (int)floor((N + 1) * (double)hz / sr);
        mel[i] = std::min(std::max(b, 0), N / 2);  // sums only run over j < N/2 (mfcc.js:56)
    }
}

void build_dct(float *dct, int nf, int nc) {
    // src/extractors/mfcc.js:67-83
    const double k = M_PI / nf;
    const double w1 = 1.0 / sqrt((double)nf), w2 = sqrt(2.0 / nf);
    for (int i = 0; i < nc; i++)
        for (int j = 0; j < nf; j++)
            dct[i + j * nc] = (float)((i == 0 ? w1 : w2) * cos(k * (i + 1) * (j + 0.5)));
}

template <typename T>
cudaError_t upload(T **dst, const std::vector<T> &src) {
    cudaError_t e = cudaMalloc((void **)dst, std::max<size_t>(1, src.size()) * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

void free_slot(Slot &s) {
    if (s.stream) cudaStreamDestroy(s.stream);
    cudaFree(s.d_fix);
    if (s.h_fix) cudaFreeHost(s.h_fix);
    if (s.h_small) cudaFreeHost(s.h_small);
    for (int t = 0; t < 2; t++) {
        if (s.h_gather[t]) cudaFreeHost(s.h_gather[t]);
        cudaFree(s.d_gather[t]);
    }
    if (s.gat_ev) cudaEventDestroy(s.gat_ev);
    cudaFree(s.d_samples);
    cudaFree(s.d_out);
    cudaFree(s.d_tab);
    if (s.h_tab) cudaFreeHost(s.h_tab);
    s = Slot();
}

mb_status ensure_table(int64_t **d_tab, int64_t **h_tab, size_t *cap, size_t entries) {
    if (*cap >= entries) return MB_OK;
    size_t want = std::max<size_t>(entries, 1024);
    want += want / 2;
    cudaFree(*d_tab);
    if (*h_tab) cudaFreeHost(*h_tab);
    *d_tab = nullptr;
    *h_tab = nullptr;
    *cap = 0;
    MB_CUDA(cudaMalloc((void **)d_tab, want * sizeof(int64_t)));
    MB_CUDA(cudaMallocHost((void **)h_tab, want * sizeof(int64_t)));
    *cap = want;
    return MB_OK;
}

mb_status check_outputs(const mb_plan *p, const mb_outputs *out) {
    if (!out) return fail(MB_ERR_INVALID_ARG, "outputs struct is NULL");
    for (int i = 0; i < kNumFields; i++)
        if (mb_has(p->mask, kFields[i].feature) && !field_ptr(*out, kFields[i]))
            return fail(MB_ERR_MISSING_OUTPUT, "output pointer for requested feature '%s' is NULL",
                        kFeatureNames[kFields[i].feature]);
    return MB_OK;
}

mb_status check_clips(const mb_plan *p, int64_t n_samples, const int64_t *off, const int64_t *len, int64_t n_clips) {
    if (n_clips < 0 || n_samples < 0) return fail(MB_ERR_INVALID_ARG, "negative clip or sample count");
    if (n_clips > 0 && (!off || !len)) return fail(MB_ERR_INVALID_ARG, "clip_offset/clip_len is NULL");
    for (int64_t c = 0; c < n_clips; c++)
        if (off[c] < 0 || len[c] < 0 || off[c] > n_samples || len[c] > n_samples - off[c])
            return fail(MB_ERR_OUT_OF_RANGE, "clip %lld [%lld, +%lld) lies outside the %lld samples given",
                        (long long)c, (long long)off[c], (long long)len[c], (long long)n_samples);
    (void)p;
    return MB_OK;
}

// Launch the plan's kernel over `n` (virtual) clips whose offsets/prefix are
// already in device memory.
// `aligned`: every frame of this call starts on a 16-byte boundary (TMA bulk copies).
// `pcm_channels` > 0: d_samples is interleaved PCM of `pcm_format` (see MbClipTable).
// `fix`/`fix_cap`: where the float32 kernels list the frames to be redone exactly (grown here; NULL: plan not adaptive).
mb_status launch(mb_plan *p, const int64_t *d_off, const int64_t *d_frame_start, int64_t n, int64_t total_frames,
                 const float *d_samples, const mb_outputs &d_out, cudaStream_t stream, int **fix, size_t *fix_cap,
                 int pcm_channels = 0, int pcm_channel = 0, int pcm_format = MB_SAMPLE_S16, uint32_t drop_mask = 0) {
    if (total_frames == 0) return MB_OK;
    MbDevPlan dev_main = p->dev;
    dev_main.mask &= ~drop_mask;
    if (total_frames > 0x7fffffff) return fail(MB_ERR_UNSUPPORTED, "more than 2^31 - 1 frames in one launch");
    MbClipTable T{d_off, d_frame_start, n, total_frames, pcm_channels, pcm_channel, pcm_format, nullptr, nullptr, nullptr, nullptr};
    const uint32_t time_only = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) | MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    if ((dev_main.mask & MB_ALL_FEATURES) == 0) return MB_OK;  // (nothing left for the device to produce)
    const bool adaptive = p->adaptive && fix != nullptr && (dev_main.mask & ~time_only) != 0;
    if (adaptive) {
        if (*fix_cap < (size_t)total_frames + 1) {
            cudaFree(*fix);  // (synchronises with whatever still uses the old list)
            *fix = nullptr;
            *fix_cap = 0;
            const size_t want = (size_t)total_frames + (size_t)total_frames / 2 + 64;
            MB_CUDA(cudaMalloc((void **)fix, want * sizeof(int)));
            *fix_cap = want;
        }
        MB_CUDA(cudaMemsetAsync(*fix, 0, sizeof(int), stream));
        T.fix_count = *fix;
        T.fix_list = *fix + 1;
    }
    const bool pcm = pcm_channels > 0;
    // the tuned kernels take float32 mono or 16-bit PCM; other payloads go through the generic kernel's loader
    const bool tuned_ok = !pcm || pcm_format == MB_SAMPLE_S16;
    // the tuned kernels take any float-aligned frame (misaligned ones bypass TMA / the float4 loads inside the
    // kernel, same bits) but store `buffer` rows 16 bytes at a time
    const bool out_ok = (uintptr_t)d_out.buffer % 16 == 0;
    if (p->use_cluster) {
        MB_CUDA(mb_launch_exact_cluster(dev_main, T, d_samples, d_out, p->num_sms, stream));
        p->launches_generic++;
    } else if (p->dev.exact && p->has_exact_warp) {
        MB_CUDA(mb_launch_exact_warp(dev_main, T, d_samples, d_out, p->num_sms, stream, p->tw_small));
        p->launches_warp++;
    } else if (p->has_big_kernel && out_ok && !pcm && ((uintptr_t)d_samples % 4 == 0)) {  // (misaligned frames: read by the lanes)
        MB_CUDA(mb_launch_big32768(dev_main, T, d_samples, d_out, p->num_sms, stream));
        p->launches_warp++;
    } else if (p->has_warp_kernel && tuned_ok && out_ok && ((uintptr_t)d_samples % (pcm ? 2 : 4) == 0)) {
        MB_CUDA(mb_launch_warp2048(dev_main, T, d_samples, d_out, p->num_sms, stream));
        p->launches_warp++;
    } else if (p->has_mf_kernel && tuned_ok && out_ok && ((uintptr_t)d_samples % (pcm ? 2 : 4) == 0)) {
        MB_CUDA(mb_launch_warpmf(dev_main, T, d_samples, d_out, p->num_sms, stream));
        p->launches_warp++;
    } else {
        MB_CUDA(mb_launch_generic(dev_main, T, d_samples, d_out, p->num_sms, stream));
        p->launches_generic++;
    }
    p->launches++;
    if (adaptive) {
        // the flagged frames again, with the reference's own FFT arithmetic (the kernel reads the count on the device:
        // nothing to do is one empty launch)
        MbClipTable Tx = T;
        Tx.fix_count = nullptr;
        Tx.fix_list = nullptr;
        Tx.sel_count = *fix;
        Tx.sel_list = *fix + 1;
        MbDevPlan dev_fix = p->dev_fix;
        dev_fix.mask &= ~drop_mask;
        if (p->N > 16384) MB_CUDA(mb_launch_exact_cluster(dev_fix, Tx, d_samples, d_out, p->num_sms, stream));
        else if (p->has_exact_warp) MB_CUDA(mb_launch_exact_warp(dev_fix, Tx, d_samples, d_out, p->num_sms, stream, p->tw_small));
        else MB_CUDA(mb_launch_generic(dev_fix, Tx, d_samples, d_out, p->num_sms, stream));
        p->launches++;
        p->launches_generic++;
    }
    return MB_OK;
}

// Boundary bookkeeping of the warp kernel: the union of Bark limits and mel edges below M, in order.
void build_warp_tables(MbWarpTables &W, const MbDevPlan &D) {
    const int M = D.M;
    memset(&W, 0, sizeof(W));
    for (int c = 0; c < 32; c++)
        for (int b = 0; b < 32; b++) {
            const double ang = 2 * M_PI * (double)(b * c) / (double)M;
            W.tw32[c * 32 + b] = make_float2((float)cos(ang), (float)sin(ang));
        }
    std::vector<int> edges;
    for (int i = 0; i <= MB_NUM_BARK_BANDS; i++) edges.push_back(D.bb[i]);
    for (int i = 0; i < MB_NUM_MEL_FILTERS + 2; i++) edges.push_back(D.mel[i]);
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::vector<int> below;  // boundaries < M
    for (int e : edges)
        if (e < M) below.push_back(e);
    W.n_slots = (int)below.size();
    if (W.n_slots > MB_WARP_MAX_SLOTS || below.empty() || below[0] != 0) {
        W.n_slots = MB_WARP_MAX_SLOTS + 1;  // signals "does not fit" to the caller
        return;
    }
    std::vector<int> head_end(32);  // lane L's head piece covers [32 L, head_end[L])
    for (int lane = 0; lane < 32; lane++) {
        const auto first = std::lower_bound(below.begin(), below.end(), 32 * lane);
        W.lane_slot_base[lane] = (int)(first - below.begin());
        W.lane_seg_start[lane] = *(std::upper_bound(below.begin(), below.end(), 32 * lane) - 1);
        uint32_t m = 0;
        for (int i = 0; i < 32; i++)
            if (std::binary_search(below.begin(), below.end(), 32 * lane + i)) m |= 1u << i;
        W.lane_bmask[lane] = m;
        head_end[lane] = (first != below.end() && *first < 32 * lane + 32) ? *first : 32 * lane + 32;
    }
    // piece ids: a lane's pieces are consecutive, head first (the kernel walks them with one pointer)
    auto head_id = [&](int lane) { return W.lane_slot_base[lane] + lane; };
    auto boundary_id = [&](int s) { return s + below[s] / 32 + 1; };
    for (int lane = 0; lane < 32; lane++) W.piece_edge[head_id(lane)] = 32 * lane;  // a piece's first bin
    for (int s = 0; s < W.n_slots; s++) W.piece_edge[boundary_id(s)] = below[s];
    // pieces of every Bark band [bb[b], bb[b+1]) and mel segment [mel[s], mel[s+1])
    int n_items = 0;
    auto add_segment = [&](int seg, int e0, int e1) {
        W.seg_ptr[seg] = n_items;
        for (int s = 0; s < W.n_slots; s++)
            if (below[s] >= e0 && below[s] < e1 && n_items < MB_WARP_MAX_ITEMS)
                W.seg_items[n_items++] = (unsigned char)boundary_id(s);
        for (int lane = 0; lane < 32; lane++)
            if (head_end[lane] > 32 * lane && 32 * lane > e0 && 32 * lane < e1 && n_items < MB_WARP_MAX_ITEMS)
                W.seg_items[n_items++] = (unsigned char)head_id(lane);
    };
    for (int b = 0; b < MB_NUM_BARK_BANDS; b++) add_segment(b, D.bb[b], D.bb[b + 1]);
    for (int s = 0; s <= MB_NUM_MEL_FILTERS; s++) add_segment(MB_NUM_BARK_BANDS + s, D.mel[s], D.mel[s + 1]);
    W.seg_ptr[MB_WARP_SEGMENTS] = n_items;
    if (n_items >= MB_WARP_MAX_ITEMS) W.n_slots = MB_WARP_MAX_SLOTS + 1;
}

// The same bookkeeping for the multi-frame warp kernel (bufferSize 256 / 512 / 1024): the warp's 32 lanes are F frames
// of A = M / 32 lanes each; a lane's pieces are consecutive (head first), lanes in order, so piece ids run
// frame-major.  Segments: f * MB_WARP_SEGMENTS + (band b | 24 + mel segment s).
void build_warp_mf_tables(MbWarpMfTables &W, const MbDevPlan &D) {
    const int M = D.M, A = M / 32, F = 32 / A;
    memset(&W, 0, sizeof(W));
    for (int c = 0; c < 32; c++)
        for (int b = 0; b < 32; b++) {
            const double ang = 2 * M_PI * (double)((c % A) * b) / (double)M;
            W.tw32[c * 32 + b] = make_float2((float)cos(ang), (float)sin(ang));
        }
    std::vector<int> edges;
    for (int i = 0; i <= MB_NUM_BARK_BANDS; i++) edges.push_back(D.bb[i]);
    for (int i = 0; i < MB_NUM_MEL_FILTERS + 2; i++) edges.push_back(D.mel[i]);
    std::sort(edges.begin(), edges.end());
    edges.erase(std::unique(edges.begin(), edges.end()), edges.end());
    std::vector<int> below;  // boundaries < M (frame-local bins)
    for (int e : edges)
        if (e < M) below.push_back(e);
    const int nb = (int)below.size();
    W.n_pieces = F * (A + nb);
    const int piece_cap = A == 4 ? MB_MF_MAX_PIECES : 256;  // what the warp's slot holds (12- / 16-byte pieces)
    if (W.n_pieces > piece_cap || below.empty() || below[0] != 0) {
        W.n_pieces = MB_MF_MAX_PIECES + 1;  // signals "does not fit" to the caller
        return;
    }
    // piece ids of one frame: row r's head, then the boundaries inside row r, rows in order
    std::vector<int> head_id(A), head_end(A), boundary_id(nb);
    int id = 0;
    for (int r = 0; r < A; r++) {
        head_id[r] = id++;
        head_end[r] = 32 * r + 32;
        uint32_t m = 0;
        for (int s = 0; s < nb; s++)
            if (below[s] >= 32 * r && below[s] < 32 * r + 32) {
                if (head_end[r] == 32 * r + 32) head_end[r] = below[s];
                boundary_id[s] = id++;
                m |= 1u << (below[s] - 32 * r);
            }
        for (int f = 0; f < F; f++) W.lane_bmask[f * A + r] = m;
    }
    const int per_frame = id;  // == A + nb
    for (int f = 0; f < F; f++)
        for (int r = 0; r < A; r++) W.lane_slot_base[f * A + r] = f * per_frame + head_id[r] - (f * A + r);
    for (int f = 0; f < F; f++) {
        for (int r = 0; r < A; r++) W.piece_edge[f * per_frame + head_id[r]] = (short)(32 * r);
        for (int s = 0; s < nb; s++) W.piece_edge[f * per_frame + boundary_id[s]] = (short)below[s];  // a piece's first bin
    }
    int n_items = 0;
    bool overflow = false;
    auto add_segment = [&](int f, int seg, int e0, int e1) {
        W.seg_ptr[f * MB_WARP_SEGMENTS + seg] = (short)n_items;
        auto push = [&](int pid) {
            if (n_items < MB_MF_MAX_ITEMS) W.seg_items[n_items++] = (unsigned short)(f * per_frame + pid);
            else overflow = true;
        };
        for (int s = 0; s < nb; s++)
            if (below[s] >= e0 && below[s] < e1) push(boundary_id[s]);
        for (int r = 0; r < A; r++)
            if (head_end[r] > 32 * r && 32 * r > e0 && 32 * r < e1) push(head_id[r]);
    };
    for (int f = 0; f < F; f++) {
        for (int b = 0; b < MB_NUM_BARK_BANDS; b++) add_segment(f, b, D.bb[b], D.bb[b + 1]);
        for (int s = 0; s <= MB_NUM_MEL_FILTERS; s++) add_segment(f, MB_NUM_BARK_BANDS + s, D.mel[s], D.mel[s + 1]);
    }
    for (int i = F * MB_WARP_SEGMENTS; i <= MB_MF_MAX_SEGMENTS; i++) W.seg_ptr[i] = (short)n_items;
    if (overflow) W.n_pieces = MB_MF_MAX_PIECES + 1;
}

void offset_outputs(mb_outputs &o, const mb_outputs &base, int64_t frame0, const MbDevPlan &D) {
    o = base;
    for (int i = 0; i < kNumFields; i++) {
        void *b = field_ptr(base, kFields[i]);
        if (b) field_ptr(o, kFields[i]) = (char *)b + (size_t)frame0 * field_elems(kFields[i], D) * 4;
    }
}

template <typename T>
__global__ void mb_fma_peak_kernel(T *out, int iters) {
    T a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const T b = (T)1.000001, c = (T)0.5;
    for (int i = 0; i < iters; i++) {
        a0 = a0 * b + c; a1 = a1 * b + c; a2 = a2 * b + c; a3 = a3 * b + c;
        a4 = a4 * b + c; a5 = a5 * b + c; a6 = a6 * b + c; a7 = a7 * b + c;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

extern "C" {

int mb_version(void) { return MB_VERSION; }
const char *mb_last_error(void) { return g_last_error.c_str(); }

const char *mb_feature_name(int feature) {
    return (feature >= 0 && feature < MB_NUM_FEATURES) ? kFeatureNames[feature] : nullptr;
}

int mb_feature_from_name(const char *name) {
    if (!name) return -1;
    for (int i = 0; i < MB_NUM_FEATURES; i++)
        if (strcmp(name, kFeatureNames[i]) == 0) return i;
    return -1;
}

mb_status mb_device_count(int *count) {
    if (!count) return fail(MB_ERR_INVALID_ARG, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(MB_ERR_NO_DEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return MB_OK;
}

int64_t mb_num_frames(int64_t clip_len, int buffer_size, int hop) {
    if (buffer_size <= 0 || hop <= 0 || clip_len < buffer_size) return 0;
    return (clip_len - buffer_size) / hop + 1;
}

mb_status mb_plan_create(mb_plan **plan, int device, int buffer_size, int hop, double sample_rate, int window,
                         uint32_t feature_mask, uint32_t flags) {
    return mb_plan_create_ex(plan, device, buffer_size, hop, sample_rate, window, feature_mask, flags, nullptr);
}

mb_status mb_plan_get_params(const mb_plan *p, mb_params *params) {
    if (!p || !params) return fail(MB_ERR_INVALID_ARG, "plan or params is NULL");
    params->num_bark_bands = p->dev.nb;
    params->num_mel_filters = p->dev.nf;
    params->num_mfcc = p->dev.nc;
    params->reserved = 0;
    params->rolloff_fraction = p->dev.rolloff_frac;
    return MB_OK;
}

mb_status mb_plan_create_ex(mb_plan **plan, int device, int buffer_size, int hop, double sample_rate, int window,
                            uint32_t feature_mask, uint32_t flags, const mb_params *params) {
    if (!plan) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    *plan = nullptr;
    // the constants of loudness.js:14, mfcc.js:15,71 and spectralRolloff.js:9 unless the caller says otherwise
    int nb = MB_NUM_BARK_BANDS, nf = MB_NUM_MEL_FILTERS, nc = MB_NUM_MFCC;
    double rolloff_frac = 0.99;
    if (params) {
        if (params->reserved != 0) return fail(MB_ERR_INVALID_ARG, "mb_params.reserved must be 0");
        if (params->num_bark_bands) nb = params->num_bark_bands;
        if (params->num_mel_filters) nf = params->num_mel_filters;
        if (params->num_mfcc) nc = params->num_mfcc;
        if (params->rolloff_fraction != 0) rolloff_frac = params->rolloff_fraction;
        if (nb < 1 || nb > MB_MAX_BARK_BANDS)
            return fail(MB_ERR_INVALID_ARG, "num_bark_bands %d outside [1, %d]", nb, MB_MAX_BARK_BANDS);
        if (nf < 1 || nf > MB_MAX_MEL_FILTERS)
            return fail(MB_ERR_INVALID_ARG, "num_mel_filters %d outside [1, %d]", nf, MB_MAX_MEL_FILTERS);
        if (nc < 1 || nc > nf) return fail(MB_ERR_INVALID_ARG, "num_mfcc %d outside [1, num_mel_filters = %d]", nc, nf);
        if (!(rolloff_frac > 0 && rolloff_frac <= 1))
            return fail(MB_ERR_INVALID_ARG, "rolloff_fraction %g outside (0, 1]", rolloff_frac);
    }
    // the tuned warp kernels are built around the reference's constants; anything else runs on the generic family
    const bool reference_params = nb == MB_NUM_BARK_BANDS && nf == MB_NUM_MEL_FILTERS && nc == MB_NUM_MFCC && rolloff_frac == 0.99;
    if (!is_power_of_two(buffer_size))  // src/meyda.js:20-22
        return fail(MB_ERR_NOT_POWER_OF_TWO, "Buffer size is not a power of two: Meyda will not run.");
    if (buffer_size < MB_MIN_BUFFER_SIZE || buffer_size > MB_MAX_BUFFER_SIZE)
        return fail(MB_ERR_UNSUPPORTED, "bufferSize %d outside [%d, %d]", buffer_size, MB_MIN_BUFFER_SIZE,
                    MB_MAX_BUFFER_SIZE);
    if ((flags & MB_FLAG_EXACT_FFT) && buffer_size > MB_MAX_EXACT_BUFFER_SIZE)
        return fail(MB_ERR_UNSUPPORTED, "exact-FFT mode supports bufferSize <= %d (got %d)", MB_MAX_EXACT_BUFFER_SIZE,
                    buffer_size);
    if (flags & ~(uint32_t)(MB_FLAG_GENERIC_KERNEL | MB_FLAG_EXACT_FFT | MB_FLAG_CLUSTER_FFT | MB_FLAG_NO_REFINE))
        return fail(MB_ERR_INVALID_ARG, "unknown plan flags 0x%x", flags);
    if (hop <= 0) return fail(MB_ERR_INVALID_ARG, "hop must be positive (got %d)", hop);
    if (!(sample_rate > 0)) return fail(MB_ERR_INVALID_ARG, "sampleRate must be positive");
    if (window != MB_WINDOW_HANNING && window != MB_WINDOW_HAMMING && window != MB_WINDOW_BLACKMAN)
        return fail(MB_ERR_INVALID_ARG, "unknown windowingFunction %d", window);
    if (feature_mask == 0 || (feature_mask & ~MB_ALL_FEATURES))
        return fail(MB_ERR_INVALID_ARG, "feature mask 0x%x is empty or has unknown bits", feature_mask);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(MB_ERR_NO_DEVICE, "no CUDA device available: the Meyda B200 path has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(MB_ERR_NO_DEVICE, "device %d out of range (have %d)", device, ndev);
    DeviceGuard guard(device);
    if (!guard.ok) return fail(MB_ERR_CUDA, "cudaSetDevice(%d) failed", device);

    mb_plan *p = new mb_plan();
    p->device = device;
    p->N = buffer_size;
    p->hop = hop;
    p->sr = sample_rate;
    p->window = window;
    p->mask = feature_mask;
    p->flags = flags;
    const int N = buffer_size, M = N / 2;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete p;
        return fail(MB_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    }
    p->num_sms = prop.multiProcessorCount;

    MbDevPlan &D = p->dev;
    D.N = N;
    D.M = M;
    D.log2M = 0;
    while ((1 << D.log2M) < M) D.log2M++;
    D.hop = hop;
    D.mask = feature_mask;
    D.inv_sqrt_N = (float)(1.0 / sqrt((double)N));
    D.sr = sample_rate;
    {  // spectralSlope.js:9-16 and spectralRolloff.js:4 constants, accumulated in the reference's order
        double fs = 0, pfs = 0;
        for (int i = 0; i < M; i++) {
            const double f = i * sample_rate / N;
            pfs += f * f;
            fs += f;
        }
        D.slope_freq_sum = fs;
        D.slope_pow_freq_sum = pfs;
        D.rolloff_bin_hz = sample_rate / (2 * (M - 1));
        double sc = 0;
        for (int i = 15; i < nb; i++) sc += 0.066 * exp(0.171 * (i + 1));
        // perceptualSharpness.js:6-8 reads spec[i + 1] for every i < min(15, length): with 15 bands or fewer the
        // last read is past the end (undefined) and the reference returns NaN
        D.sharp_const = nb >= 16 ? sc : nan("");
    }
    D.nb = nb;
    D.nf = nf;
    D.nc = nc;
    D.rolloff_frac = rolloff_frac;
    build_window(p->h_window, N, window);
    build_bark_limits(D.bb, N, sample_rate, nb);
    build_mel_bins(D.mel, N, sample_rate, nf);
    std::vector<float> dct((size_t)nc * nf), mel_inv(nf + 1);
    build_dct(dct.data(), nf, nc);
    for (int s = 0; s <= nf; s++) {
        const int w = D.mel[s + 1] - D.mel[s];
        mel_inv[s] = w > 0 ? (float)(1.0 / w) : 0.f;
    }
    std::vector<float2> twM(std::max(1, M / 2)), twN(M);
    for (int j = 0; j < M / 2; j++) twM[j] = make_float2((float)cos(2 * M_PI * j / M), (float)sin(2 * M_PI * j / M));
    for (int k = 0; k < M; k++) twN[k] = make_float2((float)cos(2 * M_PI * k / N), (float)sin(2 * M_PI * k / N));

    // the exact-FFT tables also serve the adaptive path of the float32 kernels (mb_adaptive.cuh)
    const uint32_t time_only_mask = MB_FEATURE_BIT(MB_FEAT_RMS) | MB_FEATURE_BIT(MB_FEAT_ENERGY) |
                                    MB_FEATURE_BIT(MB_FEAT_ZCR) | MB_FEATURE_BIT(MB_FEAT_BUFFER);
    const bool may_refine = !(flags & (MB_FLAG_EXACT_FFT | MB_FLAG_NO_REFINE)) && (feature_mask & ~time_only_mask) != 0;
    const bool want_exact_tables = (flags & MB_FLAG_EXACT_FFT) || may_refine;
    std::vector<double2> tw_exact;
    if (want_exact_tables) {
        // lib/jsfft/fft.js:143-164: per stage del = (cos, sin)(PI / width); f <- f * del, in doubles
        tw_exact.resize(N - 1);
        for (int width = 1; width < N; width <<= 1) {
            const double del_r = cos(M_PI / width), del_i = sin(M_PI / width);
            double f_r = 1, f_i = 0;
            for (int j = 0; j < width; j++) {
                tw_exact[width - 1 + j] = make_double2(f_r, f_i);
                if (width <= 8) {
                    p->tw_small[2 * (width - 1 + j)] = f_r;
                    p->tw_small[2 * (width - 1 + j) + 1] = f_i;
                }
                const double temp = f_r * del_r - f_i * del_i;
                f_i = f_r * del_i + f_i * del_r;
                f_r = temp;
            }
        }
    }
    std::vector<double> mel_w;
    if (want_exact_tables) {  // src/extractors/mfcc.js:45-50, evaluated in doubles as the reference does
        for (int f = 0; f < nf; f++) {
            D.mel_w_off[f] = (int)mel_w.size();
            const int e0 = D.mel[f], e1 = D.mel[f + 1], e2 = D.mel[f + 2];
            for (int k = e0; k < e1; k++) mel_w.push_back((double)(k - e0) / (double)(e1 - e0));
            for (int k = e1; k < e2; k++) mel_w.push_back((double)(e2 - k) / (double)(e2 - e1));
        }
        D.mel_w_off[nf] = (int)mel_w.size();
    }
    bool ok = upload(&p->d_mel_w_exact, mel_w) == cudaSuccess && upload(&p->d_tw_exact, tw_exact) == cudaSuccess && upload(&p->d_window, p->h_window) == cudaSuccess && upload(&p->d_dct, dct) == cudaSuccess &&
              upload(&p->d_mel_inv, mel_inv) == cudaSuccess && upload(&p->d_twM, twM) == cudaSuccess &&
              upload(&p->d_twN, twN) == cudaSuccess &&
              cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&p->tab_event, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        mb_status st = fail(MB_ERR_CUDA, "plan table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        mb_plan_destroy(p);
        return st;
    }
    p->stream = p->own_stream;
    D.window = p->d_window;
    D.dct = p->d_dct;
    D.mel_inv_width = p->d_mel_inv;
    D.twM = p->d_twM;
    D.twN = p->d_twN;
    D.tw_exact = p->d_tw_exact;
    D.mel_w_exact = p->d_mel_w_exact;
    D.exact = (flags & MB_FLAG_EXACT_FFT) ? 1 : 0;
    if (D.exact) p->kernel_name = "generic-exact";
    // the warp-per-frame exact kernel: MB_FLAG_EXACT_FFT plans of its sizes, and the second pass of the adaptive plans
    p->has_exact_warp = want_exact_tables && mb_exact_warp_supports(N) && !(flags & MB_FLAG_GENERIC_KERNEL) &&
                        (size_t)prop.sharedMemPerBlockOptin >= mb_exact_warp_smem_bytes(N) + 16384;  // (+ its static scratch)
    if (D.exact && p->has_exact_warp) p->kernel_name = N == 2048 ? "exactw2048" : N == 1024 ? "exactw1024" : "exactw512";
    if (D.exact && (N > 16384 || ((flags & MB_FLAG_CLUSTER_FFT) && N >= 64))) {
        p->use_cluster = true;
        p->kernel_name = "exact-cluster2";
    }
    D.warp_tables = nullptr;
    if (N >= 4096 && !D.exact && !(flags & MB_FLAG_GENERIC_KERNEL) &&  // (shares the generic epilogue: any parameters)
        (size_t)prop.sharedMemPerBlockOptin >= mb_big_smem_bytes(N) + 4096) {  // (+ the kernel's static scratch)
        p->has_big_kernel = true;  // bufferSize / 2048 warps per frame
        p->kernel_name = N == 32768 ? "big32768" : N == 16384 ? "big16384" : N == 8192 ? "big8192" : "big4096";
    }
    if (N == 2048 && !D.exact && !(flags & MB_FLAG_GENERIC_KERNEL) && reference_params) {
        MbWarpTables *W = new MbWarpTables();
        build_warp_tables(*W, D);
        const bool fits = W->n_slots <= MB_WARP_MAX_SLOTS && (size_t)prop.sharedMemPerBlockOptin >= mb_warp2048_smem_bytes();
        cudaError_t we = cudaSuccess;
        if (fits) {
            we = cudaMalloc((void **)&p->d_warp_tables, sizeof(MbWarpTables));
            if (we == cudaSuccess) we = cudaMemcpy(p->d_warp_tables, W, sizeof(MbWarpTables), cudaMemcpyHostToDevice);
        }
        delete W;
        if (we != cudaSuccess) {
            mb_status st = fail(MB_ERR_CUDA, "warp-kernel table upload failed: %s", cudaGetErrorString(we));
            mb_plan_destroy(p);
            return st;
        }
        if (fits) {
            D.warp_tables = p->d_warp_tables;
            p->has_warp_kernel = true;
            p->kernel_name = "warp2048";
        }
    }
    D.warp_mf_tables = nullptr;
    if ((N == 256 || N == 512 || N == 1024) && !D.exact && !(flags & MB_FLAG_GENERIC_KERNEL) && reference_params) {
        MbWarpMfTables *W = new MbWarpMfTables();
        build_warp_mf_tables(*W, D);
        const bool fits = W->n_pieces <= MB_MF_MAX_PIECES && (size_t)prop.sharedMemPerBlockOptin >= mb_warpmf_smem_bytes();
        cudaError_t we = cudaSuccess;
        if (fits) {
            we = cudaMalloc((void **)&p->d_warp_mf_tables, sizeof(MbWarpMfTables));
            if (we == cudaSuccess) we = cudaMemcpy(p->d_warp_mf_tables, W, sizeof(MbWarpMfTables), cudaMemcpyHostToDevice);
        }
        delete W;
        if (we != cudaSuccess) {
            mb_status st = fail(MB_ERR_CUDA, "warp-kernel table upload failed: %s", cudaGetErrorString(we));
            mb_plan_destroy(p);
            return st;
        }
        if (fits) {
            D.warp_mf_tables = p->d_warp_mf_tables;
            p->has_mf_kernel = true;
            p->kernel_name = N == 256 ? "warpmf256" : N == 512 ? "warpmf512" : "warpmf1024";
        }
    }
    {   // constants of the noise bounds (mb_adaptive.cuh)
        MbNoiseTables *NT = new MbNoiseTables();
        memset(NT, 0, sizeof(*NT));
        for (int b = 0; b < nb; b++) {
            const double n_b = (double)std::max(0, D.bb[b + 1] - D.bb[b]);
            NT->band_c[b] = (float)(n_b > 0 ? 2.0 * n_b + (double)kMbNoiseK * sqrt(n_b) : 0.0);
        }
        for (int f = 0; f < nf; f++) {  // total weight of filter f over the bins below N/2 (mfcc.js:40-51)
            const int e0 = D.mel[f], e1 = D.mel[f + 1], e2 = std::min(D.mel[f + 2], M);
            double W = 0;
            for (int k = e0; k < e1 && k < M; k++) W += (double)(k - e0) / (double)(e1 - e0);
            for (int k = e1; k < e2; k++) W += (double)(D.mel[f + 2] - k) / (double)(D.mel[f + 2] - e1);
            NT->mel_c1[f] = (float)(W > 0 ? 2.0 * (double)kMbNoiseK * sqrt(W / std::max(W, 1.0)) : 0.0);
            NT->mel_c2[f] = (float)(4.0 * W);
        }
        for (int q = 0; q < 5; q++) {
            double t = 0;
            for (int k = 0; k < M; k++) t += pow((double)k, 2.0 * q);
            D.noise_sqrtT[q] = kMbNoiseKap * sqrt(t);
        }
        cudaError_t ne = cudaMalloc((void **)&p->d_noise, sizeof(MbNoiseTables));
        if (ne == cudaSuccess) ne = cudaMemcpy(p->d_noise, NT, sizeof(MbNoiseTables), cudaMemcpyHostToDevice);
        delete NT;
        if (ne != cudaSuccess) {
            mb_status st = fail(MB_ERR_CUDA, "noise-bound table upload failed: %s", cudaGetErrorString(ne));
            mb_plan_destroy(p);
            return st;
        }
        D.noise = p->d_noise;
    }
    // Adaptive exactness: the float32 kernels flag the frames whose features sit in the reference's own rounding noise
    // and the exact-FFT kernel redoes exactly those (time-domain features never need it).
    p->adaptive = may_refine;  // every float32 kernel family flags (warp, multi-frame warp, generic, multi-warp-per-frame)
    p->dev_fix = D;
    p->dev_fix.mask = feature_mask & ~time_only_mask;
    p->dev_fix.exact = 1;
    p->bytes_per_frame = 0;
    for (int i = 0; i < kNumFields; i++)
        if (mb_has(feature_mask, kFields[i].feature)) p->bytes_per_frame += 4 * (int64_t)field_elems(kFields[i], D);
    *plan = p;
    return MB_OK;
}

void mb_plan_destroy(mb_plan *p) {
    if (!p) return;
    DeviceGuard guard(p->device);
    if (p->own_stream) cudaStreamSynchronize(p->own_stream);
    for (auto &s : p->slots) free_slot(s);
    for (cudaEvent_t ev : p->landed) cudaEventDestroy(ev);
    for (cudaEvent_t ev : p->landed_c) cudaEventDestroy(ev);
    if (p->h_fixlist) cudaFreeHost(p->h_fixlist);
    if (p->aux_stream) cudaStreamDestroy(p->aux_stream);
    cudaFree(p->d_window);
    cudaFree(p->d_dct);
    cudaFree(p->d_mel_inv);
    cudaFree(p->d_twM);
    cudaFree(p->d_twN);
    cudaFree(p->d_tw_exact);
    cudaFree(p->d_mel_w_exact);
    cudaFree(p->d_warp_tables);
    cudaFree(p->d_warp_mf_tables);
    cudaFree(p->d_noise);
    cudaFree(p->d_fix);
    cudaFree(p->d_tab);
    if (p->h_tab) cudaFreeHost(p->h_tab);
    if (p->tab_event) cudaEventDestroy(p->tab_event);
    if (p->own_stream) cudaStreamDestroy(p->own_stream);
    delete p;
}

mb_status mb_plan_set_stream(mb_plan *p, void *cuda_stream) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
    if (next != p->stream) {
        // the plan's clip table and flagged-frame list are ordered by the stream alone: work still in flight on
        // the old stream must not meet the next call's uploads on the new one
        DeviceGuard guard(p->device);
        MB_CUDA(cudaStreamSynchronize(p->stream));
        p->stream = next;
    }
    return MB_OK;
}

mb_status mb_plan_tables(const mb_plan *p, float *window, int32_t *bb_limits, int32_t *mel_bins) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    if (window) memcpy(window, p->h_window.data(), sizeof(float) * p->N);
    if (bb_limits) memcpy(bb_limits, p->dev.bb, sizeof(int) * (p->dev.nb + 1));
    if (mel_bins) memcpy(mel_bins, p->dev.mel, sizeof(int) * (p->dev.nf + 2));
    return MB_OK;
}

mb_status mb_query_output(const mb_plan *p, int64_t n_clips, const int64_t *clip_len, int64_t *frames_per_clip,
                          mb_layout *layout) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    if (n_clips < 0 || (n_clips > 0 && !clip_len)) return fail(MB_ERR_INVALID_ARG, "bad clip list");
    int64_t total = 0;
    for (int64_t c = 0; c < n_clips; c++) {
        if (clip_len[c] < 0) return fail(MB_ERR_INVALID_ARG, "clip %lld has negative length", (long long)c);
        const int64_t f = mb_num_frames(clip_len[c], p->N, p->hop);
        if (frames_per_clip) frames_per_clip[c] = f;
        total += f;
    }
    if (layout) {
        layout->total_frames = total;
        layout->buffer_size = p->N;
        layout->spectrum_size = p->N / 2;
        layout->feature_mask = p->mask;
        layout->reserved = 0;
        layout->bytes_per_frame = p->bytes_per_frame;
        layout->output_bytes = p->bytes_per_frame * total;
        layout->num_bark_bands = p->dev.nb;
        layout->num_mfcc = p->dev.nc;
    }
    return MB_OK;
}

static mb_status extract_device(mb_plan *p, const float *samples, int64_t n_samples, const int64_t *clip_offset,
                                const int64_t *clip_len, int64_t n_clips, const mb_outputs *out, int pcm_channels,
                                int pcm_channel, int pcm_format = MB_SAMPLE_S16) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    mb_status st = check_outputs(p, out);
    if (st != MB_OK) return st;
    st = check_clips(p, n_samples, clip_offset, clip_len, n_clips);
    if (st != MB_OK) return st;
    if (n_clips == 0) return MB_OK;
    if (!samples) return fail(MB_ERR_INVALID_ARG, "samples is NULL");
    DeviceGuard guard(p->device);
    const size_t entries = 2 * (size_t)n_clips + 1;
    if (p->tab_event_pending) {  // the staging buffer may still be in flight from the previous call
        MB_CUDA(cudaEventSynchronize(p->tab_event));
        p->tab_event_pending = false;
    }
    st = ensure_table(&p->d_tab, &p->h_tab, &p->tab_cap, entries);
    if (st != MB_OK) return st;
    int64_t *h_off = p->h_tab, *h_fs = p->h_tab + n_clips;
    int64_t total = 0;
    for (int64_t c = 0; c < n_clips; c++) {
        h_off[c] = clip_offset[c];
        h_fs[c] = total;
        total += mb_num_frames(clip_len[c], p->N, p->hop);
    }
    h_fs[n_clips] = total;
    MB_CUDA(cudaMemcpyAsync(p->d_tab, p->h_tab, entries * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
    MB_CUDA(cudaEventRecord(p->tab_event, p->stream));
    p->tab_event_pending = true;
    p->refined_on_device = p->adaptive && total > 0;
    p->refined_frames = 0;
    return launch(p, p->d_tab, p->d_tab + n_clips, n_clips, total, samples, *out, p->stream, &p->d_fix, &p->fix_cap,
                  pcm_channels, pcm_channel, pcm_format);
}

mb_status mb_extract_async(mb_plan *p, const float *samples, int64_t n_samples, const int64_t *clip_offset,
                           const int64_t *clip_len, int64_t n_clips, const mb_outputs *out) {
    return extract_device(p, samples, n_samples, clip_offset, clip_len, n_clips, out, 0, 0);
}

static mb_status check_pcm(int channels, int channel) {
    if (channels < 1 || channels > 64) return fail(MB_ERR_INVALID_ARG, "channel count %d outside [1, 64]", channels);
    if (channel < 0 || channel >= channels) return fail(MB_ERR_INVALID_ARG, "channel %d of %d", channel, channels);
    return MB_OK;
}

mb_status mb_extract_pcm16_async(mb_plan *p, const int16_t *pcm, int64_t n_sample_frames, int channels, int channel,
                                 const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                                 const mb_outputs *out) {
    mb_status st = check_pcm(channels, channel);
    if (st != MB_OK) return st;
    return extract_device(p, reinterpret_cast<const float *>(pcm), n_sample_frames, clip_offset, clip_len, n_clips, out,
                          channels, channel);
}

mb_status mb_plan_synchronize(mb_plan *p) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    DeviceGuard guard(p->device);
    MB_CUDA(cudaStreamSynchronize(p->stream));
    return MB_OK;
}

// Host-memory extract: frames are cut into chunks (possibly inside a clip);
// each chunk is copied in, processed and copied out on one of two streams so
// that PCIe traffic in both directions overlaps the kernels.
// `samples` is float32 (pcm_channels == 0) or interleaved int16 PCM; offsets count sample frames either way.
static mb_status extract_host_impl(mb_plan *p, const void *samples, const int64_t *clip_offset, const int64_t *clip_len,
                                   int64_t n_clips, const mb_outputs *out, int pcm_channels, int pcm_channel, int pcm_format);
// Block b < fix[0] (and < cap) packs bins N/2+1 .. N-1 of both complexSpectrum rows of frame fix[1 + b] into dst[b].
__global__ void mb_gather_upper_kernel(const int *__restrict__ fix, int cap, const float *__restrict__ re,
                                       const float *__restrict__ im, int N, float *__restrict__ dst) {
    const int b = blockIdx.x, cnt = min(fix[0], cap);
    if (b >= cnt) return;
    const int up = N / 2 - 1;
    const size_t row = (size_t)fix[1 + b] * N + N / 2 + 1;
    float *d = dst + (size_t)b * 2 * up;
    for (int i = threadIdx.x; i < up; i += blockDim.x) {
        d[i] = re[row + i];
        d[up + i] = im[row + i];
    }
}

// the staged small fields of the slot's last chunk into the caller's arrays (the slot's stream has been synchronized)
static void scatter_small(mb_plan *p, Slot &s, const mb_outputs *out) {
    for (const Slot::SmallPart &sp : s.small_parts) {
        const OutField &f = kFields[sp.field];
        const size_t per = (size_t)field_elems(f, p->dev) * 4;
        memcpy((char *)field_ptr(*out, f) + (size_t)s.small_g0 * per, s.h_small + sp.off, (size_t)s.small_frames * per);
    }
    s.small_parts.clear();
}

static mb_status extract_host(mb_plan *p, const void *samples, const int64_t *clip_offset, const int64_t *clip_len,
                              int64_t n_clips, const mb_outputs *out, int pcm_channels = 0, int pcm_channel = 0,
                              int pcm_format = MB_SAMPLE_S16) {
    DeviceGuard guard(p->device);
    const mb_status st = extract_host_impl(p, samples, clip_offset, clip_len, n_clips, out, pcm_channels, pcm_channel, pcm_format);
    if (st != MB_OK) {
        // copies of earlier chunks may still be writing into the caller's arrays: drain before reporting the failure
        const std::string msg = g_last_error;
        if (p->aux_stream) cudaStreamSynchronize(p->aux_stream);
        for (auto &s : p->slots) {
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.small_parts.clear();  // (a failed call's staged rows are not handed out)
            s.mir_frames = 0;
            s.mir_list = nullptr;
            s.gat_cnt = 0;
        }
        (void)cudaGetLastError();
        g_last_error = msg;
    }
    return st;
}

static mb_status extract_host_impl(mb_plan *p, const void *samples, const int64_t *clip_offset, const int64_t *clip_len,
                                   int64_t n_clips, const mb_outputs *out, int pcm_channels, int pcm_channel, int pcm_format) {
    const int N = p->N, hop = p->hop;
    const size_t frame_bytes =
        pcm_channels > 0 ? (size_t)mb_sample_bytes(pcm_format) * (size_t)pcm_channels : 4u;  // bytes per sample frame
    const int64_t bpf = std::max<int64_t>(p->bytes_per_frame, 4);
    (void)bpf;
    // chunk budget: ~64 MiB of output or ~64 MiB of fresh input, whichever is hit first
    // (64 MiB of output per chunk: ~1 ms of PCIe each, so that the pipeline's fill and the host threads' tail stay short
    // against a moderate batch; the per-chunk overhead is ~20 asynchronous copies)
    const int64_t max_frames_out = std::max<int64_t>(1, (64ll << 20) / std::max<int64_t>(p->bytes_per_frame, 4));
    const int64_t max_frames_in = std::max<int64_t>(1, (64ll << 20) / (4ll * hop));
    const int64_t chunk_frames = std::min(max_frames_out, max_frames_in);

    // `buffer` is the caller's own samples cut into frames (docs.md:19-21): for float32 input the host fills those rows
    // itself, on a few threads, while the device works -- bit-identical by construction, and a quarter of the full
    // set's output bytes (8 KB of 33 KB per frame at bufferSize 2048) never crosses PCIe.
    // (measured with one and with eight devices per host, ranks or one process: on the host wins both times,
    // 2.02 vs 1.49 M and 3.33 vs 2.82 M frames/s; mb_set_host_rows(0) is for hosts short of cores)
    const int host_rows_mode = g_host_rows.load() < 0 ? auto_host_rows() : g_host_rows.load();
    const bool host_rows = host_rows_mode != 0;
    // mode 2: Z[N-k] = conj(Z[k]) is mirrored by the host threads too (8 KB of the remaining 20.7 KB per frame at
    // bufferSize 2048 stay off PCIe); not for exact-FFT plans, whose upper half is computed like the reference's
    bool host_mirror = host_rows_mode >= 2 && mb_has(p->mask, MB_FEAT_COMPLEX_SPECTRUM) && !p->dev.exact && N >= 16;
    if (host_mirror) {  // (the half rows leave as 2-D copies: into pageable arrays those crawl, 0.49 vs 0.57 M frames/s)
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, out->complex_real) != cudaSuccess || at.type != cudaMemoryTypeHost) host_mirror = false;
        (void)cudaGetLastError();
    }
    const bool host_buffer = host_rows && pcm_channels == 0 && mb_has(p->mask, MB_FEAT_BUFFER);
    // powerSpectrum[k] = float32(amplitudeSpectrum[k]^2) (powerSpectrum.js:1-7; one float32 multiply in every kernel):
    // where both are asked for, the host squares the amplitude rows as they land and the power rows stay off PCIe too.
    const bool host_power = host_rows && mb_has(p->mask, MB_FEAT_POWER_SPECTRUM) && mb_has(p->mask, MB_FEAT_AMPLITUDE_SPECTRUM);
    const uint32_t drop_mask = (host_buffer ? MB_FEATURE_BIT(MB_FEAT_BUFFER) : 0u) | (host_power ? MB_FEATURE_BIT(MB_FEAT_POWER_SPECTRUM) : 0u);
    int64_t total_frames_call = 0;
    for (int64_t i = 0; i < n_clips; i++) total_frames_call += mb_num_frames(clip_len[i], N, hop);
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    // (with the mirrored rows the host threads move 20 KB per frame instead of 12: 8 / 12 / 16 threads on a 16-core host
    // gave 2.38 / 2.56 / 2.55 M frames/s)
    const int want_workers = g_host_threads.load() > 0 ? g_host_threads.load()
                             : host_mirror           ? (int)std::min(12u, std::max(1u, hw * 3 / 4))
                                                     : (int)std::min(kDefaultHostThreads, std::max(1u, hw / 2));
    const int n_workers = (host_buffer || host_power || host_mirror) ? (int)std::min<int64_t>(std::max<int64_t>(1, total_frames_call / 2048), want_workers) : 0;
    HostWorkers workers(n_workers);  // (its destructor, on every return path, waits for the posted work: it writes into the caller's arrays)
    if (host_buffer) {
        auto fstart = std::make_shared<std::vector<int64_t>>(n_clips + 1, 0);
        for (int64_t i = 0; i < n_clips; i++) (*fstart)[i + 1] = (*fstart)[i] + mb_num_frames(clip_len[i], N, hop);
        const float *src = (const float *)samples;
        float *dst = out->buffer;
        workers.post_range(0, total_frames_call, 2048, [=](int64_t g0, int64_t g1) {
            const std::vector<int64_t> &fs = *fstart;
            int64_t c = std::upper_bound(fs.begin(), fs.end(), g0) - fs.begin() - 1;
            for (int64_t g = g0; g < g1; g++) {
                while (g >= fs[c + 1]) c++;
                copy_row_stream(dst + g * N, src + clip_offset[c] + (g - fs[c]) * hop, N);
            }
            stream_fence();
        });
    }
    // frames [g0, g1) of the amplitude rows land when `landed` completes: square them into the power rows then
    auto post_power = [&](int64_t g0, int64_t g1, cudaEvent_t landed) {
        if (!host_power || g1 <= g0) return;
        const float *amp = out->amplitude_spectrum;
        float *pw = out->power_spectrum;
        const int64_t M = N / 2;
        workers.post_range(g0, g1, 1024, [=](int64_t a, int64_t b) {
            square_rows_stream(pw + a * M, amp + a * M, (b - a) * M);
            stream_fence();
        }, landed);
    };
    // `list` (the chunk's copy of d_fix, landed before `landed`): the frames the exact kernel redid are left alone --
    // their upper halves come from the device (settle)
    auto post_mirror = [&](int64_t g0, int64_t g1, cudaEvent_t landed, const int *list) {
        float *re = out->complex_real, *im = out->complex_imag;
        workers.post_range(g0, g1, 512, [=](int64_t a, int64_t b) {
            std::vector<char> skip;
            if (list && list[0] > 0) {
                skip.assign((size_t)(b - a), 0);
                const int cnt = (int)std::min<int64_t>(list[0], g1 - g0);
                for (int i = 1; i <= cnt; i++) {
                    const int64_t g = g0 + list[i];
                    if (g >= a && g < b) skip[(size_t)(g - a)] = 1;
                }
            }
            for (int64_t g = a; g < b; g++) {
                if (!skip.empty() && skip[(size_t)(g - a)]) continue;
                mirror_row_stream(re + g * N, N, false);
                mirror_row_stream(im + g * N, N, true);
            }
            stream_fence();
        }, landed);
    };
    // a drained slot: its staged small fields handed out; the upper halves of the frames the exact kernel redid --
    // gathered on the device behind the kernels -- are copied as one piece, queued on the slot's stream (asynchronous:
    // the slot's next kernels follow in stream order), and handed out at the slot's next drain
    auto hand_out = [&](Slot &s) {
        if (s.gat_cnt == 0) return;
        (void)cudaEventSynchronize(s.gat_ev);
        const size_t up = (size_t)(N / 2 - 1), off = (size_t)(N / 2 + 1);
        const float *src = s.h_gather[s.gat_from];
        for (int i = 0; i < s.gat_cnt; i++) {
            const int f = s.gat_list[1 + i];
            if (f < 0) continue;
            memcpy(out->complex_real + (s.gat_g0 + f) * N + off, src + (size_t)i * 2 * up, up * 4);
            memcpy(out->complex_imag + (s.gat_g0 + f) * N + off, src + (size_t)i * 2 * up + up, up * 4);
        }
        s.gat_cnt = 0;
    };
    auto settle = [&](Slot &s) -> mb_status {
        scatter_small(p, s, out);
        hand_out(s);
        if (s.mir_frames > 0 && s.mir_list) {
            const int cnt = (int)std::min<int64_t>(s.mir_list[0], s.mir_frames);
            if (cnt > 0) {
                const size_t up = (size_t)(N / 2 - 1) * 4, off = (size_t)(N / 2 + 1);
                if ((size_t)cnt > s.gather_cap) {
                    // tonal material, redone almost entirely: every upper half of the chunk from the device (the rows the
                    // host threads mirrored meanwhile hold the same bits), and the chunks that follow copy whole rows
                    host_mirror = false;
                    float *hre = out->complex_real + s.small_g0 * N + off, *him = out->complex_imag + s.small_g0 * N + off;
                    MB_CUDA(cudaMemcpy2DAsync(hre, (size_t)N * 4, s.mir_d_re + off, (size_t)N * 4, up, (size_t)s.mir_frames, cudaMemcpyDeviceToHost, s.stream));
                    MB_CUDA(cudaMemcpy2DAsync(him, (size_t)N * 4, s.mir_d_im + off, (size_t)N * 4, up, (size_t)s.mir_frames, cudaMemcpyDeviceToHost, s.stream));
                } else {
                    // (the gather kernel has finished: the slot's stream was synchronized before this call)
                    MB_CUDA(cudaMemcpyAsync(s.h_gather[s.gat_t], s.d_gather[s.gat_t], (size_t)cnt * 2 * up, cudaMemcpyDeviceToHost, p->aux_stream));
                    MB_CUDA(cudaEventRecord(s.gat_ev, p->aux_stream));
                    s.gat_list = s.mir_list;
                    s.gat_cnt = cnt;
                    s.gat_from = s.gat_t;
                    s.gat_g0 = s.small_g0;
                }
            }
        }
        s.mir_frames = 0;
        s.mir_list = nullptr;
        return MB_OK;
    };
    if (host_mirror && p->adaptive) {
        const size_t need = 2 * (size_t)total_frames_call + 16;  // every chunk: its frames + 1
        if (p->fixlist_cap < need) {
            if (p->h_fixlist) cudaFreeHost(p->h_fixlist);
            p->h_fixlist = nullptr;
            p->fixlist_cap = 0;
            MB_CUDA(cudaMallocHost((void **)&p->h_fixlist, need * sizeof(int)));
            p->fixlist_cap = need;
        }
    }
    const int64_t dev_bpf = p->bytes_per_frame - (host_buffer ? 4 * (int64_t)N : 0) - (host_power ? 2 * (int64_t)N : 0);  // bytes per frame the device produces
    struct VClip { int64_t off, frames; };
    std::vector<VClip> v;
    p->refined_frames = 0;
    p->refined_on_device = false;
    for (auto &s : p->slots) s.h_fix_used = 0;
    int64_t c = 0, f_in_clip = 0, g_done = 0;
    int chunk_idx = 0;
    mb_status st = MB_OK;
    while (c < n_clips) {
        // gather virtual clips for this chunk
        v.clear();
        int64_t frames = 0, lo = INT64_MAX, hi = 0;
        while (c < n_clips && frames < chunk_frames) {
            const int64_t nf = mb_num_frames(clip_len[c], N, hop);
            if (f_in_clip >= nf) { c++; f_in_clip = 0; continue; }
            const int64_t take = std::min(nf - f_in_clip, chunk_frames - frames);
            const int64_t off = clip_offset[c] + f_in_clip * hop;
            // The chunk's samples are staged as ONE contiguous span of the host array.  Clips that lie far apart
            // (a shuffled or repeated clip list, many short clips with hop << bufferSize) would stretch that span
            // towards the whole array: close the chunk instead when the span outgrows twice the input budget.
            const int64_t nlo = std::min(lo, off), nhi = std::max(hi, off + (take - 1) * hop + N);
            if (!v.empty() && (nhi - nlo) > 2 * (chunk_frames * (int64_t)hop + N)) break;
            v.push_back({off, take});
            lo = nlo;
            hi = nhi;
            frames += take;
            f_in_clip += take;
        }
        if (frames == 0) break;
        lo &= ~(int64_t)7;  // keep the device copy's 16-byte phase equal to the host array's (float32 and mono int16)
        Slot &s = p->slots[chunk_idx & 1];
        if (!s.stream) MB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        // the slot's previous chunk (two chunks ago) must have drained before its buffers are reused
        MB_CUDA(cudaStreamSynchronize(s.stream));
        st = settle(s);  // (the other slot's copies are in flight meanwhile)
        if (st != MB_OK) return st;
        const size_t span = (size_t)(hi - lo);
        if (s.samples_cap < span * frame_bytes) {
            cudaFree(s.d_samples);
            s.d_samples = nullptr;
            s.samples_cap = 0;
            MB_CUDA(cudaMalloc((void **)&s.d_samples, span * frame_bytes));
            s.samples_cap = span * frame_bytes;
        }
        const size_t out_bytes = std::max<size_t>(16, (size_t)frames * (size_t)dev_bpf);
        if (s.out_cap < out_bytes) {
            cudaFree(s.d_out);
            s.d_out = nullptr;
            s.out_cap = 0;
            MB_CUDA(cudaMalloc((void **)&s.d_out, out_bytes));
            s.out_cap = out_bytes;
        }
        const size_t entries = 2 * v.size() + 1;
        st = ensure_table(&s.d_tab, &s.h_tab, &s.tab_cap, entries);
        if (st != MB_OK) return st;
        int64_t acc = 0;
        for (size_t i = 0; i < v.size(); i++) {
            s.h_tab[i] = v[i].off - lo;
            s.h_tab[v.size() + i] = acc;
            acc += v[i].frames;
        }
        s.h_tab[2 * v.size()] = acc;
        MB_CUDA(cudaMemcpyAsync(s.d_tab, s.h_tab, entries * sizeof(int64_t), cudaMemcpyHostToDevice, s.stream));
        MB_CUDA(cudaMemcpyAsync(s.d_samples, (const char *)samples + (size_t)lo * frame_bytes, span * frame_bytes,
                                cudaMemcpyHostToDevice, s.stream));
        // carve the slot's output arena: the big arrays first, then the per-frame numbers and short arrays back to back
        // (they leave as one copy)
        mb_outputs d_out;
        memset(&d_out, 0, sizeof(d_out));
        size_t cursor = 0, small_begin = 0;
        for (int small = 0; small < 2; small++) {
            if (small) small_begin = cursor;
            for (int i = 0; i < kNumFields; i++) {
                if (!mb_has(p->mask & ~drop_mask, kFields[i].feature)) continue;
                if ((kFields[i].kind == 1 || kFields[i].kind == 2) == (small == 1)) continue;
                field_ptr(d_out, kFields[i]) = s.d_out + cursor;
                if (small) s.small_parts.push_back({i, cursor - small_begin});
                cursor += (size_t)frames * field_elems(kFields[i], p->dev) * 4;
            }
        }
        const size_t small_bytes = cursor - small_begin;
        if (s.small_cap < small_bytes) {
            if (s.h_small) cudaFreeHost(s.h_small);
            s.h_small = nullptr;
            s.small_cap = 0;
            MB_CUDA(cudaMallocHost((void **)&s.h_small, small_bytes));
            s.small_cap = small_bytes;
        }
        s.small_g0 = g_done;
        s.small_frames = frames;
        st = launch(p, s.d_tab, s.d_tab + v.size(), (int64_t)v.size(), frames, s.d_samples, d_out, s.stream, &s.d_fix,
                    &s.fix_cap, pcm_channels, pcm_channel, pcm_format, drop_mask);
        if (st != MB_OK) return st;
        if (p->adaptive) {  // how many frames were redone: read back with the outputs, summed after the last chunk
            if (s.h_fix_used == s.h_fix_cap) {
                MB_CUDA(cudaStreamSynchronize(s.stream));
                for (size_t i = 0; i < s.h_fix_used; i++) p->refined_frames += s.h_fix[i];
                s.h_fix_used = 0;
                if (!s.h_fix) {
                    MB_CUDA(cudaMallocHost((void **)&s.h_fix, 256 * sizeof(int)));
                    s.h_fix_cap = 256;
                }
            }
            MB_CUDA(cudaMemcpyAsync(s.h_fix + s.h_fix_used++, s.d_fix, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        // which frames the exact kernel redid, ahead of the rows (mode 2: the host threads leave those frames alone)
        const int *chunk_list = nullptr;
        if (host_mirror && p->adaptive && s.d_fix) {
            int *dst = p->h_fixlist + g_done + chunk_idx;  // (regions of frames + 1 ints, one after the other)
            MB_CUDA(cudaMemcpyAsync(dst, s.d_fix, (size_t)(frames + 1) * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            chunk_list = dst;
            // the redone frames' upper halves, packed (up to an eighth of the chunk: beyond that whole rows are copied)
            const size_t gcap = (size_t)std::max<int64_t>(8, frames / 8), gbytes = gcap * 2 * (size_t)(N / 2 - 1) * 4;
            if (!p->aux_stream) MB_CUDA(cudaStreamCreateWithFlags(&p->aux_stream, cudaStreamNonBlocking));
            if (!s.gat_ev) MB_CUDA(cudaEventCreateWithFlags(&s.gat_ev, cudaEventDisableTiming));
            if (s.gather_cap < gcap) {
                hand_out(s);  // (a copy into the old area may still be pending: finish and hand it out first)
                for (int t = 0; t < 2; t++) {
                    cudaFree(s.d_gather[t]);
                    if (s.h_gather[t]) cudaFreeHost(s.h_gather[t]);
                    s.d_gather[t] = s.h_gather[t] = nullptr;
                }
                s.gather_cap = 0;
                for (int t = 0; t < 2; t++) {
                    MB_CUDA(cudaMalloc((void **)&s.d_gather[t], gbytes));
                    MB_CUDA(cudaMallocHost((void **)&s.h_gather[t], gbytes));
                }
                s.gather_cap = gcap;
            }
            s.gat_t ^= 1;
            mb_gather_upper_kernel<<<(unsigned)s.gather_cap, 128, 0, s.stream>>>(s.d_fix, (int)s.gather_cap, d_out.complex_real, d_out.complex_imag, N, s.d_gather[s.gat_t]);
            MB_CUDA(cudaGetLastError());
        }
        // the amplitude rows first when the host squares them: that work then overlaps the rest of the chunk's copies
        // (with the mirrored rows on the host too, complexSpectrum goes first -- twice the host work hangs on it -- and
        // the amplitude rows last)
        auto record = [&](std::vector<cudaEvent_t> &evs) -> mb_status {
            if ((size_t)chunk_idx >= evs.size()) {  // (one event per chunk of the call: a worker may still be waiting on an earlier one)
                cudaEvent_t ev = nullptr;
                MB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                evs.push_back(ev);
            }
            MB_CUDA(cudaEventRecord(evs[chunk_idx], s.stream));
            return MB_OK;
        };
        for (int pass = (host_power || host_mirror) ? 0 : 1; pass < 2; pass++) {
            for (int i = 0; i < kNumFields; i++) {
                if (!mb_has(p->mask & ~drop_mask, kFields[i].feature)) continue;
                if (kFields[i].kind != 1 && kFields[i].kind != 2) continue;  // (staged: below)
                const bool is_amp = kFields[i].feature == MB_FEAT_AMPLITUDE_SPECTRUM;
                const bool first = host_mirror ? kFields[i].feature == MB_FEAT_COMPLEX_SPECTRUM : is_amp;
                if ((host_power || host_mirror) && first != (pass == 0)) continue;
                const size_t per = (size_t)field_elems(kFields[i], p->dev) * 4;
                if (host_mirror && kFields[i].feature == MB_FEAT_COMPLEX_SPECTRUM)  // bins 0 .. N/2 of every row
                    MB_CUDA(cudaMemcpy2DAsync((char *)field_ptr(*out, kFields[i]) + (size_t)g_done * per, per, field_ptr(d_out, kFields[i]),
                                              per, (size_t)(N / 2 + 1) * 4, (size_t)frames, cudaMemcpyDeviceToHost, s.stream));
                else
                    MB_CUDA(cudaMemcpyAsync((char *)field_ptr(*out, kFields[i]) + (size_t)g_done * per,
                                            field_ptr(d_out, kFields[i]), (size_t)frames * per, cudaMemcpyDeviceToHost,
                                            s.stream));
            }
            if (pass == 0 && host_mirror) {
                st = record(p->landed_c);
                if (st != MB_OK) return st;
                post_mirror(g_done, g_done + frames, p->landed_c[chunk_idx], chunk_list);
                s.mir_list = chunk_list;
                s.mir_frames = frames;
                s.mir_d_re = d_out.complex_real;
                s.mir_d_im = d_out.complex_imag;
            }
            if (host_power && pass == (host_mirror ? 1 : 0)) {  // (mode 2: the amplitude rows are the last big copy of the chunk)
                st = record(p->landed);
                if (st != MB_OK) return st;
                post_power(g_done, g_done + frames, p->landed[chunk_idx]);
            }
        }
        if (small_bytes) MB_CUDA(cudaMemcpyAsync(s.h_small, s.d_out + small_begin, small_bytes, cudaMemcpyDeviceToHost, s.stream));
        g_done += frames;
        chunk_idx++;
    }
    for (int k = 0; k < 2; k++) {
        Slot &s = p->slots[k];
        if (s.stream) MB_CUDA(cudaStreamSynchronize(s.stream));
        const bool more = s.mir_frames > 0 && s.mir_list && s.mir_list[0] > 0;
        st = settle(s);
        if (st != MB_OK) return st;
        if (more) {  // (the redone frames' upper halves, queued by settle)
            MB_CUDA(cudaStreamSynchronize(s.stream));
            hand_out(s);
        }
        for (size_t i = 0; i < s.h_fix_used; i++) p->refined_frames += s.h_fix[i];
        s.h_fix_used = 0;
    }
    return MB_OK;  // (`workers` joins here: every posted row is written before the call returns)
}

mb_status mb_extract(mb_plan *p, const float *samples, int64_t n_samples, const int64_t *clip_offset,
                     const int64_t *clip_len, int64_t n_clips, const mb_outputs *out, int mem_kind) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    if (mem_kind == MB_MEM_DEVICE) {
        mb_status st = mb_extract_async(p, samples, n_samples, clip_offset, clip_len, n_clips, out);
        if (st != MB_OK) return st;
        return mb_plan_synchronize(p);
    }
    if (mem_kind != MB_MEM_HOST) return fail(MB_ERR_INVALID_ARG, "unknown memory kind %d", mem_kind);
    mb_status st = check_outputs(p, out);
    if (st != MB_OK) return st;
    st = check_clips(p, n_samples, clip_offset, clip_len, n_clips);
    if (st != MB_OK) return st;
    if (n_clips == 0) return MB_OK;
    if (!samples) return fail(MB_ERR_INVALID_ARG, "samples is NULL");
    return extract_host(p, samples, clip_offset, clip_len, n_clips, out);
}


mb_status mb_extract_pcm16(mb_plan *p, const int16_t *pcm, int64_t n_sample_frames, int channels, int channel,
                           const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips, const mb_outputs *out,
                           int mem_kind) {
    return mb_extract_pcm(p, pcm, MB_SAMPLE_S16, n_sample_frames, channels, channel, clip_offset, clip_len, n_clips, out,
                          mem_kind);
}

mb_status mb_extract_pcm(mb_plan *p, const void *pcm, int format, int64_t n_sample_frames, int channels, int channel,
                         const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips, const mb_outputs *out,
                         int mem_kind) {
    if (!p) return fail(MB_ERR_INVALID_ARG, "plan is NULL");
    if (format != MB_SAMPLE_S16 && format != MB_SAMPLE_S24 && format != MB_SAMPLE_F32)
        return fail(MB_ERR_INVALID_ARG, "unknown sample format %d", format);
    mb_status st = check_pcm(channels, channel);
    if (st != MB_OK) return st;
    if (mem_kind == MB_MEM_DEVICE) {
        st = extract_device(p, reinterpret_cast<const float *>(pcm), n_sample_frames, clip_offset, clip_len, n_clips, out,
                            channels, channel, format);
        if (st != MB_OK) return st;
        return mb_plan_synchronize(p);
    }
    if (mem_kind != MB_MEM_HOST) return fail(MB_ERR_INVALID_ARG, "unknown memory kind %d", mem_kind);
    st = check_outputs(p, out);
    if (st != MB_OK) return st;
    st = check_clips(p, n_sample_frames, clip_offset, clip_len, n_clips);
    if (st != MB_OK) return st;
    if (n_clips == 0) return MB_OK;
    if (!pcm) return fail(MB_ERR_INVALID_ARG, "pcm is NULL");
    return extract_host(p, pcm, clip_offset, clip_len, n_clips, out, channels, channel, format);
}

// RIFF/WAVE: "RIFF" <size> "WAVE" then chunks <id> <size> <payload, padded to even>; needs "fmt " and "data".
mb_status mb_wav_parse(const void *file_bytes, int64_t n_bytes, mb_wav_info *info) {
    if (!file_bytes || !info || n_bytes < 12) return fail(MB_ERR_INVALID_ARG, "not a RIFF/WAVE file (too short)");
    const unsigned char *b = (const unsigned char *)file_bytes;
    auto u16 = [&](int64_t o) { return (uint32_t)b[o] | ((uint32_t)b[o + 1] << 8); };
    auto u32 = [&](int64_t o) { return u16(o) | (u16(o + 2) << 16); };
    if (memcmp(b, "RIFF", 4) != 0 || memcmp(b + 8, "WAVE", 4) != 0)
        return fail(MB_ERR_INVALID_ARG, "not a RIFF/WAVE file (bad magic)");
    memset(info, 0, sizeof(*info));
    bool have_fmt = false, have_data = false;
    int block_align = 0;
    int64_t data_bytes = 0;
    for (int64_t o = 12; o + 8 <= n_bytes;) {
        const int64_t size = u32(o + 4), body = o + 8;
        if (memcmp(b + o, "fmt ", 4) == 0) {
            if (size < 16 || body + 16 > n_bytes) return fail(MB_ERR_INVALID_ARG, "truncated fmt chunk");
            info->format = (int32_t)u16(body);
            info->channels = (int32_t)u16(body + 2);
            info->sample_rate = (int32_t)u32(body + 4);
            block_align = (int)u16(body + 12);
            info->bits_per_sample = (int32_t)u16(body + 14);
            if (info->format == 0xFFFE && size >= 26 && body + 26 <= n_bytes) info->format = (int32_t)u16(body + 24);  // WAVE_FORMAT_EXTENSIBLE: sub-format
            have_fmt = true;
        } else if (memcmp(b + o, "data", 4) == 0) {
            info->data_offset = body;
            data_bytes = std::min<int64_t>(size, n_bytes - body);
            have_data = true;
            break;
        }
        o = body + size + (size & 1);
    }
    if (!have_fmt || !have_data) return fail(MB_ERR_INVALID_ARG, "WAVE file without a fmt or data chunk");
    if (info->channels < 1 || info->bits_per_sample < 8 || info->bits_per_sample % 8 != 0)
        return fail(MB_ERR_INVALID_ARG, "bad fmt chunk");
    const int expect_align = info->channels * (info->bits_per_sample / 8);
    if (block_align <= 0) block_align = expect_align;
    if (block_align != expect_align)  // a frame count derived from a lying block size would point past the data
        return fail(MB_ERR_INVALID_ARG, "fmt chunk: block align %d is not channels x bytes per sample (%d)", block_align,
                    expect_align);
    info->n_sample_frames = data_bytes / block_align;
    return MB_OK;
}

mb_status mb_extract_multi(mb_plan *const *plans, int n_plans, const float *samples, int64_t n_samples,
                           const int64_t *clip_offset, const int64_t *clip_len, int64_t n_clips,
                           const mb_outputs *out) {
    if (!plans || n_plans <= 0) return fail(MB_ERR_INVALID_ARG, "no plans given");
    for (int i = 0; i < n_plans; i++) {
        if (!plans[i]) return fail(MB_ERR_INVALID_ARG, "plan %d is NULL", i);
        if (plans[i]->N != plans[0]->N || plans[i]->hop != plans[0]->hop || plans[i]->mask != plans[0]->mask ||
            plans[i]->sr != plans[0]->sr || plans[i]->window != plans[0]->window ||
            plans[i]->dev.nb != plans[0]->dev.nb || plans[i]->dev.nf != plans[0]->dev.nf ||
            plans[i]->dev.nc != plans[0]->dev.nc || plans[i]->dev.rolloff_frac != plans[0]->dev.rolloff_frac)
            return fail(MB_ERR_INVALID_ARG, "plan %d was created with different parameters than plan 0", i);
    }
    mb_plan *p0 = plans[0];
    mb_status st = check_outputs(p0, out);
    if (st != MB_OK) return st;
    st = check_clips(p0, n_samples, clip_offset, clip_len, n_clips);
    if (st != MB_OK) return st;
    if (n_clips == 0) return MB_OK;
    // contiguous clip ranges balanced on cumulative frame count
    std::vector<int64_t> prefix(n_clips + 1, 0);
    for (int64_t c = 0; c < n_clips; c++) prefix[c + 1] = prefix[c] + mb_num_frames(clip_len[c], p0->N, p0->hop);
    const int64_t total = prefix[n_clips];
    std::vector<int64_t> cut(n_plans + 1, n_clips);
    cut[0] = 0;
    for (int d = 1; d < n_plans; d++) {
        const int64_t target = total * d / n_plans;
        cut[d] = std::lower_bound(prefix.begin(), prefix.end(), target) - prefix.begin();
        cut[d] = std::min<int64_t>(std::max(cut[d], cut[d - 1]), n_clips);
    }
    std::vector<mb_status> status(n_plans, MB_OK);
    std::vector<std::string> msgs(n_plans);
    std::vector<std::thread> th;
    for (int d = 0; d < n_plans; d++) {
        th.emplace_back([&, d]() {
            const int64_t c0 = cut[d], c1 = cut[d + 1];
            if (c1 <= c0) return;
            mb_outputs o;
            offset_outputs(o, *out, prefix[c0], p0->dev);
            status[d] = mb_extract(plans[d], samples, n_samples, clip_offset + c0, clip_len + c0, c1 - c0, &o,
                                   MB_MEM_HOST);
            if (status[d] != MB_OK) msgs[d] = mb_last_error();
        });
    }
    for (auto &t : th) t.join();
    for (int d = 0; d < n_plans; d++)
        if (status[d] != MB_OK) return fail(status[d], "device shard %d: %s", d, msgs[d].c_str());
    return MB_OK;
}

// ---- measured non-tensor arithmetic peaks of a device (the FP32 / FP64 roofline denominators bench.py reports against)
mb_status mb_set_host_rows(int mode) {
    if (mode < -1 || mode > 2) return fail(MB_ERR_INVALID_ARG, "host rows mode %d (2 all, 1 buffer + powerSpectrum, 0 off, -1 default)", mode);
    g_host_rows.store(mode);
    return MB_OK;
}

int mb_get_host_rows(void) { return g_host_rows.load() < 0 ? auto_host_rows() : g_host_rows.load(); }

mb_status mb_set_host_threads(int n) {
    if (n < 0 || n > 256) return fail(MB_ERR_INVALID_ARG, "host thread count %d out of range [0, 256]", n);
    g_host_threads.store(n);
    return MB_OK;
}

mb_status mb_measure_peaks(int device, double *fp32_ffma_tflops, double *fp64_dfma_tflops) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return fail(MB_ERR_NO_DEVICE, "device %d out of range", device);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    MB_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    void *buf = nullptr;
    MB_CUDA(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto best_of = [&](bool dbl, int iters) {
        float best = 1e30f;
        for (int r = 0; r < 6; r++) {  // (the first run warms up)
            cudaEventRecord(e0);
            if (dbl) mb_fma_peak_kernel<double><<<blocks, threads>>>((double *)buf, iters);
            else mb_fma_peak_kernel<float><<<blocks, threads>>>((float *)buf, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        return 2.0 * 8 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    };
    if (fp32_ffma_tflops) *fp32_ffma_tflops = best_of(false, 1 << 15);
    if (fp64_dfma_tflops) *fp64_dfma_tflops = best_of(true, 1 << 13);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    MB_CUDA(cudaGetLastError());
    return MB_OK;
}

int64_t mb_plan_launch_count(const mb_plan *p) { return p ? p->launches : 0; }

mb_status mb_plan_refined_frames(mb_plan *p, int64_t *frames) {
    if (!p || !frames) return fail(MB_ERR_INVALID_ARG, "plan/frames is NULL");
    *frames = 0;
    if (!p->adaptive) return MB_OK;
    if (p->refined_on_device && p->d_fix) {  // device-memory call (or stream push): the count lies next to the list
        DeviceGuard guard(p->device);
        int n = 0;
        MB_CUDA(cudaStreamSynchronize(p->stream));
        MB_CUDA(cudaMemcpy(&n, p->d_fix, sizeof(int), cudaMemcpyDeviceToHost));
        *frames = n;
    } else {
        *frames = p->refined_frames;
    }
    return MB_OK;
}
const char *mb_plan_kernel_name(const mb_plan *p) { return p ? p->kernel_name : ""; }

mb_status mb_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(MB_ERR_INVALID_ARG, "ptr is NULL");
    *ptr = nullptr;
    MB_CUDA(cudaMallocHost(ptr, std::max<size_t>(bytes, 1)));
    return MB_OK;
}
void mb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

// ---- streaming (buffer-by-buffer) use, src/meyda.js:69-91

mb_status mb_stream_create(mb_stream **stream, mb_plan *plan) {
    if (!stream || !plan) return fail(MB_ERR_INVALID_ARG, "stream/plan is NULL");
    *stream = new mb_stream();
    (*stream)->plan = plan;
    return MB_OK;
}

void mb_stream_destroy(mb_stream *s) {
    if (!s) return;
    DeviceGuard guard(s->plan->device);
    cudaStreamSynchronize(s->plan->stream);
    for (auto &g : s->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(s->d_buf[0]);
    cudaFree(s->d_buf[1]);
    cudaFree(s->d_out);
    cudaFree(s->d_tab);
    if (s->h_in) cudaFreeHost(s->h_in);
    if (s->h_out) cudaFreeHost(s->h_out);
    if (s->h_tab) cudaFreeHost(s->h_tab);
    delete s;
}

int64_t mb_stream_graph_launches(const mb_stream *s) { return s ? s->graph_launches : 0; }

int64_t mb_stream_frames_after(const mb_stream *s, int64_t n_new) {
    if (!s || n_new < 0) return 0;
    const int64_t usable = n_new > s->skip ? n_new - s->skip : 0;
    return mb_num_frames(s->filled + usable, s->plan->N, s->plan->hop);
}

mb_status mb_stream_reset(mb_stream *s) {
    if (!s) return fail(MB_ERR_INVALID_ARG, "stream is NULL");
    s->filled = 0;
    s->skip = 0;
    return MB_OK;
}

static mb_status stream_push(mb_stream *s, const char *new_samples, int64_t n_new, const mb_outputs *out, int mem_kind,
                             int64_t *frames_done) {
    if (n_new < 0 || (n_new > 0 && !new_samples)) return fail(MB_ERR_INVALID_ARG, "bad sample block");
    const size_t fb = s->frame_bytes;
    if (mem_kind != MB_MEM_HOST && mem_kind != MB_MEM_DEVICE) return fail(MB_ERR_INVALID_ARG, "unknown memory kind");
    mb_plan *p = s->plan;
    DeviceGuard guard(p->device);
    if (frames_done) *frames_done = 0;
    {
        const int64_t drop = std::min(s->skip, n_new);
        new_samples += (size_t)drop * fb;
        n_new -= drop;
        s->skip -= drop;
    }
    const int64_t need = s->filled + n_new;
    if ((int64_t)s->cap < need) {
        size_t cap = (size_t)need + (size_t)need / 2 + p->N;
        char *nb[2] = {nullptr, nullptr};
        MB_CUDA(cudaMalloc((void **)&nb[0], cap * fb));
        MB_CUDA(cudaMalloc((void **)&nb[1], cap * fb));
        if (s->filled)
            MB_CUDA(cudaMemcpyAsync(nb[0], s->d_buf[s->cur], s->filled * fb, cudaMemcpyDeviceToDevice, p->stream));
        MB_CUDA(cudaStreamSynchronize(p->stream));
        cudaFree(s->d_buf[0]);
        cudaFree(s->d_buf[1]);
        s->d_buf[0] = nb[0];
        s->d_buf[1] = nb[1];
        s->cur = 0;
        s->cap = cap;
        for (auto &g : s->graphs) g.cur = -1;  // captured pushes point into the old buffers
    }
    char *buf = s->d_buf[s->cur];
    const int64_t filled_before = s->filled;
    const int64_t nf = mb_num_frames(need, p->N, p->hop);
    if (nf > 0) {
        mb_status st = check_outputs(p, out);
        if (st != MB_OK) return st;
    }
    if (nf > 0) {
        p->refined_on_device = p->adaptive;
        p->refined_frames = 0;
    }
    const int64_t consumed = nf * p->hop;  // the hop-overlap tail [consumed, need) stays for the next push
    const int64_t rest = std::max<int64_t>(0, need - consumed);

    if (mem_kind == MB_MEM_DEVICE) {
        if (n_new)
            MB_CUDA(cudaMemcpyAsync(buf + filled_before * fb, new_samples, n_new * fb, cudaMemcpyDeviceToDevice, p->stream));
        if (nf > 0) {
            const int64_t off = 0, len = need;
            mb_status st = extract_device(p, reinterpret_cast<const float *>(buf), need, &off, &len, 1, out, s->pcm_channels,
                                          s->pcm_channel);
            if (st != MB_OK) return st;
            if (rest)
                MB_CUDA(cudaMemcpyAsync(s->d_buf[s->cur ^ 1], buf + consumed * fb, rest * fb, cudaMemcpyDeviceToDevice,
                                        p->stream));
        }
        MB_CUDA(cudaStreamSynchronize(p->stream));
    } else {
        // Host memory, the latency path (one buffer per call, as onaudioprocess delivers them): the block goes
        // through pinned staging, every output leaves in ONE device-to-host copy of a packed arena, and the
        // whole push (copy in, kernel, copy out, tail move) is replayed as a CUDA graph once its shape
        // (buffer parity, fill level, block size) has been seen before.
        const size_t out_bytes = (size_t)nf * (size_t)p->bytes_per_frame;
        if (s->h_in_cap < (size_t)n_new) {
            if (s->h_in) cudaFreeHost(s->h_in);
            s->h_in = nullptr;
            s->h_in_cap = 0;
            MB_CUDA(cudaMallocHost((void **)&s->h_in, ((size_t)n_new + (size_t)n_new / 2 + 64) * fb));
            s->h_in_cap = (size_t)n_new + (size_t)n_new / 2 + 64;
            for (auto &g : s->graphs) g.cur = -1;  // captured pointers are stale
        }
        if (s->out_cap < out_bytes) {
            MB_CUDA(cudaStreamSynchronize(p->stream));
            cudaFree(s->d_out);
            if (s->h_out) cudaFreeHost(s->h_out);
            s->d_out = s->h_out = nullptr;
            s->out_cap = 0;
            MB_CUDA(cudaMalloc((void **)&s->d_out, out_bytes + out_bytes / 2));
            MB_CUDA(cudaMallocHost((void **)&s->h_out, out_bytes + out_bytes / 2));
            s->out_cap = out_bytes + out_bytes / 2;
            for (auto &g : s->graphs) g.cur = -1;
        }
        if (p->adaptive && p->fix_cap < (size_t)nf + 1) {  // (never inside a capture: the graphs hold this pointer)
            MB_CUDA(cudaStreamSynchronize(p->stream));
            cudaFree(p->d_fix);
            p->d_fix = nullptr;
            p->fix_cap = 0;
            const size_t want = std::max<size_t>(4096, 2 * (size_t)nf + 64);
            MB_CUDA(cudaMalloc((void **)&p->d_fix, want * sizeof(int)));
            p->fix_cap = want;
            for (auto &g : s->graphs) g.cur = -1;
        }
        if (!s->d_tab) {
            MB_CUDA(cudaMalloc((void **)&s->d_tab, 3 * sizeof(int64_t)));
            MB_CUDA(cudaMallocHost((void **)&s->h_tab, 3 * sizeof(int64_t)));
        }
        if (nf > 0 && s->tab_nf != nf) {  // clip table of the single clip [0, need): only its frame count varies
            MB_CUDA(cudaStreamSynchronize(p->stream));
            s->h_tab[0] = 0;
            s->h_tab[1] = 0;
            s->h_tab[2] = nf;
            MB_CUDA(cudaMemcpyAsync(s->d_tab, s->h_tab, 3 * sizeof(int64_t), cudaMemcpyHostToDevice, p->stream));
            s->tab_nf = nf;
        }
        if (n_new) memcpy(s->h_in, new_samples, (size_t)n_new * fb);
        mb_outputs d_out;
        memset(&d_out, 0, sizeof(d_out));
        size_t cursor = 0;
        for (int i = 0; i < kNumFields; i++) {
            if (!mb_has(p->mask, kFields[i].feature)) continue;
            field_ptr(d_out, kFields[i]) = s->d_out + cursor;
            cursor += (size_t)nf * field_elems(kFields[i], p->dev) * 4;
        }
        auto enqueue = [&]() -> mb_status {
            if (n_new)
                MB_CUDA(cudaMemcpyAsync(buf + filled_before * fb, s->h_in, n_new * fb, cudaMemcpyHostToDevice, p->stream));
            if (nf > 0) {
                mb_status st = launch(p, s->d_tab, s->d_tab + 1, 1, nf, reinterpret_cast<const float *>(buf), d_out, p->stream,
                                      &p->d_fix, &p->fix_cap, s->pcm_channels, s->pcm_channel);
                if (st != MB_OK) return st;
                MB_CUDA(cudaMemcpyAsync(s->h_out, s->d_out, out_bytes, cudaMemcpyDeviceToHost, p->stream));
                if (rest)
                    MB_CUDA(cudaMemcpyAsync(s->d_buf[s->cur ^ 1], buf + consumed * fb, rest * fb, cudaMemcpyDeviceToDevice,
                                            p->stream));
            }
            return MB_OK;
        };
        StreamGraph *slot_g = nullptr;
        if (nf > 0 && p->stream == p->own_stream) {
            for (auto &g : s->graphs)
                if (g.cur == s->cur && g.filled == filled_before && g.n_new == n_new) slot_g = &g;
            if (!slot_g) {  // take the least used entry for this new shape
                slot_g = &s->graphs[0];
                for (auto &g : s->graphs)
                    if (g.cur < 0 || g.seen < slot_g->seen) { slot_g = &g; if (g.cur < 0) break; }
                if (slot_g->exec) cudaGraphExecDestroy(slot_g->exec);
                *slot_g = StreamGraph();
                slot_g->cur = s->cur;
                slot_g->filled = filled_before;
                slot_g->n_new = n_new;
            }
            slot_g->seen++;
        }
        if (slot_g && slot_g->exec) {
            MB_CUDA(cudaGraphLaunch(slot_g->exec, p->stream));
            p->launches++;
            p->launches_warp += p->has_warp_kernel ? 1 : 0;
            s->graph_launches++;
        } else if (slot_g && slot_g->seen >= 2) {
            cudaGraph_t graph = nullptr;
            MB_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
            mb_status st = enqueue();
            cudaError_t ce = cudaStreamEndCapture(p->stream, &graph);
            if (st != MB_OK) {
                if (graph) cudaGraphDestroy(graph);
                return st;
            }
            if (ce != cudaSuccess) return fail(MB_ERR_CUDA, "stream capture: %s", cudaGetErrorString(ce));
            ce = cudaGraphInstantiate(&slot_g->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) {
                slot_g->exec = nullptr;
                return fail(MB_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(ce));
            }
            MB_CUDA(cudaGraphLaunch(slot_g->exec, p->stream));
            s->graph_launches++;
        } else {
            mb_status st = enqueue();
            if (st != MB_OK) return st;
        }
        MB_CUDA(cudaStreamSynchronize(p->stream));
        cursor = 0;
        for (int i = 0; i < kNumFields && nf > 0; i++) {
            if (!mb_has(p->mask, kFields[i].feature)) continue;
            const size_t bytes = (size_t)nf * field_elems(kFields[i], p->dev) * 4;
            memcpy(field_ptr(*out, kFields[i]), s->h_out + cursor, bytes);
            cursor += bytes;
        }
    }
    if (nf > 0) {
        s->skip = std::max<int64_t>(0, consumed - need);
        s->cur ^= 1;
        s->filled = rest;
    } else {
        s->filled = need;
    }
    if (frames_done) *frames_done = nf;
    return MB_OK;
}

mb_status mb_stream_push(mb_stream *s, const float *new_samples, int64_t n_new, const mb_outputs *out, int mem_kind,
                         int64_t *frames_done) {
    if (!s) return fail(MB_ERR_INVALID_ARG, "stream is NULL");
    if (s->pcm_channels) return fail(MB_ERR_INVALID_ARG, "this stream takes 16-bit PCM (mb_stream_push_pcm16)");
    return stream_push(s, reinterpret_cast<const char *>(new_samples), n_new, out, mem_kind, frames_done);
}

mb_status mb_stream_create_pcm16(mb_stream **stream, mb_plan *plan, int channels, int channel) {
    mb_status st = check_pcm(channels, channel);
    if (st != MB_OK) return st;
    st = mb_stream_create(stream, plan);
    if (st != MB_OK) return st;
    (*stream)->pcm_channels = channels;
    (*stream)->pcm_channel = channel;
    (*stream)->frame_bytes = 2u * (size_t)channels;
    return MB_OK;
}

mb_status mb_stream_push_pcm16(mb_stream *s, const int16_t *new_sample_frames, int64_t n_new, const mb_outputs *out,
                               int mem_kind, int64_t *frames_done) {
    if (!s) return fail(MB_ERR_INVALID_ARG, "stream is NULL");
    if (!s->pcm_channels) return fail(MB_ERR_INVALID_ARG, "this stream takes float32 samples (mb_stream_push)");
    return stream_push(s, reinterpret_cast<const char *>(new_sample_frames), n_new, out, mem_kind, frames_done);
}

}  // extern "C"
