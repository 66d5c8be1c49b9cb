/*
 * meyda_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar, literal CPU restatement of Meyda's per-frame feature-extraction
 * path (reference: kirbysayshi/meyda v1.1.0, JavaScript).  Every function
 * cites the reference file:line it follows.  JS semantics are kept exactly:
 * all arithmetic is IEEE double, a value is rounded to float only where the
 * reference stores into a Float32Array, there is no fused multiply-add
 * (compile with -ffp-contract=off), and loops run in the reference's order.
 *
 * PINNING: the reference ships no tests / golden vectors and no JavaScript
 * engine exists in this image.  The restatement is pinned against the
 * reference's OWN source files executed by oracle/minijs.py, an ES5-subset
 * interpreter written for that purpose: tests/golden/js_reference_vectors.npz
 * (tools/make_js_golden.py) holds what the lib/jsfft sources, src/utils.js, every
 * file under src/extractors and the compute* methods of src/meyda.js return on seventeen
 * frames, and tests/test_js_pin.py requires this file to reproduce them bit
 * for bit (Float32Array results) / to 1e-12 (Numbers).  Further guards: (i) an
 * independently written numpy restatement (oracle/meyda_oracle.py) that must
 * agree bit for bit, (ii) numpy.fft / closed-form identities, (iii) the
 * survey's spot values.  See DESIGN.md "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product path
 * (meyda_b200/) never does.
 *
 * Wiring fixes relative to the mid-refactor snapshot: the FFT
 * is taken per frame on a zero-imaginary ComplexArray; `buffer` is the raw
 * signal; perceptualSpread/Sharpness use the same loudness; mfcc's
 * audioContext.sampleRate is the plan's sample rate.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MO_NUM_BARK 24
#define MO_NUM_MEL 26
#define MO_NUM_MFCC 13

enum { MO_WIN_HANNING = 0, MO_WIN_HAMMING = 1, MO_WIN_BLACKMAN = 2 };

typedef struct mo_plan {
    int N;            /* bufferSize */
    int n;            /* N / 2 = ampSpectrum.length */
    double sr;        /* audioContext.sampleRate */
    int window;
    float *hanning;   /* Float32Array(N) */
    float *hamming;   /* Float32Array(N) */
    float *blackman;  /* Float32Array(N) */
    float *bark;      /* Float32Array(N) */
    int32_t bbLimits[MO_NUM_BARK + 1];
} mo_plan;

/* Per-frame result; Numbers are JS doubles, arrays are Float32Array. */
typedef struct mo_frame_out {
    double rms, energy, zcr;
    double centroid, flatness, slope, rolloff, spread, skewness, kurtosis;
    double loudness_total, perceptual_spread, perceptual_sharpness;
    float loudness_specific[MO_NUM_BARK];
    float mfcc[MO_NUM_MFCC];
    /* caller-provided (may be NULL): */
    float *buffer;      /* N   */
    float *cs_real;     /* N   */
    float *cs_imag;     /* N   */
    float *amp;         /* N/2 */
    float *power;       /* N/2 */
} mo_frame_out;

/* src/utils.js:13-19 */
int mo_is_power_of_two(double num) {
    while (fmod(num, 2.0) == 0.0 && num > 1.0) num /= 2.0;
    return num == 1.0;
}

/* src/meyda.js:128-138 */
static void compute_hanning(float *w, int N) {
    for (int i = 0; i < N; i++)
        w[i] = (float)(0.5 - 0.5 * cos(2 * M_PI * i / (N - 1)));
}

/* src/meyda.js:116-126  (note: cos(2*PI*(i/N - 1))) */
static void compute_hamming(float *w, int N) {
    for (int i = 0; i < N; i++)
        w[i] = (float)(0.54 - 0.46 * cos(2 * M_PI * ((double)i / N - 1)));
}

/* src/meyda.js:140-156: the reference leaves this window commented out ("UNFINISHED"); what it states is the
 * formula (the symmetric window of the MathWorks page it cites), evaluated here for every i.  No reference
 * output exists for it -- this table is pinned by its closed form only. */
static void compute_blackman(float *w, int N) {
    for (int i = 0; i < N; i++)
        w[i] = (float)(0.42 - 0.5 * cos(2 * M_PI * i / (N - 1)) + 0.08 * cos(4 * M_PI * i / (N - 1)));
}

/* src/meyda.js:170-182 -- the Hz value is stored to the Float32Array first,
 * then read back for the bark formula. */
static void compute_bark_scale(float *b, int N, double sr) {
    for (int i = 0; i < N; i++) {
        b[i] = (float)(i * sr / N);
        double v = b[i];
        b[i] = (float)(13 * atan(v / 1315.8) + 3.5 * atan(pow(v / 7518, 2)));
    }
}

/* src/extractors/loudness.js:24-45 */
static void compute_bark_band_limits(int32_t *bb, const float *bark, int nSpec, int nb) {
    double currentBandEnd = (double)bark[nSpec - 1] / nb;
    int currentBand = 1;
    for (int i = 0; i <= nb; i++) bb[i] = 0; /* Int32Array zero-initialised */
    bb[0] = 0;
    for (int i = 0; i < nSpec; i++) {
        while ((double)bark[i] > currentBandEnd) {
            if (currentBand <= nb) bb[currentBand] = i; /* OOB typed-array store is a no-op */
            currentBand++;
            currentBandEnd = currentBand * (double)bark[nSpec - 1] / nb;
        }
    }
    bb[nb] = nSpec - 1;
}

mo_plan *mo_plan_create(int N, double sr, int window) {
    if (N < 2 || !mo_is_power_of_two((double)N)) return NULL; /* src/meyda.js:20-22 */
    mo_plan *p = (mo_plan *)calloc(1, sizeof(mo_plan));
    p->N = N; p->n = N / 2; p->sr = sr; p->window = window;
    p->hanning = (float *)malloc(sizeof(float) * N);
    p->hamming = (float *)malloc(sizeof(float) * N);
    p->blackman = (float *)malloc(sizeof(float) * N);
    p->bark = (float *)malloc(sizeof(float) * N);
    compute_hanning(p->hanning, N);
    compute_hamming(p->hamming, N);
    compute_blackman(p->blackman, N);
    compute_bark_scale(p->bark, N, sr);
    compute_bark_band_limits(p->bbLimits, p->bark, p->n, MO_NUM_BARK);
    return p;
}

void mo_plan_destroy(mo_plan *p) {
    if (!p) return;
    free(p->hanning); free(p->hamming); free(p->blackman); free(p->bark); free(p);
}

const float *mo_plan_window(const mo_plan *p, int which) {
    return which == MO_WIN_BLACKMAN ? p->blackman : which == MO_WIN_HAMMING ? p->hamming : p->hanning;
}
const float *mo_plan_bark(const mo_plan *p) { return p->bark; }
const int32_t *mo_plan_bb_limits(const mo_plan *p) { return p->bbLimits; }

/* lib/jsfft/fft.js:173-183 */
static int bit_reverse_index(int index, int n) {
    int r = 0;
    while (n > 1) { r <<= 1; r += index & 1; index >>= 1; n >>= 1; }
    return r;
}

/* lib/jsfft/fft.js:185-208 -- each (i, r_i) pair swapped once. */
static void bit_reverse_complex_array(float *re, float *im, int n) {
    for (int i = 0; i < n; i++) {
        int r = bit_reverse_index(i, n);
        if (r <= i) continue; /* same pairs as the flips{} hash-set, once each */
        float s = re[r]; re[r] = re[i]; re[i] = s;
        s = im[r]; im[r] = im[i]; im[i] = s;
    }
}

/* lib/jsfft/fft.js:123-171, inverse=false.  Doubles throughout; each
 * butterfly output is rounded to float by the Float32Array store. */
void mo_fft_jsfft(float *output_r, float *output_i, int n) {
    bit_reverse_complex_array(output_r, output_i, n);
    int width = 1;
    while (width < n) {
        double del_f_r = cos(M_PI / width);
        double del_f_i = sin(M_PI / width);
        for (int i = 0; i < n / (2 * width); i++) {
            double f_r = 1, f_i = 0;
            for (int j = 0; j < width; j++) {
                int l_index = 2 * i * width + j;
                int r_index = l_index + width;
                double left_r = output_r[l_index];
                double left_i = output_i[l_index];
                double right_r = f_r * output_r[r_index] - f_i * output_i[r_index];
                double right_i = f_i * output_r[r_index] + f_r * output_i[r_index];
                output_r[l_index] = (float)(M_SQRT1_2 * (left_r + right_r));
                output_i[l_index] = (float)(M_SQRT1_2 * (left_i + right_i));
                output_r[r_index] = (float)(M_SQRT1_2 * (left_r - right_r));
                output_i[r_index] = (float)(M_SQRT1_2 * (left_i - right_i));
                double temp = f_r * del_f_r - f_i * del_f_i;
                f_i = f_r * del_f_i + f_i * del_f_r;
                f_r = temp;
            }
        }
        width <<= 1;
    }
}

/* src/utils.js:1-11 -- Math.pow(k,i) is exact for these k,i. */
static double mu(int i, const float *amp, int len) {
    double numerator = 0, denominator = 0;
    for (int k = 0; k < len; k++) {
        double kp = 1;
        for (int e = 0; e < i; e++) kp *= (double)k;
        numerator += kp * fabs((double)amp[k]);
        denominator += amp[k];
    }
    return numerator / denominator;
}

/* src/extractors/loudness.js:47-96 */
static double loudness(const mo_plan *p, const float *amp, float *specific) {
    for (int i = 0; i < MO_NUM_BARK; i++) {
        double sum = 0;
        for (int j = p->bbLimits[i]; j < p->bbLimits[i + 1]; j++) sum += amp[j];
        specific[i] = (float)pow(sum, 0.23);
    }
    double total = 0;
    for (int i = 0; i < MO_NUM_BARK; i++) total += specific[i];
    return total;
}

/* src/extractors/mfcc.js:5-95 -- the filterbank and DCT matrix are rebuilt on
 * every call, as the reference does (rebuild=1), or cached per plan is NOT
 * offered here: literal means literal. */
static void mfcc(const mo_plan *p, const float *powSpec, float *mfccs, double *fb_scratch) {
    const int bufferSize = p->N, numFilters = MO_NUM_MEL;
    const double sampleRate = p->sr;
    float melValues[MO_NUM_MEL + 2], melValuesInFreq[MO_NUM_MEL + 2];
    double fftBinsOfFreq[MO_NUM_MEL + 2];
    double lowerLimitMel = 1125 * log(1 + (0.0 / 700));
    double upperLimitMel = 1125 * log(1 + ((sampleRate / 2) / 700));
    double range = upperLimitMel - lowerLimitMel;
    double valueToAdd = range / (numFilters + 1);
    for (int i = 0; i < numFilters + 2; i++) {
        melValues[i] = (float)(i * valueToAdd);
        melValuesInFreq[i] = (float)(700 * (exp((double)melValues[i] / 1125) - 1));
        fftBinsOfFreq[i] = floor((bufferSize + 1) * (double)melValuesInFreq[i] / sampleRate);
    }
    const int cols = bufferSize / 2 + 1;
    float loggedMelBands[MO_NUM_MEL];
    for (int j = 0; j < numFilters; j++) {
        double *row = fb_scratch;
        for (int i = 0; i < cols; i++) row[i] = 0;
        for (int i = (int)fftBinsOfFreq[j]; i < (int)fftBinsOfFreq[j + 1]; i++)
            row[i] = (i - fftBinsOfFreq[j]) / (fftBinsOfFreq[j + 1] - fftBinsOfFreq[j]);
        for (int i = (int)fftBinsOfFreq[j + 1]; i < (int)fftBinsOfFreq[j + 2]; i++)
            row[i] = (fftBinsOfFreq[j + 2] - i) / (fftBinsOfFreq[j + 2] - fftBinsOfFreq[j + 1]);
        /* mfcc.js:53-65: Float32Array accumulator => f32-rounded running sum */
        loggedMelBands[j] = 0;
        for (int q = 0; q < bufferSize / 2; q++) {
            row[q] = row[q] * (double)powSpec[q];
            loggedMelBands[j] = (float)((double)loggedMelBands[j] + row[q]);
        }
        loggedMelBands[j] = (float)log((double)loggedMelBands[j]);
    }
    /* mfcc.js:67-93 */
    double k = M_PI / numFilters;
    double w1 = 1.0 / sqrt((double)numFilters);
    double w2 = sqrt(2.0 / numFilters);
    const int numCoeffs = MO_NUM_MFCC;
    float dctMatrix[MO_NUM_MFCC * MO_NUM_MEL];
    for (int i = 0; i < numCoeffs; i++)
        for (int j = 0; j < numFilters; j++) {
            int idx = i + j * numCoeffs;
            dctMatrix[idx] = (float)((i == 0 ? w1 : w2) * cos(k * (i + 1) * (j + 0.5)));
        }
    for (int c = 0; c < numCoeffs; c++) {
        double v = 0;
        for (int f = 0; f < numFilters; f++) v += (double)dctMatrix[c + f * numCoeffs] * (double)loggedMelBands[f];
        mfccs[c] = (float)(v / numCoeffs);
    }
}

/* One buffer through the intended per-frame pipeline, src/meyda.js:69-91 with
 * the per-frame FFT of SURVEY 2.3-1, then every extractor. `work` holds
 * 2*N floats + (N/2+1) doubles of scratch. */
void mo_frame(const mo_plan *p, const float *signal, mo_frame_out *o, void *work) {
    const int N = p->N, n = p->n;
    const double sr = p->sr;
    float *re = (float *)work, *im = re + N;
    double *fb_scratch = (double *)(im + N);
    const float *win = mo_plan_window(p, p->window);

    /* computeWindow src/meyda.js:158-168; ComplexArray.map lib/jsfft/complex_array.js:54-70 */
    for (int i = 0; i < N; i++) { re[i] = (float)((double)signal[i] * (double)win[i]); im[i] = 0.0f; }
    mo_fft_jsfft(re, im, N);

    /* computeAmplitude src/meyda.js:104-114 */
    float *amp = o->amp ? o->amp : (float *)malloc(sizeof(float) * n);
    for (int i = 0; i < n; i++)
        amp[i] = (float)sqrt((double)re[i] * (double)re[i] + (double)im[i] * (double)im[i]);

    if (o->buffer) memcpy(o->buffer, signal, sizeof(float) * N);      /* docs.md:19-21 */
    if (o->cs_real) memcpy(o->cs_real, re, sizeof(float) * N);        /* complexSpectrum.js:1-3 */
    if (o->cs_imag) memcpy(o->cs_imag, im, sizeof(float) * N);

    /* powerSpectrum.js:1-7 */
    float *power = o->power ? o->power : (float *)malloc(sizeof(float) * n);
    for (int i = 0; i < n; i++) power[i] = (float)((double)amp[i] * (double)amp[i]);

    /* rms.js:1-11, energy.js:1-7 */
    double acc = 0;
    for (int i = 0; i < N; i++) acc += (double)signal[i] * (double)signal[i];
    o->rms = sqrt(acc / N);
    acc = 0;
    for (int i = 0; i < N; i++) { double a = fabs((double)signal[i]); acc += a * a; }
    o->energy = acc;

    /* zcr.js:1-9 -- signal[N] is undefined: both comparisons false. */
    int z = 0;
    for (int i = 0; i + 1 < N; i++)
        if ((signal[i] >= 0 && signal[i + 1] < 0) || (signal[i] < 0 && signal[i + 1] >= 0)) z++;
    o->zcr = z;

    /* spectralCentroid.js, spectralSpread.js, spectralSkewness.js, spectralKurtosis.js */
    double m1 = mu(1, amp, n), m2 = mu(2, amp, n), m3 = mu(3, amp, n), m4 = mu(4, amp, n);
    o->centroid = m1;
    o->spread = sqrt(m2 - pow(m1, 2));
    o->skewness = (2 * pow(m1, 3) - 3 * m1 * m2 + m3) / pow(sqrt(m2 - pow(m1, 2)), 3);
    o->kurtosis = (-3 * pow(m1, 4) + 6 * m1 * m2 - 4 * m1 * m3 + m4) / pow(sqrt(m2 - pow(m1, 2)), 4);

    /* spectralFlatness.js:1-10 */
    {
        double numerator = 0, denominator = 0;
        for (int i = 0; i < n; i++) { numerator += log((double)amp[i]); denominator += amp[i]; }
        o->flatness = exp(numerator / n) * n / denominator;
    }
    /* spectralSlope.js:1-18 */
    {
        double ampSum = 0, freqSum = 0, powFreqSum = 0, ampFreqSum = 0;
        for (int i = 0; i < n; i++) {
            ampSum += amp[i];
            double curFreq = i * sr / N;
            powFreqSum += curFreq * curFreq;
            freqSum += curFreq;
            ampFreqSum += curFreq * amp[i];
        }
        o->slope = (n * ampFreqSum - freqSum * ampSum) / (ampSum * (powFreqSum - pow(freqSum, 2)));
    }
    /* spectralRolloff.js:1-16 */
    {
        double nyqBin = sr / (2 * (n - 1));
        double ec = 0;
        for (int i = 0; i < n; i++) ec += amp[i];
        double threshold = 0.99 * ec;
        int q = n - 1;
        while (ec > threshold && q >= 0) { ec -= amp[q]; --q; }
        o->rolloff = (q + 1) * nyqBin;
    }
    /* loudness.js, perceptualSpread.js:1-14, perceptualSharpness.js:1-16 */
    {
        float *spec = o->loudness_specific;
        double total = loudness(p, amp, spec);
        o->loudness_total = total;
        double max = 0;
        for (int i = 0; i < MO_NUM_BARK; i++) if (spec[i] > max) max = spec[i];
        o->perceptual_spread = pow((total - max) / total, 2);
        double output = 0;
        for (int i = 0; i < MO_NUM_BARK; i++) {
            if (i < 15) output += (i + 1) * (double)spec[i + 1];
            else output += 0.066 * exp(0.171 * (i + 1));
        }
        output *= 0.11 / total;
        o->perceptual_sharpness = output;
    }
    mfcc(p, power, o->mfcc, fb_scratch);
    if (!o->amp) free(amp);
    if (!o->power) free(power);
}

size_t mo_work_bytes(const mo_plan *p) {
    return sizeof(float) * 2 * (size_t)p->N + sizeof(double) * ((size_t)p->N / 2 + 1);
}

/* SoA batch outputs (any pointer may be NULL); frame-major. scalars: 13
 * doubles per frame in the order of mo_scalar_names(). */
typedef struct mo_batch_out {
    double *scalars;    /* [frames][13] */
    float *specific;    /* [frames][24] */
    float *mfcc;        /* [frames][13] */
    float *buffer;      /* [frames][N]  */
    float *cs_real;     /* [frames][N]  */
    float *cs_imag;     /* [frames][N]  */
    float *amp;         /* [frames][N/2] */
    float *power;       /* [frames][N/2] */
} mo_batch_out;

#define MO_NUM_SCALARS 13

int64_t mo_num_frames(int64_t len, int N, int hop) {
    return len < N ? 0 : (len - N) / hop + 1;
}

/* ring > 0: results go to slot ring_base + (frame index % ring) instead of the
 * frame's own row, so a long timing run needs only a bounded output arena. */
static void batch_range(const mo_plan *p, const float *samples, int hop, int64_t f0, int64_t f1,
                        const mo_batch_out *out, int64_t out_base, int64_t ring, int64_t ring_base) {
    void *work = malloc(mo_work_bytes(p));
    const int N = p->N, n = p->n;
    for (int64_t f = f0; f < f1; f++) {
        mo_frame_out o;
        memset(&o, 0, sizeof(o));
        int64_t g = ring > 0 ? ring_base + ((out_base + f) % ring) : out_base + f;
        o.buffer = out->buffer ? out->buffer + g * N : NULL;
        o.cs_real = out->cs_real ? out->cs_real + g * N : NULL;
        o.cs_imag = out->cs_imag ? out->cs_imag + g * N : NULL;
        o.amp = out->amp ? out->amp + g * n : NULL;
        o.power = out->power ? out->power + g * n : NULL;
        mo_frame(p, samples + f * hop, &o, work);
        if (out->scalars) {
            double *s = out->scalars + g * MO_NUM_SCALARS;
            s[0] = o.rms; s[1] = o.energy; s[2] = o.zcr; s[3] = o.centroid; s[4] = o.flatness;
            s[5] = o.slope; s[6] = o.rolloff; s[7] = o.spread; s[8] = o.skewness; s[9] = o.kurtosis;
            s[10] = o.loudness_total; s[11] = o.perceptual_spread; s[12] = o.perceptual_sharpness;
        }
        if (out->specific) memcpy(out->specific + g * MO_NUM_BARK, o.loudness_specific, sizeof(float) * MO_NUM_BARK);
        if (out->mfcc) memcpy(out->mfcc + g * MO_NUM_MFCC, o.mfcc, sizeof(float) * MO_NUM_MFCC);
    }
    free(work);
}

/* One clip, frames f of [f*hop, f*hop+N); returns the number of frames. */
int64_t mo_extract_clip(const mo_plan *p, const float *samples, int64_t len, int hop,
                        const mo_batch_out *out, int64_t out_base) {
    int64_t nf = mo_num_frames(len, p->N, hop);
    batch_range(p, samples, hop, 0, nf, out, out_base, 0, 0);
    return nf;
}

/* Multi-threaded driver for the CPU baseline: clips of equal length laid out
 * back to back; frames are split evenly over `threads` pthreads. */
typedef struct {
    const mo_plan *p; const float *samples; int64_t clip_len; int hop; int64_t fpc;
    int64_t g0, g1; const mo_batch_out *out; int64_t ring, ring_base;
} mo_job;

static void *job_main(void *arg) {
    mo_job *j = (mo_job *)arg;
    int64_t g = j->g0;
    while (g < j->g1) {
        int64_t clip = g / j->fpc, f0 = g % j->fpc;
        int64_t f1 = f0 + (j->g1 - g);
        if (f1 > j->fpc) f1 = j->fpc;
        batch_range(j->p, j->samples + clip * j->clip_len, j->hop, f0, f1, j->out, clip * j->fpc, j->ring,
                    j->ring_base);
        g += f1 - f0;
    }
    return NULL;
}

/* ring_per_thread > 0: the output arrays hold threads * ring_per_thread frames. */
int64_t mo_extract_threads(const mo_plan *p, const float *samples, int64_t n_clips, int64_t clip_len,
                           int hop, const mo_batch_out *out, int threads, int64_t ring_per_thread) {
    int64_t fpc = mo_num_frames(clip_len, p->N, hop);
    int64_t total = fpc * n_clips;
    if (total == 0) return 0;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256];
    mo_job jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t].p = p; jobs[t].samples = samples; jobs[t].clip_len = clip_len; jobs[t].hop = hop;
        jobs[t].fpc = fpc; jobs[t].out = out;
        jobs[t].g0 = total * t / threads; jobs[t].g1 = total * (t + 1) / threads;
        jobs[t].ring = ring_per_thread; jobs[t].ring_base = ring_per_thread * t;
        pthread_create(&tid[t], NULL, job_main, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(tid[t], NULL);
    return total;
}
