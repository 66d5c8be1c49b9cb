// addon.cc -- Node.js N-API veneer over the C ABI of include/meyda_b200.h.
//
// NOT BUILT OR RUN IN THIS REPO'S IMAGE: there is no `node` and no node_api.h
// here.  It is the reference-side binding a maintainer
// adds on a machine with Node >= 18 + CUDA; everything below the C ABI is
// exercised from Python (meyda_b200/_capi.py binds the same entry points).
//
//   const native = require('./build/Release/meyda_b200.node')
//   const plan = native.createPlan({bufferSize, hop, sampleRate, window, featureMask, device, flags})
//   const out  = native.extract(plan, samples /*Float32Array*/, offsets /*BigInt64Array*/, lengths /*BigInt64Array*/)
//   const out  = native.extractPcm16(plan, pcm /*Int16Array, interleaved*/, channels, channel, offsets, lengths)
//   const out  = await native.extractAsync(plan, samples, offsets, lengths)     // napi_async_work on the libuv pool
//   const out  = await native.extractPcm16Async(plan, pcm, channels, channel, offsets, lengths)
//   const info = native.wavInfo(fileBytes /*Uint8Array*/)   // {format, channels, sampleRate, bitsPerSample, dataOffset, sampleFrames}
//   const out  = native.extractMulti([plan0, plan1, ...], samples, offsets, lengths)   // one plan per device, clip-sharded
//   const prm  = native.getParams(plan)            // {numBarkBands, numMelFilters, numMfccCoefficients, rolloffFraction}
//   const n    = native.refinedFrames(plan)        // frames of the last call redone with the exact FFT (adaptive plans)
//   native.setHostThreads(n)                        // host threads for the rows the device does not produce (0: automatic)
//   const st   = native.createStream(plan)         // the onaudioprocess cadence (src/meyda.js:69-91)
//   const out  = native.streamPush(st, block /*Float32Array*/)   // features of the frames this block completes
//   native.streamReset(st); native.destroyStream(st)
//   native.destroyPlan(plan)
//
// Output arrays live in page-locked memory (mb_host_alloc, wrapped as external ArrayBuffers and released by
// mb_host_free when collected), so the device copies straight into them, asynchronously.
//
// `out` maps mb_outputs field names to typed arrays laid out exactly as the C
// ABI documents (frame-major SoA); js/meyda_b200.js turns them into the
// reference's per-frame `get([...])` objects.
#include <node_api.h>
#include <stdint.h>
#include <string.h>

#include "../include/meyda_b200.h"

#define NAPI_OK_OR_THROW(env, call)                                   \
    do {                                                              \
        if ((call) != napi_ok) {                                      \
            napi_throw_error((env), NULL, "N-API call failed: " #call); \
            return NULL;                                              \
        }                                                             \
    } while (0)

static napi_value throw_mb(napi_env env, mb_status st) {
    // src/meyda.js:20-22 throws `new Error(...)` with this exact text for MB_ERR_NOT_POWER_OF_TWO
    napi_throw_error(env, st == MB_ERR_NOT_POWER_OF_TWO ? "ERR_MEYDA_BUFFER_SIZE" : "ERR_MEYDA_NATIVE", mb_last_error());
    return NULL;
}

static int32_t get_i32(napi_env env, napi_value obj, const char *key, int32_t dflt) {
    napi_value v;
    bool has = false;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
    int32_t out = dflt;
    if (napi_get_named_property(env, obj, key, &v) == napi_ok) napi_get_value_int32(env, v, &out);
    return out;
}
static double get_f64(napi_env env, napi_value obj, const char *key, double dflt) {
    napi_value v;
    bool has = false;
    if (napi_has_named_property(env, obj, key, &has) != napi_ok || !has) return dflt;
    double out = dflt;
    if (napi_get_named_property(env, obj, key, &v) == napi_ok) napi_get_value_double(env, v, &out);
    return out;
}

static void plan_finalize(napi_env, void *data, void *) { mb_plan_destroy((mb_plan *)data); }

static napi_value CreatePlan(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    const int32_t n = get_i32(env, argv[0], "bufferSize", 0);
    mb_plan *plan = NULL;
    // 0 = the reference's constant (loudness.js:14, mfcc.js:15,71, spectralRolloff.js:9)
    mb_params prm = {get_i32(env, argv[0], "numBarkBands", 0), get_i32(env, argv[0], "numMelFilters", 0),
                     get_i32(env, argv[0], "numMfccCoefficients", 0), 0, get_f64(env, argv[0], "rolloffFraction", 0.0)};
    mb_status st = mb_plan_create_ex(&plan, get_i32(env, argv[0], "device", 0), n, get_i32(env, argv[0], "hop", n),
                                     get_f64(env, argv[0], "sampleRate", 44100.0), get_i32(env, argv[0], "window", 0),
                                     (uint32_t)get_i32(env, argv[0], "featureMask", (int32_t)MB_ALL_FEATURES),
                                     (uint32_t)get_i32(env, argv[0], "flags", 0), &prm);
    if (st != MB_OK) return throw_mb(env, st);
    napi_value ext;
    NAPI_OK_OR_THROW(env, napi_create_external(env, plan, plan_finalize, NULL, &ext));
    return ext;
}

struct FieldDesc { const char *name; int feature; int kind; size_t offset; };  // kind 0:1 1:N 2:N/2 3:Bark bands (24) 4:mfcc coefficients (13)
#define F(name, feat, kind) {#name, feat, kind, offsetof(mb_outputs, name)}
static const FieldDesc kFields[] = {
    F(buffer, MB_FEAT_BUFFER, 1), F(rms, MB_FEAT_RMS, 0), F(energy, MB_FEAT_ENERGY, 0), F(zcr, MB_FEAT_ZCR, 0),
    F(complex_real, MB_FEAT_COMPLEX_SPECTRUM, 1), F(complex_imag, MB_FEAT_COMPLEX_SPECTRUM, 1),
    F(amplitude_spectrum, MB_FEAT_AMPLITUDE_SPECTRUM, 2), F(power_spectrum, MB_FEAT_POWER_SPECTRUM, 2),
    F(spectral_centroid, MB_FEAT_SPECTRAL_CENTROID, 0), F(spectral_flatness, MB_FEAT_SPECTRAL_FLATNESS, 0),
    F(spectral_slope, MB_FEAT_SPECTRAL_SLOPE, 0), F(spectral_rolloff, MB_FEAT_SPECTRAL_ROLLOFF, 0),
    F(spectral_spread, MB_FEAT_SPECTRAL_SPREAD, 0), F(spectral_skewness, MB_FEAT_SPECTRAL_SKEWNESS, 0),
    F(spectral_kurtosis, MB_FEAT_SPECTRAL_KURTOSIS, 0), F(loudness_specific, MB_FEAT_LOUDNESS, 3),
    F(loudness_total, MB_FEAT_LOUDNESS, 0), F(perceptual_spread, MB_FEAT_PERCEPTUAL_SPREAD, 0),
    F(perceptual_sharpness, MB_FEAT_PERCEPTUAL_SHARPNESS, 0), F(mfcc, MB_FEAT_MFCC, 4)};

static void host_finalize(napi_env, void *data, void *) { mb_host_free(data); }

// A typed array over page-locked memory; plain (pageable) ArrayBuffer if the allocation fails.
static napi_status make_output_array(napi_env env, bool is_int, size_t elems, void **data, napi_value *ta) {
    napi_value ab;
    void *pinned = NULL;
    napi_status ns = napi_generic_failure;
    if (elems > 0 && mb_host_alloc(&pinned, elems * 4) == MB_OK) {
        ns = napi_create_external_arraybuffer(env, pinned, elems * 4, host_finalize, NULL, &ab);
        if (ns != napi_ok) mb_host_free(pinned);  // (engines that forbid external buffers: fall back below)
        else *data = pinned;
    }
    if (ns != napi_ok) {
        ns = napi_create_arraybuffer(env, elems * 4, data, &ab);
        if (ns != napi_ok) return ns;
    }
    return napi_create_typedarray(env, is_int ? napi_int32_array : napi_float32_array, elems, ab, 0, ta);
}

// The result object for `lay`: one typed array per requested field (pointers also written into *out) plus
// totalFrames / numBarkBands / numMfccCoefficients.
static napi_value make_result(napi_env env, const mb_layout &lay, int64_t frames, mb_outputs *out);

// One extract call: arguments parsed and outputs allocated on the JS thread, the blocking C-ABI call either made in
// place (extract / extractPcm16) or on the libuv pool (extractAsync / extractPcm16Async, SURVEY.md 8b "Threading").
struct Job {
    mb_plan *plan = NULL;
    bool pcm16 = false;
    void *samples = NULL, *offs = NULL, *lens = NULL;
    size_t n_samples = 0, n_clips = 0;
    int32_t channels = 1, channel = 0;
    mb_outputs out;
    // async only
    mb_status status = MB_OK;
    char error[512];
    napi_async_work work = NULL;
    napi_deferred deferred = NULL;
    napi_ref keep_result = NULL, keep_args = NULL;  // the typed arrays must outlive the pool thread's use of them
};

static mb_status run_job(Job *j) {
    if (j->pcm16)  // the WAV data chunk as it is: conversion and channel pick happen on the device
        return mb_extract_pcm16(j->plan, (const int16_t *)j->samples, (int64_t)(j->n_samples / (j->channels > 0 ? j->channels : 1)),
                                j->channels, j->channel, (const int64_t *)j->offs, (const int64_t *)j->lens, (int64_t)j->n_clips,
                                &j->out, MB_MEM_HOST);
    return mb_extract(j->plan, (const float *)j->samples, (int64_t)j->n_samples, (const int64_t *)j->offs, (const int64_t *)j->lens,
                      (int64_t)j->n_clips, &j->out, MB_MEM_HOST);
}

// (plan, Float32Array, offsets, lengths)  /  (plan, Int16Array, channels, channel, offsets, lengths).
// Returns the result object (typed arrays allocated, not yet filled) or NULL with an exception pending.
static napi_value prepare_job(napi_env env, napi_callback_info info, bool pcm16, Job *j, napi_value *args_array) {
    size_t argc = 6;
    napi_value argv[6];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    j->pcm16 = pcm16;
    memset(&j->out, 0, sizeof(j->out));
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&j->plan));
    napi_typedarray_type tt;
    size_t n_len = 0;
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[1], &tt, &j->n_samples, &j->samples, NULL, NULL));
    if (tt != (pcm16 ? napi_int16_array : napi_float32_array)) {
        napi_throw_type_error(env, NULL, pcm16 ? "pcm must be an Int16Array" : "samples must be a Float32Array");
        return NULL;
    }
    if (pcm16) {
        NAPI_OK_OR_THROW(env, napi_get_value_int32(env, argv[2], &j->channels));
        NAPI_OK_OR_THROW(env, napi_get_value_int32(env, argv[3], &j->channel));
    }
    const int a0 = pcm16 ? 4 : 2;
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[a0], &tt, &j->n_clips, &j->offs, NULL, NULL));
    if (tt != napi_bigint64_array) { napi_throw_type_error(env, NULL, "offsets must be a BigInt64Array"); return NULL; }
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[a0 + 1], &tt, &n_len, &j->lens, NULL, NULL));
    if (tt != napi_bigint64_array || n_len != j->n_clips) { napi_throw_type_error(env, NULL, "lengths must match offsets"); return NULL; }

    mb_layout lay;
    mb_status st = mb_query_output(j->plan, (int64_t)j->n_clips, (const int64_t *)j->lens, NULL, &lay);
    if (st != MB_OK) return throw_mb(env, st);
    napi_value result = make_result(env, lay, lay.total_frames, &j->out);
    if (!result) return NULL;
    if (args_array) {  // plan + every input array, kept alive while the pool thread reads them
        NAPI_OK_OR_THROW(env, napi_create_array_with_length(env, argc, args_array));
        for (size_t i = 0; i < argc; i++) NAPI_OK_OR_THROW(env, napi_set_element(env, *args_array, (uint32_t)i, argv[i]));
    }
    return result;
}

static napi_value make_result(napi_env env, const mb_layout &lay, int64_t frames_n, mb_outputs *out) {
    napi_value result;
    NAPI_OK_OR_THROW(env, napi_create_object(env, &result));
    memset(out, 0, sizeof(*out));
    for (const FieldDesc &f : kFields) {
        if (!((lay.feature_mask >> f.feature) & 1u)) continue;
        const size_t per = f.kind == 0 ? 1 : f.kind == 1 ? (size_t)lay.buffer_size : f.kind == 2 ? (size_t)lay.spectrum_size
                                         : f.kind == 3 ? (size_t)lay.num_bark_bands : (size_t)lay.num_mfcc;
        napi_value ta;
        void *data = NULL;
        NAPI_OK_OR_THROW(env, make_output_array(env, strcmp(f.name, "zcr") == 0, per * (size_t)frames_n, &data, &ta));
        *(void **)((char *)out + f.offset) = data;
        NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, f.name, ta));
    }
    napi_value frames, nbands, ncoefs;
    NAPI_OK_OR_THROW(env, napi_create_int64(env, frames_n, &frames));
    NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, "totalFrames", frames));
    NAPI_OK_OR_THROW(env, napi_create_int32(env, lay.num_bark_bands, &nbands));  // row widths of loudness_specific and mfcc
    NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, "numBarkBands", nbands));
    NAPI_OK_OR_THROW(env, napi_create_int32(env, lay.num_mfcc, &ncoefs));
    NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, "numMfccCoefficients", ncoefs));
    return result;
}

static napi_value ExtractImpl(napi_env env, napi_callback_info info, bool pcm16) {
    Job j;
    napi_value result = prepare_job(env, info, pcm16, &j, NULL);
    if (!result) return NULL;
    const mb_status st = run_job(&j);
    if (st != MB_OK) return throw_mb(env, st);
    return result;
}

static napi_value Extract(napi_env env, napi_callback_info info) { return ExtractImpl(env, info, false); }
static napi_value ExtractPcm16(napi_env env, napi_callback_info info) { return ExtractImpl(env, info, true); }

// ---- the same calls as Promises: mb_extract* block, so they run on the libuv pool (one plan must not be used by two
// jobs at once: the C ABI is thread-compatible, not thread-safe).  No N-API call is made off the JS thread.
static void job_execute(napi_env, void *data) {
    Job *j = (Job *)data;
    j->status = run_job(j);
    if (j->status != MB_OK) {  // mb_last_error() is thread-local: take it on this thread
        strncpy(j->error, mb_last_error(), sizeof(j->error) - 1);
        j->error[sizeof(j->error) - 1] = 0;
    }
}

static void job_complete(napi_env env, napi_status status, void *data) {
    Job *j = (Job *)data;
    napi_value result = NULL;
    if (status == napi_ok && j->status == MB_OK && napi_get_reference_value(env, j->keep_result, &result) == napi_ok) {
        napi_resolve_deferred(env, j->deferred, result);
    } else {
        napi_value code, msg, err;
        const char *text = status != napi_ok ? "extract was cancelled" : j->error;
        napi_create_string_utf8(env, j->status == MB_ERR_NOT_POWER_OF_TWO ? "ERR_MEYDA_BUFFER_SIZE" : "ERR_MEYDA_NATIVE",
                                NAPI_AUTO_LENGTH, &code);
        napi_create_string_utf8(env, text, NAPI_AUTO_LENGTH, &msg);
        napi_create_error(env, code, msg, &err);
        napi_reject_deferred(env, j->deferred, err);
    }
    napi_delete_reference(env, j->keep_result);
    napi_delete_reference(env, j->keep_args);
    napi_delete_async_work(env, j->work);
    delete j;
}

static napi_value ExtractAsyncImpl(napi_env env, napi_callback_info info, bool pcm16) {
    Job *j = new Job();
    napi_value args, promise, name;
    napi_value result = prepare_job(env, info, pcm16, j, &args);
    if (!result) { delete j; return NULL; }
    if (napi_create_reference(env, result, 1, &j->keep_result) != napi_ok ||
        napi_create_reference(env, args, 1, &j->keep_args) != napi_ok ||
        napi_create_promise(env, &j->deferred, &promise) != napi_ok ||
        napi_create_string_utf8(env, "meyda_b200.extract", NAPI_AUTO_LENGTH, &name) != napi_ok ||
        napi_create_async_work(env, NULL, name, job_execute, job_complete, j, &j->work) != napi_ok ||
        napi_queue_async_work(env, j->work) != napi_ok) {
        if (j->keep_result) napi_delete_reference(env, j->keep_result);
        if (j->keep_args) napi_delete_reference(env, j->keep_args);
        if (j->work) napi_delete_async_work(env, j->work);
        delete j;
        napi_throw_error(env, NULL, "could not queue the extract job");
        return NULL;
    }
    return promise;
}

static napi_value ExtractAsync(napi_env env, napi_callback_info info) { return ExtractAsyncImpl(env, info, false); }
static napi_value ExtractPcm16Async(napi_env env, napi_callback_info info) { return ExtractAsyncImpl(env, info, true); }

// wavInfo(Uint8Array) -> the fields of mb_wav_info (replaces the header half of decodeAudioData, lib/bufferLoader.js:28-38)
static napi_value WavInfo(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    napi_typedarray_type tt;
    size_t n = 0;
    void *bytes = NULL;
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[0], &tt, &n, &bytes, NULL, NULL));
    if (tt != napi_uint8_array) { napi_throw_type_error(env, NULL, "file bytes must be a Uint8Array"); return NULL; }
    mb_wav_info wi;
    const mb_status st = mb_wav_parse(bytes, (int64_t)n, &wi);
    if (st != MB_OK) return throw_mb(env, st);
    napi_value result, v;
    NAPI_OK_OR_THROW(env, napi_create_object(env, &result));
    const struct { const char *name; int64_t value; } fields[] = {
        {"format", wi.format}, {"channels", wi.channels}, {"sampleRate", wi.sample_rate},
        {"bitsPerSample", wi.bits_per_sample}, {"dataOffset", wi.data_offset}, {"sampleFrames", wi.n_sample_frames}};
    for (const auto &f : fields) {
        NAPI_OK_OR_THROW(env, napi_create_int64(env, f.value, &v));
        NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, f.name, v));
    }
    return result;
}

// extractMulti([plan, ...], Float32Array, offsets, lengths): mb_extract_multi -- contiguous clip ranges balanced on frame
// count, one plan per device, results exactly where a single-device call would put them.
static napi_value ExtractMulti(napi_env env, napi_callback_info info) {
    size_t argc = 4;
    napi_value argv[4];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    uint32_t n_plans = 0;
    NAPI_OK_OR_THROW(env, napi_get_array_length(env, argv[0], &n_plans));
    if (n_plans == 0 || n_plans > 64) { napi_throw_type_error(env, NULL, "plans must be an array of 1..64 plans"); return NULL; }
    mb_plan *plans[64];
    for (uint32_t i = 0; i < n_plans; i++) {
        napi_value e;
        NAPI_OK_OR_THROW(env, napi_get_element(env, argv[0], i, &e));
        NAPI_OK_OR_THROW(env, napi_get_value_external(env, e, (void **)&plans[i]));
    }
    napi_typedarray_type tt;
    size_t n_samples = 0, n_clips = 0, n_len = 0;
    void *samples = NULL, *offs = NULL, *lens = NULL;
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[1], &tt, &n_samples, &samples, NULL, NULL));
    if (tt != napi_float32_array) { napi_throw_type_error(env, NULL, "samples must be a Float32Array"); return NULL; }
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[2], &tt, &n_clips, &offs, NULL, NULL));
    if (tt != napi_bigint64_array) { napi_throw_type_error(env, NULL, "offsets must be a BigInt64Array"); return NULL; }
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[3], &tt, &n_len, &lens, NULL, NULL));
    if (tt != napi_bigint64_array || n_len != n_clips) { napi_throw_type_error(env, NULL, "lengths must match offsets"); return NULL; }
    mb_layout lay;
    mb_status st = mb_query_output(plans[0], (int64_t)n_clips, (const int64_t *)lens, NULL, &lay);
    if (st != MB_OK) return throw_mb(env, st);
    mb_outputs out;
    napi_value result = make_result(env, lay, lay.total_frames, &out);
    if (!result) return NULL;
    st = mb_extract_multi(plans, (int)n_plans, (const float *)samples, (int64_t)n_samples, (const int64_t *)offs,
                          (const int64_t *)lens, (int64_t)n_clips, &out);
    if (st != MB_OK) return throw_mb(env, st);
    return result;
}

// getParams(plan): the constants the plan runs with (mb_plan_get_params)
static napi_value GetParams(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1], result, v;
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    mb_plan *plan = NULL;
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&plan));
    mb_params prm;
    const mb_status st = mb_plan_get_params(plan, &prm);
    if (st != MB_OK) return throw_mb(env, st);
    NAPI_OK_OR_THROW(env, napi_create_object(env, &result));
    const struct { const char *name; int32_t value; } ints[] = {
        {"numBarkBands", prm.num_bark_bands}, {"numMelFilters", prm.num_mel_filters}, {"numMfccCoefficients", prm.num_mfcc}};
    for (const auto &f : ints) {
        NAPI_OK_OR_THROW(env, napi_create_int32(env, f.value, &v));
        NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, f.name, v));
    }
    NAPI_OK_OR_THROW(env, napi_create_double(env, prm.rolloff_fraction, &v));
    NAPI_OK_OR_THROW(env, napi_set_named_property(env, result, "rolloffFraction", v));
    return result;
}

// refinedFrames(plan): how many frames of the last call the adaptive plan redid with the reference's own FFT arithmetic
static napi_value RefinedFrames(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1], v;
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    mb_plan *plan = NULL;
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&plan));
    int64_t n = 0;
    const mb_status st = mb_plan_refined_frames(plan, &n);
    if (st != MB_OK) return throw_mb(env, st);
    NAPI_OK_OR_THROW(env, napi_create_int64(env, n, &v));
    return v;
}

// setHostThreads(n): host threads a host-memory call uses for the `buffer` and powerSpectrum rows (mb_set_host_threads)
static napi_value SetHostThreads(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    int32_t n = 0;
    NAPI_OK_OR_THROW(env, napi_get_value_int32(env, argv[0], &n));
    const mb_status st = mb_set_host_threads(n);
    if (st != MB_OK) return throw_mb(env, st);
    return NULL;
}

// setHostRows(mode): which rows of a host-memory call the host threads produce (2 / 1 / 0, -1: automatic; mb_set_host_rows)
static napi_value SetHostRows(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    int32_t mode = -1;
    NAPI_OK_OR_THROW(env, napi_get_value_int32(env, argv[0], &mode));
    const mb_status st = mb_set_host_rows(mode);
    if (st != MB_OK) return throw_mb(env, st);
    return NULL;
}

// getHostRows() -> 0 | 1 | 2: the mode in force (the automatic choice resolved)
static napi_value GetHostRows(napi_env env, napi_callback_info) {
    napi_value v;
    NAPI_OK_OR_THROW(env, napi_create_int32(env, mb_get_host_rows(), &v));
    return v;
}

// ---- streaming, the reference's actual usage model: one buffer per onaudioprocess event (src/meyda.js:69-91)
static void stream_finalize(napi_env, void *data, void *) { mb_stream_destroy((mb_stream *)data); }

// createStream(plan) -> stream.  The plan must outlive the stream (the facade keeps both).
static napi_value CreateStream(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1], ext;
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    mb_plan *plan = NULL;
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&plan));
    mb_stream *st = NULL;
    const mb_status s = mb_stream_create(&st, plan);
    if (s != MB_OK) return throw_mb(env, s);
    NAPI_OK_OR_THROW(env, napi_create_external(env, st, stream_finalize, NULL, &ext));
    return ext;
}

// streamPush(stream, plan, Float32Array block) -> the result object for the frames this block completes (0 or more)
static napi_value StreamPush(napi_env env, napi_callback_info info) {
    size_t argc = 3;
    napi_value argv[3];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    mb_stream *st = NULL;
    mb_plan *plan = NULL;
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&st));
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[1], (void **)&plan));
    napi_typedarray_type tt;
    size_t n_new = 0;
    void *block = NULL;
    NAPI_OK_OR_THROW(env, napi_get_typedarray_info(env, argv[2], &tt, &n_new, &block, NULL, NULL));
    if (tt != napi_float32_array) { napi_throw_type_error(env, NULL, "block must be a Float32Array"); return NULL; }
    const int64_t nf = mb_stream_frames_after(st, (int64_t)n_new);
    mb_layout lay;
    const int64_t one = 0;
    mb_status s = mb_query_output(plan, 0, &one, NULL, &lay);  // (shapes only: feature mask, row widths)
    if (s != MB_OK) return throw_mb(env, s);
    mb_outputs out;
    napi_value result = make_result(env, lay, nf, &out);
    if (!result) return NULL;
    int64_t done = 0;
    s = mb_stream_push(st, (const float *)block, (int64_t)n_new, &out, MB_MEM_HOST, &done);
    if (s != MB_OK) return throw_mb(env, s);
    return result;
}

static napi_value StreamReset(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    mb_stream *st = NULL;
    NAPI_OK_OR_THROW(env, napi_get_value_external(env, argv[0], (void **)&st));
    const mb_status s = mb_stream_reset(st);
    if (s != MB_OK) return throw_mb(env, s);
    return NULL;
}

static napi_value DestroyStream(napi_env env, napi_callback_info info) {
    (void)env; (void)info;  // externals are released by their finalizers (stream_finalize)
    return NULL;
}

static napi_value DestroyPlan(napi_env env, napi_callback_info info) {
    size_t argc = 1;
    napi_value argv[1];
    NAPI_OK_OR_THROW(env, napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
    void *plan = NULL;
    if (napi_remove_wrap(env, argv[0], &plan) != napi_ok) { /* externals are released by the finalizer */ }
    return NULL;
}

static napi_value Init(napi_env env, napi_value exports) {
    napi_property_descriptor props[] = {
        {"createPlan", NULL, CreatePlan, NULL, NULL, NULL, napi_default, NULL},
        {"extract", NULL, Extract, NULL, NULL, NULL, napi_default, NULL},
        {"extractPcm16", NULL, ExtractPcm16, NULL, NULL, NULL, napi_default, NULL},
        {"extractAsync", NULL, ExtractAsync, NULL, NULL, NULL, napi_default, NULL},
        {"extractPcm16Async", NULL, ExtractPcm16Async, NULL, NULL, NULL, napi_default, NULL},
        {"wavInfo", NULL, WavInfo, NULL, NULL, NULL, napi_default, NULL},
        {"destroyPlan", NULL, DestroyPlan, NULL, NULL, NULL, napi_default, NULL},
        {"extractMulti", NULL, ExtractMulti, NULL, NULL, NULL, napi_default, NULL},
        {"getParams", NULL, GetParams, NULL, NULL, NULL, napi_default, NULL},
        {"refinedFrames", NULL, RefinedFrames, NULL, NULL, NULL, napi_default, NULL},
        {"setHostThreads", NULL, SetHostThreads, NULL, NULL, NULL, napi_default, NULL},
        {"setHostRows", NULL, SetHostRows, NULL, NULL, NULL, napi_default, NULL},
        {"getHostRows", NULL, GetHostRows, NULL, NULL, NULL, napi_default, NULL},
        {"createStream", NULL, CreateStream, NULL, NULL, NULL, napi_default, NULL},
        {"streamPush", NULL, StreamPush, NULL, NULL, NULL, napi_default, NULL},
        {"streamReset", NULL, StreamReset, NULL, NULL, NULL, napi_default, NULL},
        {"destroyStream", NULL, DestroyStream, NULL, NULL, NULL, napi_default, NULL},
    };
    napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
    return exports;
}
NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
