"""ctypes binding of include/meyda_b200.h (the C ABI a Node N-API addon would
bind the same way).  Loading fails loudly when the CUDA
library has not been built: there is no CPU fallback behind this module."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MEYDA_B200_LIB points at an alternative build of the same library (kernel tuning experiments).
LIB_PATH = os.environ.get("MEYDA_B200_LIB") or os.path.join(_HERE, "_lib", "libmeyda_b200.so")

MB_OK = 0
MB_ERR_INVALID_ARG, MB_ERR_NOT_POWER_OF_TWO, MB_ERR_UNSUPPORTED, MB_ERR_CUDA = 1, 2, 3, 4
MB_ERR_NO_DEVICE, MB_ERR_MISSING_OUTPUT, MB_ERR_OUT_OF_RANGE = 5, 6, 7
MB_MEM_HOST, MB_MEM_DEVICE = 0, 1
MB_WINDOW = {"hanning": 0, "hamming": 1, "blackman": 2}
MB_FLAG_GENERIC_KERNEL = 1
MB_FLAG_EXACT_FFT = 2
MB_FLAG_CLUSTER_FFT = 4
MB_FLAG_NO_REFINE = 8
MB_NUM_FEATURES = 18
MB_SAMPLE_S16, MB_SAMPLE_S24, MB_SAMPLE_F32 = 1, 2, 3

# every symbol include/meyda_b200.h declares
EXPORTS = [
    "mb_version", "mb_last_error", "mb_feature_name", "mb_feature_from_name", "mb_device_count",
    "mb_num_frames", "mb_plan_create", "mb_plan_destroy", "mb_plan_set_stream", "mb_plan_tables",
    "mb_query_output", "mb_extract", "mb_extract_async", "mb_plan_synchronize", "mb_extract_multi",
    "mb_plan_launch_count", "mb_plan_kernel_name", "mb_host_alloc", "mb_host_free",
    "mb_stream_create", "mb_stream_destroy", "mb_stream_frames_after", "mb_stream_push", "mb_stream_reset",
    "mb_extract_pcm16", "mb_extract_pcm16_async", "mb_wav_parse", "mb_stream_graph_launches", "mb_extract_pcm",
    "mb_stream_create_pcm16", "mb_stream_push_pcm16", "mb_plan_create_ex", "mb_plan_get_params",
    "mb_plan_refined_frames", "mb_measure_peaks", "mb_set_host_threads", "mb_set_host_rows",
    "mb_get_host_rows",
]

# (field name in mb_outputs, feature name, per-frame length as a function of N and of the plan's Bark-band and
#  mfcc-coefficient counts: 24 and 13 unless the plan was created with other parameters)
OUTPUT_FIELDS = [
    ("buffer", "buffer", lambda N, nb=24, nc=13: N),
    ("rms", "rms", lambda N, nb=24, nc=13: 1),
    ("energy", "energy", lambda N, nb=24, nc=13: 1),
    ("zcr", "zcr", lambda N, nb=24, nc=13: 1),
    ("complex_real", "complexSpectrum", lambda N, nb=24, nc=13: N),
    ("complex_imag", "complexSpectrum", lambda N, nb=24, nc=13: N),
    ("amplitude_spectrum", "amplitudeSpectrum", lambda N, nb=24, nc=13: N // 2),
    ("power_spectrum", "powerSpectrum", lambda N, nb=24, nc=13: N // 2),
    ("spectral_centroid", "spectralCentroid", lambda N, nb=24, nc=13: 1),
    ("spectral_flatness", "spectralFlatness", lambda N, nb=24, nc=13: 1),
    ("spectral_slope", "spectralSlope", lambda N, nb=24, nc=13: 1),
    ("spectral_rolloff", "spectralRolloff", lambda N, nb=24, nc=13: 1),
    ("spectral_spread", "spectralSpread", lambda N, nb=24, nc=13: 1),
    ("spectral_skewness", "spectralSkewness", lambda N, nb=24, nc=13: 1),
    ("spectral_kurtosis", "spectralKurtosis", lambda N, nb=24, nc=13: 1),
    ("loudness_specific", "loudness", lambda N, nb=24, nc=13: nb),
    ("loudness_total", "loudness", lambda N, nb=24, nc=13: 1),
    ("perceptual_spread", "perceptualSpread", lambda N, nb=24, nc=13: 1),
    ("perceptual_sharpness", "perceptualSharpness", lambda N, nb=24, nc=13: 1),
    ("mfcc", "mfcc", lambda N, nb=24, nc=13: nc),
]


class Outputs(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _, _ in OUTPUT_FIELDS]


class Layout(C.Structure):
    _fields_ = [("total_frames", C.c_int64), ("buffer_size", C.c_int32), ("spectrum_size", C.c_int32),
                ("feature_mask", C.c_uint32), ("reserved", C.c_int32), ("output_bytes", C.c_int64),
                ("bytes_per_frame", C.c_int64), ("num_bark_bands", C.c_int32), ("num_mfcc", C.c_int32)]


class Params(C.Structure):
    """mb_params: what the reference keeps as constants (loudness.js:14, mfcc.js:15,71, spectralRolloff.js:9)."""
    _fields_ = [("num_bark_bands", C.c_int32), ("num_mel_filters", C.c_int32), ("num_mfcc", C.c_int32),
                ("reserved", C.c_int32), ("rolloff_fraction", C.c_double)]


class WavInfo(C.Structure):
    _fields_ = [("format", C.c_int32), ("channels", C.c_int32), ("sample_rate", C.c_int32),
                ("bits_per_sample", C.c_int32), ("data_offset", C.c_int64), ("n_sample_frames", C.c_int64)]


class MeydaNativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(message)
        self.status = status


_lib = None


def lib():
    """The loaded C-ABI library.  Raises if it was never built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "meyda_b200: %s is missing. Build it with `python -m meyda_b200.build` (needs nvcc); "
            "there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    i64p, vp = C.POINTER(C.c_int64), C.c_void_p
    L.mb_version.restype = C.c_int
    L.mb_last_error.restype = C.c_char_p
    L.mb_feature_name.restype = C.c_char_p
    L.mb_feature_name.argtypes = [C.c_int]
    L.mb_feature_from_name.argtypes = [C.c_char_p]
    L.mb_device_count.argtypes = [C.POINTER(C.c_int)]
    L.mb_num_frames.restype = C.c_int64
    L.mb_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
    L.mb_plan_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint32,
                                 C.c_uint32]
    L.mb_plan_create_ex.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_uint32,
                                    C.c_uint32, C.POINTER(Params)]
    L.mb_plan_get_params.argtypes = [vp, C.POINTER(Params)]
    L.mb_plan_destroy.argtypes = [vp]
    L.mb_plan_destroy.restype = None
    L.mb_plan_set_stream.argtypes = [vp, vp]
    L.mb_plan_tables.argtypes = [vp, vp, vp, vp]
    L.mb_query_output.argtypes = [vp, C.c_int64, i64p, i64p, C.POINTER(Layout)]
    L.mb_extract.argtypes = [vp, vp, C.c_int64, i64p, i64p, C.c_int64, C.POINTER(Outputs), C.c_int]
    L.mb_extract_async.argtypes = [vp, vp, C.c_int64, i64p, i64p, C.c_int64, C.POINTER(Outputs)]
    L.mb_plan_synchronize.argtypes = [vp]
    L.mb_extract_multi.argtypes = [C.POINTER(vp), C.c_int, vp, C.c_int64, i64p, i64p, C.c_int64, C.POINTER(Outputs)]
    L.mb_plan_launch_count.restype = C.c_int64
    L.mb_plan_launch_count.argtypes = [vp]
    L.mb_plan_refined_frames.argtypes = [vp, i64p]
    L.mb_measure_peaks.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.mb_set_host_threads.argtypes = [C.c_int]
    L.mb_set_host_rows.argtypes = [C.c_int]
    L.mb_get_host_rows.restype = C.c_int
    L.mb_get_host_rows.argtypes = []
    L.mb_plan_kernel_name.restype = C.c_char_p
    L.mb_plan_kernel_name.argtypes = [vp]
    L.mb_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.mb_host_free.argtypes = [vp]
    L.mb_host_free.restype = None
    L.mb_stream_create.argtypes = [C.POINTER(vp), vp]
    L.mb_stream_destroy.argtypes = [vp]
    L.mb_stream_destroy.restype = None
    L.mb_stream_frames_after.restype = C.c_int64
    L.mb_stream_frames_after.argtypes = [vp, C.c_int64]
    L.mb_stream_push.argtypes = [vp, vp, C.c_int64, C.POINTER(Outputs), C.c_int, i64p]
    L.mb_stream_reset.argtypes = [vp]
    L.mb_stream_create_pcm16.argtypes = [C.POINTER(vp), vp, C.c_int, C.c_int]
    L.mb_stream_push_pcm16.argtypes = [vp, vp, C.c_int64, C.POINTER(Outputs), C.c_int, i64p]
    L.mb_stream_graph_launches.restype = C.c_int64
    L.mb_stream_graph_launches.argtypes = [vp]
    L.mb_extract_pcm16.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, i64p, i64p, C.c_int64, C.POINTER(Outputs),
                                   C.c_int]
    L.mb_extract_pcm16_async.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_int, i64p, i64p, C.c_int64,
                                         C.POINTER(Outputs)]
    L.mb_wav_parse.argtypes = [vp, C.c_int64, C.POINTER(WavInfo)]
    L.mb_extract_pcm.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int, C.c_int, i64p, i64p, C.c_int64, C.POINTER(Outputs),
                                 C.c_int]
    _lib = L
    return L


def check(status: int):
    if status != MB_OK:
        raise MeydaNativeError(status, lib().mb_last_error().decode("utf-8", "replace"))


def feature_names():
    L = lib()
    return [L.mb_feature_name(i).decode() for i in range(MB_NUM_FEATURES)]
