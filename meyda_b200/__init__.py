"""meyda_b200 -- B200-native (sm_100a) implementation of Meyda's per-frame
feature-extraction hot path behind the reference's extractor API.

  csrc/      CUDA kernels + the C ABI of include/meyda_b200.h
  _capi.py   ctypes binding of that ABI
  meyda.py   host-side mirror of the reference interface (Meyda, get, featureInfo, extract)
  build.py   nvcc build of _lib/libmeyda_b200.so
"""
from .meyda import (AudioContext, ExtractResult, FEATURES, Meyda, MeydaError, MeydaNativeError, Plan, Stream,
                    extract, extract_multi, extract_wav, feature_mask, featureInfo, isPowerOfTwo, set_host_rows, get_host_rows,
                    set_host_threads, wav_info)

__all__ = ["AudioContext", "ExtractResult", "FEATURES", "Meyda", "MeydaError", "MeydaNativeError", "Plan", "Stream",
           "extract", "extract_multi", "extract_wav", "feature_mask", "featureInfo", "isPowerOfTwo", "set_host_rows", "get_host_rows",
           "set_host_threads", "wav_info"]
__version__ = "0.1.0"
