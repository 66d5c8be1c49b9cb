"""Host-side mirror of the reference's extractor API over the CUDA C ABI.

The reference (kirbysayshi/meyda v1.1.0) is JavaScript and no JS toolchain
exists in this image, so the host layer that would be `js/meyda_b200.js` over
an N-API addon is mirrored here in Python over ctypes, keeping the reference's
names, argument meaning and error behaviour:

  Meyda(audioContext, src, bufSize, callback)   src/meyda.js:17
  meyda.get(feature | [features])               src/meyda.js:244-261
  meyda.start(features) / stop()                src/meyda.js:233-241
  meyda.setSource(src)                          src/meyda.js:229-231
  meyda.windowingFunction = "hanning"|"hamming" src/meyda.js:41, docs.md:5-11
  meyda.featureInfo                             src/feature-info.js:3-64

plus the batched call the north star adds: extract(clips, bufferSize, hop, ...).
Every feature value comes from the CUDA kernels; nothing here computes audio
features on the CPU.
"""
from __future__ import annotations

import ctypes as C
import sys
from typing import Callable, Iterable, Sequence

import numpy as np

from . import _capi
from ._capi import MeydaNativeError, Outputs, OUTPUT_FIELDS

# src/feature-info.js:3-64 -- lower-case types as in the code (README spells
# them 'Number'/'Array'; readme.md:84-85).
featureInfo = {
    "buffer": {"type": "array"},
    "rms": {"type": "number"},
    "energy": {"type": "number"},
    "zcr": {"type": "number"},
    "complexSpectrum": {"type": "multipleArrays", "arrayNames": {"1": "real", "2": "imag"}},
    "amplitudeSpectrum": {"type": "array"},
    "powerSpectrum": {"type": "array"},
    "spectralCentroid": {"type": "number"},
    "spectralFlatness": {"type": "number"},
    "spectralSlope": {"type": "number"},
    "spectralRolloff": {"type": "number"},
    "spectralSpread": {"type": "number"},
    "spectralSkewness": {"type": "number"},
    "spectralKurtosis": {"type": "number"},
    "loudness": {"type": "multipleArrays", "arrayNames": {"1": "total", "2": "specific"}},
    "perceptualSpread": {"type": "number"},
    "perceptualSharpness": {"type": "number"},
    "mfcc": {"type": "array"},
}
FEATURES = list(featureInfo)
_FEATURE_BIT = {name: i for i, name in enumerate(FEATURES)}


class MeydaError(Exception):
    """JS `Error` thrown by the reference (src/meyda.js:20-26,259)."""


def isPowerOfTwo(num) -> bool:
    """src/utils.js:13-19."""
    try:
        num = float(num)
    except (TypeError, ValueError):
        return False
    if num != num:
        return False
    while num % 2 == 0 and num > 1:
        num /= 2
    return num == 1


def feature_mask(features: Iterable[str]) -> int:
    m = 0
    for f in features:
        m |= 1 << _FEATURE_BIT[f]
    return m


def _normalize_clips(clips):
    """-> (flat float32 samples, int64 offsets, int64 lengths)."""
    if isinstance(clips, dict):
        data = np.ascontiguousarray(clips["data"], dtype=np.float32)
        return data, np.ascontiguousarray(clips["offsets"], dtype=np.int64), np.ascontiguousarray(
            clips["lengths"], dtype=np.int64)
    if isinstance(clips, np.ndarray) and clips.ndim == 1:
        clips = [clips]
    if isinstance(clips, np.ndarray) and clips.ndim == 2:
        data = np.ascontiguousarray(clips, dtype=np.float32)
        n, L = data.shape
        return data.reshape(-1), np.arange(n, dtype=np.int64) * L, np.full(n, L, dtype=np.int64)
    arrs = [np.ascontiguousarray(c, dtype=np.float32).reshape(-1) for c in clips]
    lengths = np.array([len(a) for a in arrs], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64) if len(arrs) else np.zeros(0, np.int64)
    data = np.concatenate(arrs) if arrs else np.zeros(0, np.float32)
    return data, offsets, lengths


def set_host_threads(n: int) -> None:
    """Host threads a host-memory extract uses for the rows the device does not produce (`buffer`, powerSpectrum);
    0 = automatic (mb_set_host_threads, include/meyda_b200.h)."""
    _capi.check(_capi.lib().mb_set_host_threads(int(n)))


def set_host_rows(mode: int) -> None:
    """1: host-memory extracts produce the `buffer` / powerSpectrum rows on the host while the device works; 2: the
    mirrored half of complexSpectrum as well; 0: the device produces every row and all are copied back; -1 (default):
    2 where one device is visible and the host has twelve or more cores, else 1 (mb_set_host_rows,
    include/meyda_b200.h).  Same bits in every mode."""
    _capi.check(_capi.lib().mb_set_host_rows(int(mode)))


def get_host_rows() -> int:
    """The host-rows mode in force (0, 1 or 2: the automatic choice resolved; mb_get_host_rows)."""
    return int(_capi.lib().mb_get_host_rows())


def pinned_empty(shape, dtype) -> np.ndarray:
    """A numpy array over page-locked host memory (mb_host_alloc); freed with mb_host_free when collected."""
    import weakref
    L = _capi.lib()
    nbytes = max(1, int(np.prod(shape)) * np.dtype(dtype).itemsize)
    ptr = C.c_void_p()
    _capi.check(L.mb_host_alloc(C.byref(ptr), nbytes))
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, L.mb_host_free, C.c_void_p(ptr.value))  # (arr keeps buf alive through its base chain)
    return arr


class Plan:
    """One (device, bufferSize, hop, sampleRate, window, feature set): the
    tables `new Meyda(...)` precomputes, resident on the GPU."""

    def __init__(self, bufferSize: int, hop: int | None = None, sampleRate: float = 44100.0,
                 windowingFunction: str = "hanning", features: Sequence[str] = tuple(FEATURES),
                 device: int = 0, flags: int = 0, numBarkBands: int | None = None, numMelFilters: int | None = None,
                 numMfccCoefficients: int | None = None, rolloffFraction: float | None = None):
        """numBarkBands is the NUM_BARK_BANDS option of the reference's Loudness constructor
        (src/extractors/loudness.js:14); numMelFilters, numMfccCoefficients and rolloffFraction replace the local
        constants 26, 13 (mfcc.js:15,71) and 0.99 (spectralRolloff.js:9).  None keeps the reference's value."""
        if not isPowerOfTwo(bufferSize):
            raise MeydaError("Buffer size is not a power of two: Meyda will not run.")
        if windowingFunction not in _capi.MB_WINDOW:
            raise MeydaError("unknown windowingFunction %r" % (windowingFunction,))
        self.bufferSize = int(bufferSize)
        self.hop = int(bufferSize if hop is None else hop)
        self.sampleRate = float(sampleRate)
        self.windowingFunction = windowingFunction
        self.features = list(features)
        self.device = device
        self.mask = feature_mask(self.features)
        self._L = _capi.lib()
        self._h = C.c_void_p()
        prm = _capi.Params(int(numBarkBands or 0), int(numMelFilters or 0), int(numMfccCoefficients or 0), 0,
                           float(rolloffFraction or 0.0))
        _capi.check(self._L.mb_plan_create_ex(C.byref(self._h), device, self.bufferSize, self.hop, self.sampleRate,
                                              _capi.MB_WINDOW[windowingFunction], self.mask, flags, C.byref(prm)))
        _capi.check(self._L.mb_plan_get_params(self._h, C.byref(prm)))
        self.numBarkBands, self.numMelFilters = int(prm.num_bark_bands), int(prm.num_mel_filters)
        self.numMfccCoefficients, self.rolloffFraction = int(prm.num_mfcc), float(prm.rolloff_fraction)

    def close(self):
        if getattr(self, "_h", None):
            self._L.mb_plan_destroy(self._h)
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    @property
    def kernel_name(self) -> str:
        return self._L.mb_plan_kernel_name(self._h).decode()

    @property
    def launch_count(self) -> int:
        return int(self._L.mb_plan_launch_count(self._h))

    @property
    def refined_frames(self) -> int:
        """Frames of the last extract call (or stream push) that were redone with the reference's own FFT
        arithmetic because their features sit in its rounding noise (adaptive float32 plans; else 0)."""
        n = C.c_int64(0)
        _capi.check(self._L.mb_plan_refined_frames(self._h, C.byref(n)))
        return int(n.value)

    def set_stream(self, cuda_stream: int | None):
        """Launch on this cudaStream_t handle; None restores the plan's own stream.  Handle 0 (the legacy
        default stream, e.g. torch's default `current_stream().cuda_stream`) is passed as cudaStreamLegacy,
        because a NULL handle means "the plan's own stream" at the C ABI.  Device-memory calls are ordered
        only against work on the stream they run on: buffers produced on another stream (a `torch.zeros`
        on torch's stream, say) must be synchronised by the caller or share the stream."""
        if cuda_stream == 0:
            cuda_stream = 1  # cudaStreamLegacy
        _capi.check(self._L.mb_plan_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        _capi.check(self._L.mb_plan_synchronize(self._h))

    def tables(self) -> dict:
        win = np.zeros(self.bufferSize, np.float32)
        bb = np.zeros(self.numBarkBands + 1, np.int32)
        mel = np.zeros(self.numMelFilters + 2, np.int32)
        _capi.check(self._L.mb_plan_tables(self._h, win.ctypes.data, bb.ctypes.data, mel.ctypes.data))
        return {"window": win, "bbLimits": bb, "melBins": mel}

    def query(self, lengths: np.ndarray):
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        per = np.zeros(len(lengths), np.int64)
        lay = _capi.Layout()
        _capi.check(self._L.mb_query_output(self._h, len(lengths), lengths.ctypes.data_as(C.POINTER(C.c_int64)),
                                            per.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(lay)))
        return per, lay

    def output_shapes(self, total_frames: int) -> dict:
        """field name -> (shape, dtype) for the requested features."""
        shapes = {}
        for field, feat, per in OUTPUT_FIELDS:
            if feat in self.features:
                n = per(self.bufferSize, self.numBarkBands, self.numMfccCoefficients)
                scalar = n == 1 and field not in ("loudness_specific", "mfcc")
                shapes[field] = ((total_frames,) if scalar else (total_frames, n),
                                 np.int32 if field == "zcr" else np.float32)
        return shapes

    def alloc_host_outputs(self, total_frames: int, pinned: bool | None = None) -> dict:
        """Output arrays for `total_frames` frames.  pinned (the default for results of a megabyte and more): page-locked
        memory from mb_host_alloc, which lets the device copy straight into them asynchronously (pageable arrays make
        every copy a staged, synchronous one); the memory is returned by mb_host_free when the array is collected."""
        shapes = self.output_shapes(total_frames)
        if pinned is None:
            pinned = sum(int(np.prod(s)) * 4 for s, _ in shapes.values()) >= (1 << 20)
        if not pinned:
            return {k: np.zeros(s, dtype=d) for k, (s, d) in shapes.items()}
        return {k: pinned_empty(s, d) for k, (s, d) in shapes.items()}

    @staticmethod
    def pack_outputs(ptrs: dict) -> Outputs:
        o = Outputs()
        for k, v in ptrs.items():
            setattr(o, k, v)
        return o

    def _check_out(self, out: dict, total_frames: int):
        """Caller-supplied output arrays must be what the C ABI will write: shape, dtype, C-contiguous."""
        for k, (shape, dtype) in self.output_shapes(total_frames).items():
            a = out.get(k)
            if not isinstance(a, np.ndarray) or a.shape != shape or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
                raise TypeError("output %r must be a C-contiguous %s array of shape %s" % (k, np.dtype(dtype).name, shape))

    def extract_host(self, data: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, out: dict | None = None):
        """MB_MEM_HOST call: numpy in, numpy out (dict field -> array)."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        per, lay = self.query(lengths)
        if out is None:
            out = self.alloc_host_outputs(int(lay.total_frames))
        else:
            self._check_out(out, int(lay.total_frames))
        o = self.pack_outputs({k: v.ctypes.data for k, v in out.items()})
        i64p = C.POINTER(C.c_int64)
        _capi.check(self._L.mb_extract(self._h, data.ctypes.data, data.size, offsets.ctypes.data_as(i64p),
                                       lengths.ctypes.data_as(i64p), len(lengths), C.byref(o), _capi.MB_MEM_HOST))
        return out, per

    def extract_pcm16_host(self, pcm: np.ndarray, offsets: np.ndarray, lengths: np.ndarray, channel: int = 0,
                           out: dict | None = None):
        """MB_MEM_HOST call on 16-bit PCM: `pcm` is int16 [sample_frames] (mono) or [sample_frames, channels]
        (interleaved, as a WAV data chunk holds it); offsets/lengths count sample frames.  The int16 -> float32
        conversion of decodeAudioData (s / 32768) and the channel pick happen inside the kernels' frame load."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        channels = 1 if pcm.ndim == 1 else pcm.shape[1]
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        per, lay = self.query(lengths)
        if out is None:
            out = self.alloc_host_outputs(int(lay.total_frames))
        o = self.pack_outputs({k: v.ctypes.data for k, v in out.items()})
        i64p = C.POINTER(C.c_int64)
        _capi.check(self._L.mb_extract_pcm16(self._h, pcm.ctypes.data, pcm.shape[0], channels, channel,
                                             offsets.ctypes.data_as(i64p), lengths.ctypes.data_as(i64p), len(lengths),
                                             C.byref(o), _capi.MB_MEM_HOST))
        return out, per

    def extract_pcm_host(self, data: np.ndarray, sample_format: int, channels: int, offsets: np.ndarray,
                         lengths: np.ndarray, channel: int = 0, out: dict | None = None):
        """MB_MEM_HOST call on a WAV payload of any supported sample format (mb_extract_pcm): `data` is the
        interleaved sample block as int16 / float32 values or, for packed 24-bit, as bytes (uint8);
        sample_format is _capi.MB_SAMPLE_S16 / _S24 / _F32.  offsets/lengths count sample frames."""
        data = np.ascontiguousarray(data)
        bytes_per = {_capi.MB_SAMPLE_S16: 2, _capi.MB_SAMPLE_S24: 3}.get(sample_format, 4)  # (the ABI rejects unknown formats)
        n_sample_frames = data.nbytes // (bytes_per * channels)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        per, lay = self.query(lengths)
        if out is None:
            out = self.alloc_host_outputs(int(lay.total_frames))
        o = self.pack_outputs({k: v.ctypes.data for k, v in out.items()})
        i64p = C.POINTER(C.c_int64)
        _capi.check(self._L.mb_extract_pcm(self._h, data.ctypes.data, sample_format, n_sample_frames, channels, channel,
                                           offsets.ctypes.data_as(i64p), lengths.ctypes.data_as(i64p), len(lengths),
                                           C.byref(o), _capi.MB_MEM_HOST))
        return out, per

    def extract_pcm16_device(self, pcm_ptr: int, n_sample_frames: int, channels: int, channel: int,
                             offsets: np.ndarray, lengths: np.ndarray, out_ptrs: dict, sync: bool = True):
        """MB_MEM_DEVICE call on 16-bit PCM already resident on the plan's device."""
        o = self.pack_outputs(out_ptrs)
        i64p = C.POINTER(C.c_int64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        args = [self._h, C.c_void_p(pcm_ptr), n_sample_frames, channels, channel, offsets.ctypes.data_as(i64p),
                lengths.ctypes.data_as(i64p), len(lengths), C.byref(o)]
        if sync:
            _capi.check(self._L.mb_extract_pcm16(*args, _capi.MB_MEM_DEVICE))
        else:
            _capi.check(self._L.mb_extract_pcm16_async(*args))

    def extract_device(self, samples_ptr: int, n_samples: int, offsets: np.ndarray, lengths: np.ndarray,
                       out_ptrs: dict, sync: bool = True):
        """MB_MEM_DEVICE call on raw device pointers (e.g. torch tensors' data_ptr())."""
        o = self.pack_outputs(out_ptrs)
        i64p = C.POINTER(C.c_int64)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        lengths = np.ascontiguousarray(lengths, dtype=np.int64)
        fn = self._L.mb_extract if sync else self._L.mb_extract_async
        args = [self._h, C.c_void_p(samples_ptr), n_samples, offsets.ctypes.data_as(i64p),
                lengths.ctypes.data_as(i64p), len(lengths), C.byref(o)]
        if sync:
            args.append(_capi.MB_MEM_DEVICE)
        _capi.check(fn(*args))


def extract_multi(plans: Sequence[Plan], data: np.ndarray, offsets: np.ndarray, lengths: np.ndarray,
                  out: dict | None = None):
    """Clip-sharded MB_MEM_HOST extraction over several devices (no collectives)."""
    p0 = plans[0]
    data = np.ascontiguousarray(data, dtype=np.float32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lengths = np.ascontiguousarray(lengths, dtype=np.int64)
    per, lay = p0.query(lengths)
    if out is None:
        out = p0.alloc_host_outputs(int(lay.total_frames))
    else:
        p0._check_out(out, int(lay.total_frames))
    o = Plan.pack_outputs({k: v.ctypes.data for k, v in out.items()})
    handles = (C.c_void_p * len(plans))(*[p.handle for p in plans])
    i64p = C.POINTER(C.c_int64)
    _capi.check(p0._L.mb_extract_multi(handles, len(plans), data.ctypes.data, data.size,
                                       offsets.ctypes.data_as(i64p), lengths.ctypes.data_as(i64p), len(lengths),
                                       C.byref(o)))
    return out, per


class ExtractResult:
    """Per-feature typed arrays for every frame, plus `get`-shaped frame views."""

    def __init__(self, features, arrays: dict, frames_per_clip: np.ndarray, bufferSize: int):
        self.features = list(features)
        self.arrays = arrays
        self.frames_per_clip = frames_per_clip
        self.clip_frame_start = np.concatenate([[0], np.cumsum(frames_per_clip)]).astype(np.int64)
        self.total_frames = int(self.clip_frame_start[-1])
        self.bufferSize = bufferSize

    def __len__(self):
        return self.total_frames

    def __getitem__(self, feature: str):
        """Whole-batch arrays of one feature, in the reference's value shape."""
        a = self.arrays
        if feature == "complexSpectrum":
            return {"real": a["complex_real"], "imag": a["complex_imag"]}
        if feature == "loudness":
            return {"specific": a["loudness_specific"], "total": a["loudness_total"]}
        for field, feat, _ in OUTPUT_FIELDS:
            if feat == feature:
                return a[field]
        raise KeyError(feature)

    def value(self, i: int, feature: str):
        """What `get(feature)` returned for frame i (number | array | object)."""
        a = self.arrays
        if feature == "complexSpectrum":
            return {"real": a["complex_real"][i], "imag": a["complex_imag"][i]}
        if feature == "loudness":
            return {"specific": a["loudness_specific"][i], "total": float(a["loudness_total"][i])}
        v = self[feature][i]
        if featureInfo[feature]["type"] == "number":
            return int(v) if feature == "zcr" else float(v)
        return v

    def frame(self, i: int, features=None) -> dict:
        """What `get([...features])` returned for frame i."""
        return {f: self.value(i, f) for f in (self.features if features is None else features)}

    def clip(self, c: int) -> range:
        return range(int(self.clip_frame_start[c]), int(self.clip_frame_start[c + 1]))


def _split_features(features):
    """List form: unknown names are reported and dropped (src/meyda.js:248-254).
    String form: unknown name is a TypeError, as calling undefined is in JS."""
    if isinstance(features, str):
        if features not in featureInfo:
            raise TypeError("Cannot read property 'process' of undefined (feature %r)" % (features,))
        return [features], True
    if isinstance(features, (list, tuple)):
        known = []
        for f in features:
            if isinstance(f, str) and f in featureInfo:
                if f not in known:
                    known.append(f)
            else:
                print("TypeError: unknown feature %r" % (f,), file=sys.stderr)  # console.error(e)
        return known, False
    raise MeydaError("Invalid Feature Format")


def extract(clips, bufferSize: int, hop: int | None = None, sampleRate: float = 44100.0,
            windowingFunction: str = "hanning", features=tuple(FEATURES),
            callback: Callable[[dict], None] | None = None, devices: Sequence[int] | None = None,
            flags: int = 0, **params) -> ExtractResult:
    """Batched drop-in for "construct Meyda, feed every buffer, get(features)".
    **params: numBarkBands / numMelFilters / numMfccCoefficients / rolloffFraction (see Plan).

    clips: 1-D array, list of 1-D arrays, 2-D [clips, samples] array, or
    {"data", "offsets", "lengths"}.  If `callback` is given it is invoked once
    per frame, in order, with the `get([...])`-shaped object (the
    buffer-by-buffer contract, src/meyda.js:87-89)."""
    feats, _single = _split_features(features)
    if not feats:
        raise MeydaError("Invalid Feature Format")
    data, offsets, lengths = _normalize_clips(clips)
    devices = [0] if devices is None else list(devices)
    plans = [Plan(bufferSize, hop, sampleRate, windowingFunction, feats, device=d, flags=flags, **params)
             for d in devices]
    try:
        if len(plans) == 1:
            arrays, per = plans[0].extract_host(data, offsets, lengths)
        else:
            arrays, per = extract_multi(plans, data, offsets, lengths)
    finally:
        for p in plans:
            p.close()
    res = ExtractResult(feats, arrays, per, int(bufferSize))
    if callback is not None:
        for i in range(res.total_frames):
            callback(res.frame(i))
    return res


def wav_info(file_bytes) -> dict:
    """RIFF/WAVE header of an in-memory file (mb_wav_parse): format, channels, sampleRate, bitsPerSample,
    dataOffset, sampleFrames."""
    buf = np.frombuffer(bytes(file_bytes), dtype=np.uint8)
    info = _capi.WavInfo()
    _capi.check(_capi.lib().mb_wav_parse(buf.ctypes.data, buf.size, C.byref(info)))
    return {"format": info.format, "channels": info.channels, "sampleRate": info.sample_rate,
            "bitsPerSample": info.bits_per_sample, "dataOffset": info.data_offset,
            "sampleFrames": info.n_sample_frames}


def extract_wav(files, bufferSize: int, hop: int | None = None, windowingFunction: str = "hanning",
                features=tuple(FEATURES), channel: int = 0, callback: Callable[[dict], None] | None = None,
                device: int = 0, flags: int = 0) -> ExtractResult:
    """Replaces BufferLoader + decodeAudioData + getChannelData(channel) (lib/bufferLoader.js:13-44,
    src/meyda.js:72) in front of `extract`: `files` is one path / bytes object or a list of them, each a
    16-bit PCM WAV (or 24-bit PCM / 32-bit float, which take the generic kernels); the samples go to the GPU as the
    file holds them and are converted (s / 32768, s / 8388608) inside the frame load.  All files must share the
    channel count, sample format and sample rate (which becomes the plan's sampleRate)."""
    feats, _single = _split_features(features)
    if not feats:
        raise MeydaError("Invalid Feature Format")
    if isinstance(files, (str, bytes, bytearray, memoryview)):
        files = [files]
    blobs = []
    for f in files:
        if isinstance(f, str):
            with open(f, "rb") as fh:
                f = fh.read()
        blobs.append(bytes(f))
    infos = [wav_info(b) for b in blobs]
    kinds = {(1, 16): (_capi.MB_SAMPLE_S16, 2), (1, 24): (_capi.MB_SAMPLE_S24, 3), (3, 32): (_capi.MB_SAMPLE_F32, 4)}
    for i in infos:
        if (i["format"], i["bitsPerSample"]) not in kinds:
            raise MeydaError("extract_wav takes 16- or 24-bit integer PCM or 32-bit float (got format %d, %d bits)"
                             % (i["format"], i["bitsPerSample"]))
        if (i["channels"], i["sampleRate"], i["format"], i["bitsPerSample"]) != \
                (infos[0]["channels"], infos[0]["sampleRate"], infos[0]["format"], infos[0]["bitsPerSample"]):
            raise MeydaError("all WAV files of one call must share channel count, sample rate and sample format")
    ch = infos[0]["channels"]
    fmt, sbytes = kinds[(infos[0]["format"], infos[0]["bitsPerSample"])]
    fb = sbytes * ch  # bytes per sample frame
    # clips start on multiples of 8 sample frames so that mono 16-bit frames stay 16-byte aligned (TMA bulk loads)
    lengths = np.array([i["sampleFrames"] for i in infos], dtype=np.int64)
    padded = (lengths + 7) // 8 * 8
    offsets = np.concatenate([[0], np.cumsum(padded)[:-1]]).astype(np.int64)
    payload = np.zeros(int(padded.sum()) * fb, dtype=np.uint8)  # the data chunks, back to back, untouched
    for o, b, i in zip(offsets, blobs, infos):
        n = int(i["sampleFrames"]) * fb
        payload[int(o) * fb:int(o) * fb + n] = np.frombuffer(b, dtype=np.uint8, count=n, offset=i["dataOffset"])
    plan = Plan(bufferSize, hop, float(infos[0]["sampleRate"]), windowingFunction, feats, device=device, flags=flags)
    try:
        arrays, per = plan.extract_pcm_host(payload, fmt, ch, offsets, lengths, channel=channel)
    finally:
        plan.close()
    res = ExtractResult(feats, arrays, per, int(bufferSize))
    if callback is not None:
        for i in range(res.total_frames):
            callback(res.frame(i))
    return res


class AudioContext:
    """Stand-in for the Web Audio context the reference constructor needs:
    only `.sampleRate` is read (src/meyda.js:29)."""

    def __init__(self, sampleRate: float = 44100.0):
        self.sampleRate = float(sampleRate)


class Meyda:
    """The reference's class (src/meyda.js:15-263) over the CUDA path.

    `src` is the source signal (1-D float array) instead of a Web Audio node.
    `process()` plays it through: every complete buffer is framed on the GPU in
    one batch, then delivered buffer by buffer -- `get()` sees the current
    buffer, and `callback(get(features))` fires per buffer while started.
    """

    def __init__(self, audioContext, src, bufSize, callback=None, hop=None, device: int = 0):
        if not isPowerOfTwo(bufSize):  # src/meyda.js:20-22
            raise MeydaError("Buffer size is not a power of two: Meyda will not run.")
        if not audioContext:  # src/meyda.js:24-26
            raise MeydaError("AudioContext wasn't specified: Meyda will not run.")
        self.audioContext = audioContext
        self.bufferSize = int(bufSize)
        self.hop = int(bufSize if hop is None else hop)
        self.sampleRate = float(audioContext.sampleRate)
        self.windowingFunction = "hanning"  # src/meyda.js:41
        self.featureInfo = featureInfo
        self.featureExtractors = {name: (lambda n=name: self.get(n)) for name in FEATURES}
        self.EXTRACTION_STARTED = False
        self._featuresToExtract = None
        self._callback = callback
        self._device = device
        self._source = None
        self._result = None
        self._result_window = None
        self._cursor = -1
        self.signal = None
        self.setSource(src)

    def setSource(self, _src):
        self._source = None if _src is None else np.ascontiguousarray(_src, dtype=np.float32).reshape(-1)
        self._result = None
        self._cursor = -1

    def start(self, features):
        self._featuresToExtract = features
        self.EXTRACTION_STARTED = True

    def stop(self):
        self._featuresToExtract = None
        self.EXTRACTION_STARTED = False

    def _ensure_result(self):
        if self._result is None or self._result_window != self.windowingFunction:
            if self._source is None:
                raise MeydaError("no source set")
            self._result = extract(self._source, self.bufferSize, self.hop, self.sampleRate, self.windowingFunction,
                                   FEATURES, devices=[self._device])
            self._result_window = self.windowingFunction

    def process(self, max_buffers: int | None = None) -> int:
        """Deliver the source buffer by buffer (the onaudioprocess loop,
        src/meyda.js:69-91).  Returns the number of buffers delivered."""
        self._ensure_result()
        n = self._result.total_frames
        done = 0
        while self._cursor + 1 < n and (max_buffers is None or done < max_buffers):
            self._cursor += 1
            self.signal = self._result.arrays["buffer"][self._cursor]
            if callable(self._callback) and self.EXTRACTION_STARTED:
                self._callback(self.get(self._featuresToExtract))
            done += 1
        return done

    def get(self, feature):
        """src/meyda.js:244-261."""
        if isinstance(feature, (list, tuple)):
            self._require_buffer()
            results = {}
            for name in feature:
                try:
                    if not isinstance(name, str) or name not in featureInfo:
                        raise TypeError("Cannot read property 'process' of undefined (feature %r)" % (name,))
                    results[name] = self._result.value(self._cursor, name)
                except TypeError as e:
                    print("%s: %s" % (type(e).__name__, e), file=sys.stderr)  # console.error(e)
            return results
        if isinstance(feature, str):
            if feature not in featureInfo:
                raise TypeError("Cannot read property 'process' of undefined (feature %r)" % (feature,))
            self._require_buffer()
            return self._result.value(self._cursor, feature)
        raise MeydaError("Invalid Feature Format")

    def _require_buffer(self):
        self._ensure_result()
        if self._cursor < 0:
            if self._result.total_frames == 0:
                raise MeydaError("source is shorter than one buffer")
            self._cursor = 0
            self.signal = self._result.arrays["buffer"][0]


class Stream:
    """Stateful buffer-by-buffer extractor (mb_stream_*): push samples, get the
    features of every frame they complete."""

    def __init__(self, plan: Plan, pcm16_channels: int = 0, channel: int = 0):
        """pcm16_channels > 0: the stream takes int16 blocks of that many interleaved channels (a capture device's
        or WAV reader's format) and `channel` of them is analysed; the conversion happens in the kernels."""
        self.plan = plan
        self._L = plan._L
        self._h = C.c_void_p()
        self.pcm16_channels = int(pcm16_channels)
        if self.pcm16_channels:
            _capi.check(self._L.mb_stream_create_pcm16(C.byref(self._h), plan.handle, self.pcm16_channels, channel))
        else:
            _capi.check(self._L.mb_stream_create(C.byref(self._h), plan.handle))

    def close(self):
        if getattr(self, "_h", None):
            # the native stream points into its plan: if the plan is already gone (garbage collection order at
            # interpreter exit, or a plan closed first) there is nothing safe left to release
            if getattr(self.plan, "_h", None):
                self._L.mb_stream_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self):
        _capi.check(self._L.mb_stream_reset(self._h))

    @property
    def graph_launches(self) -> int:
        """Pushes replayed as a CUDA graph so far (a repeating push shape is captured on its second occurrence)."""
        return int(self._L.mb_stream_graph_launches(self._h))

    def push_into(self, samples: np.ndarray, out: dict) -> int:
        """push() into caller-owned arrays (from Plan.alloc_host_outputs); returns the frames completed."""
        o = Plan.pack_outputs({k: v.ctypes.data for k, v in out.items()})
        done = C.c_int64(0)
        _capi.check(self._L.mb_stream_push(self._h, samples.ctypes.data, samples.size, C.byref(o), _capi.MB_MEM_HOST,
                                           C.byref(done)))
        return int(done.value)

    def push(self, samples) -> ExtractResult:
        if self.pcm16_channels:
            x = np.ascontiguousarray(samples, dtype=np.int16).reshape(-1, self.pcm16_channels)
            n, fn = x.shape[0], self._L.mb_stream_push_pcm16
        else:
            x = np.ascontiguousarray(samples, dtype=np.float32).reshape(-1)
            n, fn = x.size, self._L.mb_stream_push
        nf = int(self._L.mb_stream_frames_after(self._h, n))
        out = self.plan.alloc_host_outputs(nf)
        o = Plan.pack_outputs({k: v.ctypes.data for k, v in out.items()})
        done = C.c_int64(0)
        _capi.check(fn(self._h, x.ctypes.data, n, C.byref(o), _capi.MB_MEM_HOST, C.byref(done)))
        assert done.value == nf
        return ExtractResult(self.plan.features, out, np.array([nf], np.int64), self.plan.bufferSize)


__all__ = ["Meyda", "AudioContext", "MeydaError", "MeydaNativeError", "featureInfo", "FEATURES", "extract",
           "extract_multi", "Plan", "Stream", "ExtractResult", "isPowerOfTwo", "feature_mask"]
