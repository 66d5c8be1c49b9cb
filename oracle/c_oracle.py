"""ctypes binding of oracle/meyda_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Builds oracle/_build/libmeyda_oracle.so on demand (gcc, -ffp-contract=off).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmeyda_oracle.so")
SCALAR_NAMES = ["rms", "energy", "zcr", "spectralCentroid", "spectralFlatness", "spectralSlope",
                "spectralRolloff", "spectralSpread", "spectralSkewness", "spectralKurtosis",
                "loudness.total", "perceptualSpread", "perceptualSharpness"]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "meyda_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class _BatchOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in
                ("scalars", "specific", "mfcc", "buffer", "cs_real", "cs_imag", "amp", "power")]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.mo_plan_create.restype = C.c_void_p
        _lib.mo_plan_create.argtypes = [C.c_int, C.c_double, C.c_int]
        _lib.mo_plan_destroy.argtypes = [C.c_void_p]
        _lib.mo_plan_window.restype = C.POINTER(C.c_float)
        _lib.mo_plan_window.argtypes = [C.c_void_p, C.c_int]
        _lib.mo_plan_bark.restype = C.POINTER(C.c_float)
        _lib.mo_plan_bark.argtypes = [C.c_void_p]
        _lib.mo_plan_bb_limits.restype = C.POINTER(C.c_int32)
        _lib.mo_plan_bb_limits.argtypes = [C.c_void_p]
        _lib.mo_fft_jsfft.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        _lib.mo_num_frames.restype = C.c_int64
        _lib.mo_num_frames.argtypes = [C.c_int64, C.c_int, C.c_int]
        _lib.mo_extract_threads.restype = C.c_int64
        _lib.mo_extract_threads.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int,
                                            C.POINTER(_BatchOut), C.c_int, C.c_int64]
    return _lib


def fft(real: np.ndarray, imag: np.ndarray):
    re = np.array(real, dtype=np.float32, copy=True)
    im = np.array(imag, dtype=np.float32, copy=True)
    lib().mo_fft_jsfft(re.ctypes.data, im.ctypes.data, len(re))
    return re, im


def extract(signal, bufferSize: int, hop=None, sr: float = 44100.0, window: str = "hanning",
            arrays: bool = True, threads: int = 1, n_clips: int = 1, ring_per_thread: int = 0) -> dict:
    """Every feature of every frame of `n_clips` equal-length clips laid back to
    back in `signal`.  Same result shapes as meyda_oracle.extract."""
    L = lib()
    hop = bufferSize if hop is None else hop
    sig = np.ascontiguousarray(signal, dtype=np.float32)
    clip_len = len(sig) // n_clips
    plan = L.mo_plan_create(bufferSize, float(sr), {"hanning": 0, "hamming": 1, "blackman": 2}[window])
    if not plan:
        raise ValueError("Buffer size is not a power of two: Meyda will not run.")
    try:
        nf = L.mo_num_frames(clip_len, bufferSize, hop) * n_clips
        total = nf
        if ring_per_thread:  # timing runs: bounded output arena, results are overwritten
            nf = ring_per_thread * threads
        N, n = bufferSize, bufferSize // 2
        bufs = {"scalars": np.zeros((nf, 13), np.float64), "specific": np.zeros((nf, 24), np.float32),
                "mfcc": np.zeros((nf, 13), np.float32)}
        if arrays:
            bufs.update(buffer=np.zeros((nf, N), np.float32), cs_real=np.zeros((nf, N), np.float32),
                        cs_imag=np.zeros((nf, N), np.float32), amp=np.zeros((nf, n), np.float32),
                        power=np.zeros((nf, n), np.float32))
        bo = _BatchOut(**{k: v.ctypes.data for k, v in bufs.items()})
        got = L.mo_extract_threads(plan, sig.ctypes.data, n_clips, clip_len, hop, C.byref(bo), threads,
                                   ring_per_thread)
        assert got == total
    finally:
        L.mo_plan_destroy(plan)
    out = {"frames_processed": total} if ring_per_thread else {}
    for i, name in enumerate(SCALAR_NAMES):
        if name != "loudness.total":
            out[name] = bufs["scalars"][:, i].copy()
    out["loudness"] = {"specific": bufs["specific"], "total": bufs["scalars"][:, 10].copy()}
    out["mfcc"] = bufs["mfcc"]
    if arrays:
        out["buffer"] = bufs["buffer"]
        out["complexSpectrum"] = {"real": bufs["cs_real"], "imag": bufs["cs_imag"]}
        out["amplitudeSpectrum"] = bufs["amp"]
        out["powerSpectrum"] = bufs["power"]
    return out


def plan_tables(bufferSize: int, sr: float = 44100.0):
    L = lib()
    plan = L.mo_plan_create(bufferSize, float(sr), 0)
    try:
        han = np.ctypeslib.as_array(L.mo_plan_window(plan, 0), (bufferSize,)).copy()
        ham = np.ctypeslib.as_array(L.mo_plan_window(plan, 1), (bufferSize,)).copy()
        bark = np.ctypeslib.as_array(L.mo_plan_bark(plan), (bufferSize,)).copy()
        bb = np.ctypeslib.as_array(L.mo_plan_bb_limits(plan), (25,)).copy()
    finally:
        L.mo_plan_destroy(plan)
    return {"hanning": han, "hamming": ham, "bark": bark, "bbLimits": bb}
