"""Parity comparison between the CUDA path (field-name dict of arrays, as
meyda_b200.Plan.extract_host returns) and the oracle (feature-name dict, as
oracle.meyda_oracle.extract returns).  Tolerances are BASELINE.json's:

  zcr, buffer            bit-exact
  spectralRolloff        exact bin (discrete output)
  spectra                |gpu-ref| <= 1e-4 * max_k|ref_frame|  and per-bin relative
                         <= 1e-4 on bins >= 1e-3 * peak (float32 FFT error is absolute,
                         SURVEY.md section 7)
  numbers, loudness, mfcc   1e-3 relative OR absolute (slope: relative only -- its
                         magnitude is ~1e-7, an absolute 1e-3 would be vacuous)
  NaN / +-Inf            same positions and signs
"""
from __future__ import annotations

import numpy as np

SPECTRA_TOL = 1e-4
NUMBER_TOL = 1e-3

NUMBER_FIELDS = {
    "rms": "rms", "energy": "energy", "spectral_centroid": "spectralCentroid",
    "spectral_flatness": "spectralFlatness", "spectral_slope": "spectralSlope",
    "spectral_spread": "spectralSpread", "spectral_skewness": "spectralSkewness",
    "spectral_kurtosis": "spectralKurtosis", "perceptual_spread": "perceptualSpread",
    "perceptual_sharpness": "perceptualSharpness",
}


def _special_match(g, r):
    """NaN/Inf positions and signs identical; returns the finite mask."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    assert g.shape == r.shape, (g.shape, r.shape)
    fin_g, fin_r = np.isfinite(g), np.isfinite(r)
    bad = fin_g != fin_r
    assert not bad.any(), "finite/non-finite mismatch at %s: gpu=%s ref=%s" % (
        np.argwhere(bad)[:5].tolist(), g[bad][:5], r[bad][:5])
    nf = ~fin_r
    same = (np.isnan(g[nf]) & np.isnan(r[nf])) | (g[nf] == r[nf])
    assert same.all(), "NaN/Inf kind mismatch: gpu=%s ref=%s" % (g[nf][~same][:5], r[nf][~same][:5])
    return fin_r


def number_violations(g, r, tol=NUMBER_TOL, relative_only=False):
    """Indices whose error exceeds `tol` both relatively and absolutely."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    fin = _special_match(g, r)
    err = np.abs(np.where(fin, g - r, 0.0))
    rel_ok = err <= tol * np.abs(np.where(fin, r, 1.0))
    ok = rel_ok if relative_only else (rel_ok | (err <= tol))
    return np.argwhere(~ok), err


def assert_numbers(name, g, r, tol=NUMBER_TOL, relative_only=False, allow=0):
    bad, err = number_violations(g, r, tol, relative_only)
    assert len(bad) <= allow, "%s: %d values outside %g (allowed %d); worst err %g at %s (gpu=%s ref=%s)" % (
        name, len(bad), tol, allow, err.max(), bad[:3].tolist(),
        np.asarray(g)[tuple(bad[0])] if len(bad) else None, np.asarray(r)[tuple(bad[0])] if len(bad) else None)
    return len(bad)


def assert_spectrum(name, g, r, tol=SPECTRA_TOL, peak=None):
    """g, r: [frames, bins].  `peak` overrides the per-frame reference peak
    (complexSpectrum real/imag share the frame's complex peak)."""
    g = np.atleast_2d(np.asarray(g, dtype=np.float64))
    r = np.atleast_2d(np.asarray(r, dtype=np.float64))
    fin = _special_match(g, r)
    if g.size == 0:
        return
    rr = np.where(fin, r, 0.0)
    pk = np.abs(rr).max(axis=1, keepdims=True) if peak is None else np.asarray(peak, dtype=np.float64).reshape(-1, 1)
    err = np.abs(np.where(fin, g - r, 0.0))
    lim = tol * pk
    bad = err > lim
    assert not bad.any(), "%s: peak-relative error %g > %g at %s" % (
        name, (err / np.maximum(pk, 1e-300)).max(), tol, np.argwhere(bad)[:3].tolist())
    big = np.abs(rr) >= 1e-3 * pk
    rel = err[big] / np.abs(rr[big])
    assert rel.size == 0 or rel.max() <= tol, "%s: per-bin relative error %g > %g" % (name, rel.max(), tol)


def compare_all(gpu: dict, ref: dict, N: int, sr: float = 44100.0, allow_moment_outliers: int = 0) -> dict:
    """Compare whatever features `gpu` holds.  Returns {feature: outlier count}."""
    out = {}
    n = N // 2
    if "buffer" in gpu:
        assert np.array_equal(gpu["buffer"].view(np.uint32), ref["buffer"].view(np.uint32)), "buffer not bit-exact"
    if "zcr" in gpu:
        assert np.array_equal(gpu["zcr"].astype(np.int64), ref["zcr"].astype(np.int64)), "zcr not bit-exact"
    if "complex_real" in gpu:
        rr, ri = ref["complexSpectrum"]["real"], ref["complexSpectrum"]["imag"]
        pk = np.sqrt(rr.astype(np.float64) ** 2 + ri.astype(np.float64) ** 2).max(axis=1) if len(rr) else None
        assert_spectrum("complexSpectrum.real", gpu["complex_real"], rr, peak=pk)
        assert_spectrum("complexSpectrum.imag", gpu["complex_imag"], ri, peak=pk)
    if "amplitude_spectrum" in gpu:
        assert_spectrum("amplitudeSpectrum", gpu["amplitude_spectrum"], ref["amplitudeSpectrum"])
    if "power_spectrum" in gpu:
        assert_spectrum("powerSpectrum", gpu["power_spectrum"], ref["powerSpectrum"], tol=2.5 * SPECTRA_TOL)
    if "spectral_rolloff" in gpu:
        bin_hz = sr / (2 * (n - 1))
        gb = np.rint(gpu["spectral_rolloff"].astype(np.float64) / bin_hz)
        rb = np.rint(ref["spectralRolloff"] / bin_hz)
        assert np.array_equal(gb, rb), "rolloff bin mismatch at %s" % np.argwhere(gb != rb)[:5].tolist()
        assert_numbers("spectralRolloff", gpu["spectral_rolloff"], ref["spectralRolloff"], tol=1e-6, relative_only=True)
    for field, feat in NUMBER_FIELDS.items():
        if field in gpu:
            allow = allow_moment_outliers if feat in ("spectralSkewness", "spectralKurtosis", "spectralSpread") else 0
            out[feat] = assert_numbers(feat, gpu[field], ref[feat], relative_only=(feat == "spectralSlope"), allow=allow)
    if "loudness_specific" in gpu:
        assert_numbers("loudness.specific", gpu["loudness_specific"], ref["loudness"]["specific"])
        assert_numbers("loudness.total", gpu["loudness_total"], ref["loudness"]["total"])
    if "mfcc" in gpu:
        assert_numbers("mfcc", gpu["mfcc"], ref["mfcc"])
    return out
