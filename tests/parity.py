"""Parity comparison between the CUDA path (field-name dict of arrays, as
meyda_b200.Plan.extract_host returns) and the oracle (feature-name dict, as
oracle.meyda_oracle.extract returns).  Tolerances are BASELINE.json's:

  zcr, buffer            bit-exact
  spectralRolloff        exact bin (discrete output)
  spectra                |gpu-ref| <= 1e-4 * max_k|ref_frame|  and per-bin relative
                         <= 1e-4 on bins >= 1e-2 * peak (float32 FFT error is absolute,
                         ~1e-8 * peak per bin: SURVEY.md section 7)
  numbers, loudness, mfcc   1e-3 relative OR absolute (slope: relative, or 1e-13
                         absolute -- its magnitude is ~1e-7, an absolute 1e-3 would
                         be vacuous, and on a flat spectrum it is pure cancellation;
                         rms / energy: relative only; flatness: the absolute floor is
                         1e-3 x the batch's median flatness)
  NaN / +-Inf            same positions and signs

Noise band (float32-FFT kernels only).  Some features amplify the spectrum's
rounding-noise floor without bound: x^0.23 and ln(x) of a band that holds
nothing but FFT rounding noise, or k^3/k^4-weighted sums over ~n empty bins of
a pure tone.  On such frames the REFERENCE's own value is set by its float32
per-stage rounding (lib/jsfft/fft.js:158-161), and no FFT that is not bit
identical can land within 1e-3 of it.  `noise_band` (oracle.noise_band) is, per
value, how far the feature moves when the exact spectrum is perturbed by noise
of the reference's own measured level; a value may miss the tolerance only if
it is within NOISE_BAND_FACTOR x that band.  Such values are counted and
returned, never hidden; the exact-FFT mode (MB_FLAG_EXACT_FFT) must need none
and gets bit-identical spectra.
"""
from __future__ import annotations

import numpy as np

SPECTRA_TOL = 1e-4
NUMBER_TOL = 1e-3
SLOPE_ABS_TOL = 1e-13
NOISE_BAND_FACTOR = 8.0

NUMBER_FIELDS = {
    "rms": "rms", "energy": "energy", "spectral_centroid": "spectralCentroid",
    "spectral_flatness": "spectralFlatness", "spectral_slope": "spectralSlope",
    "spectral_spread": "spectralSpread", "spectral_skewness": "spectralSkewness",
    "spectral_kurtosis": "spectralKurtosis", "perceptual_spread": "perceptualSpread",
    "perceptual_sharpness": "perceptualSharpness",
}


def _special_match(g, r, name=""):
    """NaN/Inf positions and signs identical; returns the finite mask."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    assert g.shape == r.shape, (name, g.shape, r.shape)
    fin_g, fin_r = np.isfinite(g), np.isfinite(r)
    bad = fin_g != fin_r
    assert not bad.any(), "%s: finite/non-finite mismatch at %s: gpu=%s ref=%s" % (
        name, np.argwhere(bad)[:5].tolist(), g[bad][:5], r[bad][:5])
    nf = ~fin_r
    same = (np.isnan(g[nf]) & np.isnan(r[nf])) | (g[nf] == r[nf])
    assert same.all(), "%s: NaN/Inf kind mismatch: gpu=%s ref=%s" % (name, g[nf][~same][:5], r[nf][~same][:5])
    return fin_r


def assert_numbers(name, g, r, tol=NUMBER_TOL, abs_tol=None, band=None):
    """Pass: relative error <= tol, or absolute error <= abs_tol (default tol),
    or inside NOISE_BAND_FACTOR x band.  Returns how many values needed the band."""
    g = np.asarray(g, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    fin = _special_match(g, r, name)
    err = np.abs(np.where(fin, g - r, 0.0))
    ok = (err <= tol * np.abs(np.where(fin, r, 1.0))) | (err <= (tol if abs_tol is None else abs_tol))
    banded = 0
    if band is not None and not ok.all():
        in_band = ~ok & (err <= NOISE_BAND_FACTOR * np.asarray(band, dtype=np.float64))
        banded = int(in_band.sum())
        ok = ok | in_band
    bad = np.argwhere(~ok)
    assert len(bad) == 0, "%s: %d values outside %g; worst err %g at %s (gpu=%s ref=%s band=%s)" % (
        name, len(bad), tol, err[~ok].max(), bad[:3].tolist(), g[tuple(bad[0])], r[tuple(bad[0])],
        None if band is None else np.asarray(band)[tuple(bad[0])])
    return banded


def assert_spectrum(name, g, r, tol=SPECTRA_TOL, peak=None):
    """g, r: [frames, bins].  `peak` overrides the per-frame reference peak
    (complexSpectrum real/imag share the frame's complex peak)."""
    g = np.atleast_2d(np.asarray(g, dtype=np.float64))
    r = np.atleast_2d(np.asarray(r, dtype=np.float64))
    fin = _special_match(g, r, name)
    if g.size == 0:
        return
    rr = np.where(fin, r, 0.0)
    pk = np.abs(rr).max(axis=1, keepdims=True) if peak is None else np.asarray(peak, dtype=np.float64).reshape(-1, 1)
    err = np.abs(np.where(fin, g - r, 0.0))
    bad = err > tol * pk
    assert not bad.any(), "%s: peak-relative error %g > %g at %s" % (
        name, (err / np.maximum(pk, 1e-300)).max(), tol, np.argwhere(bad)[:3].tolist())
    big = (np.abs(rr) >= 1e-2 * pk) & (np.abs(rr) > 0)
    rel = err[big] / np.abs(rr[big])
    assert rel.size == 0 or rel.max() <= tol, "%s: per-bin relative error %g > %g" % (name, rel.max(), tol)


def assert_bits(name, g, r):
    g, r = np.asarray(g), np.asarray(r, dtype=g.dtype)
    same = (g.view(np.uint32) == r.view(np.uint32)) | (np.isnan(g) & np.isnan(r))
    assert same.all(), "%s: %d of %d values not bit-identical, first at %s (gpu=%r ref=%r)" % (
        name, int((~same).sum()), same.size, np.argwhere(~same)[:3].tolist(), g[~same][:3], r[~same][:3])


def compare_all(gpu: dict, ref: dict, N: int, sr: float = 44100.0, noise_band: dict | None = None,
                exact: bool = False) -> dict:
    """Compare whatever features `gpu` holds.  Returns {feature: values that
    needed the noise band}.  exact=True: spectra must be bit-identical and
    every number within 5e-6 (float32 output rounding + libm)."""
    out = {}
    n = N // 2
    tol = 5e-6 if exact else NUMBER_TOL
    nb = (lambda k: None) if (noise_band is None or exact) else (lambda k: noise_band.get(k))
    if "buffer" in gpu:
        assert_bits("buffer", gpu["buffer"], ref["buffer"])
    if "zcr" in gpu:
        assert np.array_equal(gpu["zcr"].astype(np.int64), ref["zcr"].astype(np.int64)), "zcr not bit-exact"
    if "complex_real" in gpu:
        rr, ri = ref["complexSpectrum"]["real"], ref["complexSpectrum"]["imag"]
        if exact:
            assert_bits("complexSpectrum.real", gpu["complex_real"], rr)
            assert_bits("complexSpectrum.imag", gpu["complex_imag"], ri)
        else:
            pk = np.sqrt(rr.astype(np.float64) ** 2 + ri.astype(np.float64) ** 2).max(axis=1) if len(rr) else None
            assert_spectrum("complexSpectrum.real", gpu["complex_real"], rr, peak=pk)
            assert_spectrum("complexSpectrum.imag", gpu["complex_imag"], ri, peak=pk)
    if "amplitude_spectrum" in gpu:
        if exact:
            assert_bits("amplitudeSpectrum", gpu["amplitude_spectrum"], ref["amplitudeSpectrum"])
        else:
            assert_spectrum("amplitudeSpectrum", gpu["amplitude_spectrum"], ref["amplitudeSpectrum"])
    if "power_spectrum" in gpu:
        if exact:
            assert_bits("powerSpectrum", gpu["power_spectrum"], ref["powerSpectrum"])
        else:
            assert_spectrum("powerSpectrum", gpu["power_spectrum"], ref["powerSpectrum"], tol=2.5 * SPECTRA_TOL)
    if "spectral_rolloff" in gpu:
        bin_hz = sr / (2 * (n - 1))
        gb = np.rint(gpu["spectral_rolloff"].astype(np.float64) / bin_hz)
        rb = np.rint(ref["spectralRolloff"] / bin_hz)
        assert np.array_equal(gb, rb), "rolloff bin mismatch at %s" % np.argwhere(gb != rb)[:5].tolist()
        assert_numbers("spectralRolloff", gpu["spectral_rolloff"], ref["spectralRolloff"], tol=1e-6, abs_tol=0.0)
    for field, feat in NUMBER_FIELDS.items():
        if field in gpu:
            # The flat "1e-3 relative OR absolute" rule is vacuous where a feature's magnitude is far below 1e-3
            # (a wrong formula would pass): rms and energy are plain sums and are held to the relative bound alone,
            # slope (~1e-7) to 1e-13, flatness (1e-5 .. 1 on tonal .. noisy frames) to a floor scaled to the batch.
            abs_tol = None
            if feat == "spectralSlope":
                abs_tol = SLOPE_ABS_TOL
            elif feat in ("rms", "energy"):
                abs_tol = 1.2e-38  # (outputs are float32: a reference value below its normal range cannot be held)
            elif feat == "spectralFlatness":
                fin = np.isfinite(np.asarray(ref[feat], dtype=np.float64))
                med = float(np.median(np.abs(np.asarray(ref[feat], dtype=np.float64)[fin]))) if fin.any() else 0.0
                abs_tol = min(tol, max(1e-7, tol * med))
            out[feat] = assert_numbers(feat, gpu[field], ref[feat], tol=tol, abs_tol=abs_tol, band=nb(feat))
    if "loudness_specific" in gpu:
        nl = nb("loudness")
        out["loudness.specific"] = assert_numbers("loudness.specific", gpu["loudness_specific"],
                                                  ref["loudness"]["specific"], tol=tol,
                                                  band=None if nl is None else nl["specific"])
        out["loudness.total"] = assert_numbers("loudness.total", gpu["loudness_total"], ref["loudness"]["total"],
                                               tol=tol, band=None if nl is None else nl["total"])
    if "mfcc" in gpu:
        out["mfcc"] = assert_numbers("mfcc", gpu["mfcc"], ref["mfcc"], tol=tol, band=nb("mfcc"))
    return {k: v for k, v in out.items() if v}
