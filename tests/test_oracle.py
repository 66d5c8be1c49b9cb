"""CPU checks of the oracle (oracle/): the C and numpy restatements against each
other, against numpy.fft / closed forms, against the survey's spot values
 and against the committed golden fixtures.  The reference has
no tests of its own (package.json:25), so this is what pins the oracle."""
import numpy as np
import pytest

from oracle import c_oracle, meyda_oracle as mo

SR = 44100.0


def test_fft_matches_numpy_fft_conj_unitary():
    rng = np.random.default_rng(1)
    for n in (8, 64, 512, 2048):
        x = rng.standard_normal(n).astype(np.float32)
        re, im = mo.fft_jsfft(x[None])
        want = np.conj(np.fft.fft(x.astype(np.float64))) / np.sqrt(n)  # +i sign, 1/sqrt(N)
        scale = np.abs(want).max()
        assert np.abs(re[0] - want.real).max() < 4e-7 * scale * np.log2(n)
        assert np.abs(im[0] - want.imag).max() < 4e-7 * scale * np.log2(n)
        cre, cim = c_oracle.fft(x, np.zeros(n, np.float32))
        assert np.array_equal(cre, re[0]) and np.array_equal(cim, im[0])


def test_fft_parseval_and_impulse():
    n = 1024
    x = np.zeros(n, np.float32)
    x[3] = 1.0
    re, im = mo.fft_jsfft(x[None])
    assert np.allclose(np.hypot(re, im), 1 / np.sqrt(n), rtol=1e-5)  # flat spectrum
    rng = np.random.default_rng(2)
    y = rng.standard_normal(n).astype(np.float32)
    re, im = mo.fft_jsfft(y[None])
    assert np.isclose((re.astype(np.float64) ** 2 + im.astype(np.float64) ** 2).sum(), (y.astype(np.float64) ** 2).sum(), rtol=1e-5)


def test_window_tables_closed_form():
    for N in (256, 2048):
        h = mo.hanning(N)
        assert h[0] == 0 and abs(h[N - 1]) < 1e-7 and np.allclose(h, h[::-1], atol=1e-7)  # symmetric, N-1
        m = mo.hamming(N)
        assert np.isclose(m[0], 0.08, atol=1e-7) and np.isclose(m[N // 2], 1.0, atol=1e-6)  # periodic
        t = c_oracle.plan_tables(N)
        assert np.array_equal(t["hanning"], h) and np.array_equal(t["hamming"], m)
        assert np.array_equal(t["bark"], mo.bark_scale(N, SR))
        assert np.array_equal(t["bbLimits"], mo.bark_band_limits(mo.bark_scale(N, SR), N // 2))


def test_bark_limits_and_mel_bins_survey_values():
    assert mo.bark_band_limits(mo.bark_scale(2048, SR), 1024).tolist() == [
        0, 5, 10, 15, 20, 26, 32, 38, 45, 53, 62, 72, 84, 98, 115, 137, 163, 195, 233, 278, 332, 399, 492, 648, 1023]
    assert mo.bark_band_limits(mo.bark_scale(512, SR), 256).tolist() == [
        0, 2, 3, 4, 5, 7, 8, 10, 12, 14, 16, 18, 21, 25, 29, 34, 41, 49, 59, 70, 83, 100, 123, 162, 255]
    assert mo.mel_bins(2048, SR).astype(int).tolist() == [
        0, 4, 9, 15, 21, 29, 37, 47, 58, 71, 85, 101, 120, 141, 165, 192, 223, 258, 298, 344, 396, 455, 522, 598,
        685, 784, 896, 1024]
    assert mo.mel_bins(512, SR).astype(int).tolist() == [
        0, 1, 2, 3, 5, 7, 9, 11, 14, 17, 21, 25, 30, 35, 41, 48, 55, 64, 74, 86, 99, 113, 130, 149, 171, 196, 224,
        256]
    assert mo.mel_bins(256, SR).astype(int).tolist()[:6] == [0, 0, 1, 1, 2, 3]


def test_survey_spot_values(golden_audio):
    """SURVEY.md 8(a) [probe] values: sound1.wav, N=512, hop=512, hanning."""
    r = c_oracle.extract(golden_audio["sound1"], 512, 512, SR)
    assert len(r["rms"]) == 325
    f = 0
    want = dict(rms=0.0050815644, energy=0.013221016, zcr=26, spectralCentroid=32.01215955,
                spectralSpread=49.93446702, spectralSkewness=2.12649896, spectralKurtosis=3.63576656,
                spectralFlatness=0.26976081, spectralRolloff=18937.0588, perceptualSpread=0.88579710,
                perceptualSharpness=0.79931365)
    for k, v in want.items():
        assert np.isclose(r[k][f], v, rtol=2e-8 * 10, atol=0), (k, r[k][f], v)
    assert np.isclose(r["loudness"]["total"][f], 8.12518571, rtol=1e-8)
    assert np.allclose(r["mfcc"][f][:4], [0.56195509, 0.11801300, 0.12774231, 0.06569117], rtol=2e-7)
    assert np.allclose(r["amplitudeSpectrum"][f][:3], [0.017596455, 0.022795303, 0.029011661], rtol=2e-7)
    assert np.isclose(r["spectralSlope"][f], 2.68e-07, rtol=2e-3)
    assert r["zcr"][100] == 16 and np.isclose(r["spectralCentroid"][100], 14.73148645, rtol=1e-9)
    assert np.isclose(r["spectralKurtosis"][100], 18.64583624, rtol=1e-9)
    assert np.isclose(r["spectralRolloff"][100], 7782.35294, rtol=1e-8)


def _same(a, b, rtol=0.0):
    if isinstance(a, dict):
        return all(_same(a[k], b[k], rtol) for k in a)
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype == np.float32:
        return np.array_equal(a.view(np.uint32), b.view(np.uint32)) or np.array_equal(a, b, equal_nan=True)
    return np.allclose(a, b, rtol=rtol, atol=1e-11 if rtol else 0, equal_nan=True)


@pytest.mark.parametrize("N,hop", [(256, 256), (512, 512), (1024, 300), (2048, 512)])
def test_numpy_and_c_restatements_agree(golden_audio, N, hop):
    """Two independently written restatements: float32 arrays bit for bit,
    float64 numbers to 1e-7 (Math.pow vs repeated multiply, under cancellation)."""
    for name in ("sound1", "sound3"):
        x = golden_audio[name][: 40 * N]
        a = mo.extract(x, N, hop, SR)
        b = c_oracle.extract(x, N, hop, SR)
        for k in a:
            assert _same(a[k], b[k], rtol=1e-7), (name, N, k)


def test_hamming_path_agrees(golden_audio):
    x = golden_audio["sound2"][:8192]
    a = mo.extract(x, 1024, 512, SR, window="hamming")
    b = c_oracle.extract(x, 1024, 512, SR, window="hamming")
    for k in a:
        assert _same(a[k], b[k], rtol=1e-7), k


def test_blackman_window_closed_form_and_path(golden_audio):
    """The window the reference leaves commented out (src/meyda.js:140-156): its stated formula.  Symmetric,
    zero-ish at the ends, 1 in the middle of an odd length; both oracles agree on the whole path."""
    for N in (16, 512, 2048):
        w = mo.blackman(N).astype(np.float64)
        assert np.allclose(w, w[::-1], atol=1e-7) and abs(w[0]) < 1e-7 and w.max() <= 1.0
        assert np.allclose(w, np.blackman(N), atol=2e-7)  # numpy's is the same symmetric window
    x = golden_audio["sound1"][:8192]
    a = mo.extract(x, 512, 256, SR, window="blackman")
    b = c_oracle.extract(x, 512, 256, SR, window="blackman")
    for k in a:
        assert _same(a[k], b[k], rtol=1e-7), k


def test_golden_fixture_matches_oracle(golden_audio, golden_features):
    """Guards the oracle against drift; also exercises the numpy path on full clips."""
    names = [str(s) for s in golden_features["scalar_names"]]
    for clip in ("sound1", "sound2", "sound3"):
        for N in (256, 512, 1024, 2048):
            key = f"{clip}/{N}"
            r = mo.extract(golden_audio[clip], N, N, SR)
            g = golden_features[f"{key}/scalars"]
            for i, nm in enumerate(names):
                got = r["loudness"]["total"] if nm == "loudness.total" else r[nm]
                assert np.allclose(got, g[:, i], rtol=1e-7, atol=1e-11, equal_nan=True), (key, nm)
            assert _same(r["mfcc"], golden_features[f"{key}/mfcc"])
            assert _same(r["loudness"]["specific"], golden_features[f"{key}/specific"])
            pick = golden_features[f"{key}/frames"]
            assert _same(r["amplitudeSpectrum"][pick], golden_features[f"{key}/amp"])
            assert _same(r["complexSpectrum"]["imag"][pick], golden_features[f"{key}/imag"])


def test_frame_counts_survey():
    want = {"sound1": (166400, [650, 325, 162, 81]), "sound2": (441001, [1722, 861, 430, 215]),
            "sound3": (562688, [2198, 1099, 549, 274])}
    for _, (length, counts) in want.items():
        assert [mo.num_frames(length, N, N) for N in (256, 512, 1024, 2048)] == counts


def test_degenerate_frames():
    """SURVEY.md section 9 degenerate-frame contract."""
    N = 512
    z = np.zeros(N, np.float32)
    for impl in (mo.extract, c_oracle.extract):
        r = impl(z, N, N, SR)
        assert r["rms"][0] == 0 and r["energy"][0] == 0 and r["zcr"][0] == 0
        for k in ("spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis", "spectralFlatness",
                  "spectralSlope", "perceptualSpread"):
            assert np.isnan(r[k][0]), k
        assert np.isclose(r["spectralRolloff"][0], 256 * SR / (2 * 255))
        assert np.isposinf(r["perceptualSharpness"][0])
        assert np.isnan(r["mfcc"][0]).all()
        assert (r["loudness"]["specific"][0] == 0).all() and r["loudness"]["total"][0] == 0
    # N = 256: filter 1 has zero weight => every mfcc is +-inf, pattern [-inf x8, +inf x5]
    x = mo.synth_clip(0, 256)
    r = c_oracle.extract(x, 256, 256, SR)
    assert np.isneginf(r["mfcc"][0][:8]).all() and np.isposinf(r["mfcc"][0][8:]).all()
    # NaN sample: zcr ignores it, everything spectral is NaN
    x = mo.synth_clip(1, N).copy()
    x[7] = np.nan
    r = c_oracle.extract(x, N, N, SR)
    assert np.isnan(r["spectralCentroid"][0]) and np.isfinite(r["zcr"][0])


def test_is_power_of_two_and_errors():
    assert mo.is_power_of_two(1) and mo.is_power_of_two(4096) and not mo.is_power_of_two(0)
    assert not mo.is_power_of_two(768)
    with pytest.raises(ValueError, match="Buffer size is not a power of two"):
        mo.extract(np.zeros(600, np.float32), 600)
    with pytest.raises(ValueError, match="Buffer size is not a power of two"):
        c_oracle.extract(np.zeros(600, np.float32), 600)
