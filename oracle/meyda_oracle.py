"""meyda_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatement (vectorised over frames) of Meyda's per-frame feature path
(reference: kirbysayshi/meyda v1.1.0, JavaScript, /root/reference).  Written
independently of oracle/meyda_oracle.c; tests require the two to agree bit for
bit.  JS semantics: float64 arithmetic, float32 rounding exactly where the
reference stores into a Float32Array, no FMA, sequential accumulation order
(np.cumsum is a running sum, not pairwise).

PINNING: the reference has no tests/golden vectors and no JS engine is available
in this image; both restatements are pinned against the reference's own .js
files executed by oracle/minijs.py (tests/test_js_pin.py, DESIGN.md "Oracle").

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this module; meyda_b200/ never does.
"""
from __future__ import annotations

import struct

import numpy as np

NUM_BARK_BANDS = 24  # src/meyda.js:214
NUM_MEL_FILTERS = 26  # src/extractors/mfcc.js:15
NUM_MFCC = 13  # src/extractors/mfcc.js:71

f32 = np.float32
f64 = np.float64

# src/feature-info.js:3-64 (lower-case types as in the code)
FEATURE_INFO = {
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
ALL_FEATURES = list(FEATURE_INFO)


def is_power_of_two(num) -> bool:
    """src/utils.js:13-19."""
    num = float(num)
    while num % 2 == 0 and num > 1:
        num /= 2
    return num == 1


def read_wav_pcm16(path: str):
    """Mono 16-bit PCM RIFF reader; int16/32768 -> float32 (decodeAudioData
    convention; replaces lib/bufferLoader.js:13-44).  Returns (float32[], sr)."""
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, sz = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[pos + 8:pos + 24])
        elif cid == b"data":
            data = b[pos + 8:pos + 8 + sz]
        pos += 8 + sz + (sz & 1)
    assert fmt is not None and fmt[0] == 1 and fmt[1] == 1 and fmt[5] == 16
    pcm = np.frombuffer(data[: len(data) // 2 * 2], dtype="<i2")
    return pcm16_to_float(pcm), float(fmt[2])


def degenerate_frame(name: str, N: int) -> np.ndarray:
    """Named synthetic frames shared by the JS-vector generator and the tests."""
    t = np.arange(N)
    if name == "silence":
        x = np.zeros(N)
    elif name == "impulse":
        x = np.eye(1, N, 7)[0]
    elif name == "square":
        x = np.where((t // 16) % 2 == 0, 1.0, -1.0)
    elif name == "dc":
        x = np.full(N, 0.5)
    elif name == "tone":
        x = 0.8 * np.sin(2 * np.pi * 32 * t / N)
    elif name == "tiny":
        x = np.random.default_rng(5).standard_normal(N) * 1e-30
    elif name == "negzero":
        x = np.where(t % 2 == 0, -0.0, 0.0)
    elif name == "nan":
        x = synth_clip(2, N).astype(np.float64)
        x[N // 3] = np.nan
    else:
        raise KeyError(name)
    return x.astype(f32)


def pcm16_to_float(pcm: np.ndarray) -> np.ndarray:
    return (pcm.astype(f32) / f32(32768.0)).astype(f32)


def hanning(N: int) -> np.ndarray:
    """src/meyda.js:128-138 (symmetric, N-1 denominator)."""
    i = np.arange(N, dtype=f64)
    return (0.5 - 0.5 * np.cos(2 * np.pi * i / (N - 1))).astype(f32)


def hamming(N: int) -> np.ndarray:
    """src/meyda.js:116-126 (cos(2*PI*(i/N - 1)))."""
    i = np.arange(N, dtype=f64)
    return (0.54 - 0.46 * np.cos(2 * np.pi * (i / N - 1))).astype(f32)


def blackman(N: int) -> np.ndarray:
    """src/meyda.js:140-156: commented out in the reference ("UNFINISHED"); the formula it states, for every i."""
    i = np.arange(N, dtype=f64)
    return (0.42 - 0.5 * np.cos(2 * np.pi * i / (N - 1)) + 0.08 * np.cos(4 * np.pi * i / (N - 1))).astype(f32)


def window_table(N: int, name: str) -> np.ndarray:
    if name == "hanning":
        return hanning(N)
    if name == "hamming":
        return hamming(N)
    if name == "blackman":
        return blackman(N)
    raise ValueError("unknown windowingFunction %r" % (name,))


def bark_scale(N: int, sr: float) -> np.ndarray:
    """src/meyda.js:170-182; the Hz value takes a float32 round trip."""
    i = np.arange(N, dtype=f64)
    hz = (i * sr / N).astype(f32).astype(f64)
    return (13 * np.arctan(hz / 1315.8) + 3.5 * np.arctan((hz / 7518) ** 2)).astype(f32)


def bark_band_limits(bark: np.ndarray, n_spec: int, nb: int = NUM_BARK_BANDS) -> np.ndarray:
    """src/extractors/loudness.js:24-45."""
    last = float(bark[n_spec - 1])
    end = last / nb
    band = 1
    bb = np.zeros(nb + 1, dtype=np.int32)
    for i in range(n_spec):
        while float(bark[i]) > end:
            if band <= nb:
                bb[band] = i
            band += 1
            end = band * last / nb
    bb[nb] = n_spec - 1
    return bb


def mel_bins(N: int, sr: float, num_filters: int = NUM_MEL_FILTERS) -> np.ndarray:
    """src/extractors/mfcc.js:7-38: 28 FFT-bin edges (floats holding ints)."""
    lower = 1125 * np.log(1 + 0.0 / 700)
    upper = 1125 * np.log(1 + (sr / 2) / 700)
    value_to_add = (upper - lower) / (num_filters + 1)
    mel = (np.arange(num_filters + 2, dtype=f64) * value_to_add).astype(f32)
    hz = (700 * (np.exp(mel.astype(f64) / 1125) - 1)).astype(f32)
    return np.floor((N + 1) * hz.astype(f64) / sr)


def mel_filterbank(N: int, sr: float, num_filters: int = NUM_MEL_FILTERS) -> np.ndarray:
    """src/extractors/mfcc.js:40-51: float64 [26][N/2+1]."""
    bins = mel_bins(N, sr, num_filters)
    fb = np.zeros((num_filters, N // 2 + 1), dtype=f64)
    for j in range(num_filters):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def dct_matrix(num_filters: int = NUM_MEL_FILTERS, num_coeffs: int = NUM_MFCC) -> np.ndarray:
    """src/extractors/mfcc.js:67-83: float32 [i + j*13]."""
    k = np.pi / num_filters
    w1 = 1.0 / np.sqrt(f64(num_filters))
    w2 = np.sqrt(2.0 / num_filters)
    d = np.zeros(num_coeffs * num_filters, dtype=f32)
    for i in range(num_coeffs):
        for j in range(num_filters):
            d[i + j * num_coeffs] = f32((w1 if i == 0 else w2) * np.cos(k * (i + 1) * (j + 0.5)))
    return d


def _bit_reverse_perm(n: int) -> np.ndarray:
    """lib/jsfft/fft.js:173-208: out[i] = in[rev(i)] (pair swaps)."""
    bits = n.bit_length() - 1
    idx = np.arange(n)
    rev = np.zeros(n, dtype=np.int64)
    for b in range(bits):
        rev |= ((idx >> b) & 1) << (bits - 1 - b)
    return rev


def fft_jsfft(real: np.ndarray, imag: np.ndarray | None = None):
    """lib/jsfft/fft.js:123-171 on a [frames, n] float32 batch (forward)."""
    real = np.atleast_2d(np.asarray(real, dtype=f32))
    F, n = real.shape
    assert n & (n - 1) == 0
    perm = _bit_reverse_perm(n)
    o_r = real[:, perm].copy()
    o_i = np.zeros_like(o_r) if imag is None else np.atleast_2d(np.asarray(imag, dtype=f32))[:, perm].copy()
    sqrt1_2 = f64(np.sqrt(0.5))  # Math.SQRT1_2
    width = 1
    while width < n:
        del_r, del_i = np.cos(np.pi / width), np.sin(np.pi / width)
        fr = np.empty(width, dtype=f64)
        fi = np.empty(width, dtype=f64)
        a, b = 1.0, 0.0
        for j in range(width):  # twiddle recurrence, fft.js:162-164
            fr[j], fi[j] = a, b
            a, b = a * del_r - b * del_i, a * del_i + b * del_r
        v_r = o_r.reshape(F, n // (2 * width), 2, width)
        v_i = o_i.reshape(F, n // (2 * width), 2, width)
        l_r, l_i = v_r[:, :, 0, :].astype(f64), v_i[:, :, 0, :].astype(f64)
        x_r, x_i = v_r[:, :, 1, :].astype(f64), v_i[:, :, 1, :].astype(f64)
        r_r = fr * x_r - fi * x_i
        r_i = fi * x_r + fr * x_i
        v_r[:, :, 0, :] = (sqrt1_2 * (l_r + r_r)).astype(f32)
        v_i[:, :, 0, :] = (sqrt1_2 * (l_i + r_i)).astype(f32)
        v_r[:, :, 1, :] = (sqrt1_2 * (l_r - r_r)).astype(f32)
        v_i[:, :, 1, :] = (sqrt1_2 * (l_i - r_i)).astype(f32)
        width <<= 1
    return o_r, o_i


def num_frames(length: int, N: int, hop: int) -> int:
    return 0 if length < N else (length - N) // hop + 1


def frame_signal(signal: np.ndarray, N: int, hop: int) -> np.ndarray:
    signal = np.ascontiguousarray(signal, dtype=f32)
    nf = num_frames(len(signal), N, hop)
    if nf == 0:
        return np.zeros((0, N), dtype=f32)
    return np.lib.stride_tricks.as_strided(signal, (nf, N), (hop * 4, 4), writeable=False)


def _seqsum(x: np.ndarray) -> np.ndarray:
    """Sequential (ascending) float64 accumulation along axis 1."""
    x = np.asarray(x, dtype=f64)
    if x.shape[1] == 0:
        return np.zeros(x.shape[0], dtype=f64)
    return np.cumsum(x, axis=1)[:, -1]


def _mu(i: int, amp64: np.ndarray) -> np.ndarray:
    """src/utils.js:1-11."""
    k = np.arange(amp64.shape[1], dtype=f64) ** i
    with np.errstate(all="ignore"):
        return _seqsum(k * np.abs(amp64)) / _seqsum(amp64)


def exact_spectrum(windowed: np.ndarray) -> np.ndarray:
    """The transform jsfft approximates, in float64: conj(DFT)/sqrt(N)."""
    w = np.atleast_2d(np.asarray(windowed)).astype(f64)
    return np.conj(np.fft.fft(w, axis=1)) / np.sqrt(f64(w.shape[1]))


NOISE_FEATURES = ["spectralCentroid", "spectralFlatness", "spectralSlope", "spectralSpread", "spectralSkewness",
                  "spectralKurtosis", "loudness", "perceptualSpread", "perceptualSharpness", "mfcc"]


def noise_band(signal: np.ndarray, bufferSize: int, hop: int | None = None, sr: float = 44100.0,
               window: str = "hanning", draws: int = 8, seed: int = 20261018, chunk: int = 2048,
               level: float = 2.0) -> dict:
    """How far each spectral feature moves under spectral noise of the size of
    the REFERENCE's own FFT rounding noise (test infrastructure for the
    float32-FFT kernels, see tests/parity.py).

    Per frame: sigma = level * rms_k |Z_jsfft[k] - Z_exact[k]| (the reference's
    actual noise level on that frame; level 2 because the float32 kernels measure
    1.6x the reference's noise (tools/fft_noise.py) and gpu - ref carries both).  The band of a feature is the largest of
    |feat(jsfft) - feat(exact)| and |feat(exact + white noise of that sigma) -
    feat(exact)| over `draws` draws.  Returns {feature: band array}; loudness ->
    {'specific','total'}.
    """
    hop = bufferSize if hop is None else hop
    frames_all = frame_signal(signal, bufferSize, hop)
    rng = np.random.default_rng(seed)
    parts = []
    for c0 in range(0, max(len(frames_all), 1), chunk):
        frames = frames_all[c0:c0 + chunk]
        N = bufferSize
        win = window_table(N, window)
        windowed = (frames.astype(f64) * win.astype(f64)).astype(f32)
        zx = exact_spectrum(windowed) if len(frames) else np.zeros((0, N), np.complex128)
        rr, ri = fft_jsfft(windowed) if len(frames) else (np.zeros((0, N), f32),) * 2
        with np.errstate(all="ignore"):
            sigma = level * np.sqrt(np.mean(np.abs((rr.astype(f64) + 1j * ri.astype(f64)) - zx) ** 2, axis=1,
                                            keepdims=True))
            base = extract_frames(frames, sr, window, NOISE_FEATURES, (zx.real, zx.imag))
            ref = extract_frames(frames, sr, window, NOISE_FEATURES, (rr, ri))
            band = _absdiff(ref, base)
            for _ in range(draws):
                nz = (rng.standard_normal(zx.shape) + 1j * rng.standard_normal(zx.shape)) * (sigma / np.sqrt(2.0))
                zi = zx + nz
                band = _maxdict(band, _absdiff(extract_frames(frames, sr, window, NOISE_FEATURES, (zi.real, zi.imag)), base))
        parts.append(band)
    return _concat(parts)


def _absdiff(a: dict, b: dict) -> dict:
    with np.errstate(all="ignore"):
        return {k: ({s: np.nan_to_num(np.abs(a[k][s].astype(f64) - b[k][s].astype(f64)), nan=0.0, posinf=0.0)
                     for s in a[k]} if isinstance(a[k], dict)
                    else np.nan_to_num(np.abs(a[k].astype(f64) - b[k].astype(f64)), nan=0.0, posinf=0.0)) for k in a}


def _maxdict(a: dict, b: dict) -> dict:
    return {k: ({s: np.maximum(a[k][s], b[k][s]) for s in a[k]} if isinstance(a[k], dict) else np.maximum(a[k], b[k]))
            for k in a}


def extract_frames(frames: np.ndarray, sr: float = 44100.0, window: str = "hanning",
                   features=None, fft="jsfft", params: dict | None = None) -> dict:
    """All requested features for a [F, N] float32 batch of raw frames.

    Number features -> float64[F]; arrays -> float32[F, len];
    complexSpectrum -> {'real','imag'}; loudness -> {'specific','total'}.

    fft="jsfft" is the reference's arithmetic.  fft="float64" swaps in a
    mathematically exact transform (numpy.fft in float64, same sign and 1/sqrt(N)
    scaling, rounded to float32 once): NOT the reference -- it measures how far the
    reference's own per-stage float32 rounding moves each feature (tests/parity.py).

    params: numBarkBands (the NUM_BARK_BANDS option of loudness.js:14), numMelFilters / numMfccCoefficients
    (the local constants of mfcc.js:15,71) and rolloffFraction (spectralRolloff.js:9) when not 24 / 26 / 13 / 0.99.
    """
    params = params or {}
    nb = int(params.get("numBarkBands") or NUM_BARK_BANDS)
    nf = int(params.get("numMelFilters") or NUM_MEL_FILTERS)
    nc = int(params.get("numMfccCoefficients") or NUM_MFCC)
    rolloff_fraction = float(params.get("rolloffFraction") or 0.99)
    feats = ALL_FEATURES if features is None else ([features] if isinstance(features, str) else list(features))
    frames = np.atleast_2d(np.asarray(frames, dtype=f32))
    F, N = frames.shape
    if not is_power_of_two(N):
        raise ValueError("Buffer size is not a power of two: Meyda will not run.")  # src/meyda.js:20-22
    n = N // 2
    out: dict = {}
    sig64 = frames.astype(f64)
    with np.errstate(all="ignore"):
        win = window_table(N, window)
        windowed = (sig64 * win.astype(f64)).astype(f32)  # src/meyda.js:158-168
        if isinstance(fft, tuple):  # an explicit (real, imag) spectrum, see noise_band()
            re, im = (np.asarray(a, dtype=f32) for a in fft)
        elif fft == "jsfft":
            re, im = fft_jsfft(windowed)
        elif fft == "float64":
            z = exact_spectrum(windowed)
            re, im = z.real.astype(f32), z.imag.astype(f32)
        else:
            raise ValueError(fft)
        amp = np.sqrt(re[:, :n].astype(f64) ** 2 + im[:, :n].astype(f64) ** 2).astype(f32)  # src/meyda.js:104-114
        amp64 = amp.astype(f64)
        power = (amp64 * amp64).astype(f32)  # powerSpectrum.js:1-7

        need = set(feats)
        if "buffer" in need:
            out["buffer"] = frames.copy()
        if "complexSpectrum" in need:
            out["complexSpectrum"] = {"real": re, "imag": im}
        if "amplitudeSpectrum" in need:
            out["amplitudeSpectrum"] = amp
        if "powerSpectrum" in need:
            out["powerSpectrum"] = power
        if "rms" in need:  # rms.js:1-11
            out["rms"] = np.sqrt(_seqsum(sig64 * sig64) / N)
        if "energy" in need:  # energy.js:1-7
            out["energy"] = _seqsum(np.abs(sig64) ** 2)
        if "zcr" in need:  # zcr.js:1-9
            a, b = frames[:, :-1], frames[:, 1:]
            out["zcr"] = (((a >= 0) & (b < 0)) | ((a < 0) & (b >= 0))).sum(axis=1).astype(f64)
        if need & {"spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"}:
            m1, m2, m3, m4 = (_mu(i, amp64) for i in (1, 2, 3, 4))
            if "spectralCentroid" in need:
                out["spectralCentroid"] = m1
            if "spectralSpread" in need:
                out["spectralSpread"] = np.sqrt(m2 - m1 ** 2)
            if "spectralSkewness" in need:
                out["spectralSkewness"] = (2 * m1 ** 3 - 3 * m1 * m2 + m3) / np.sqrt(m2 - m1 ** 2) ** 3
            if "spectralKurtosis" in need:
                out["spectralKurtosis"] = (-3 * m1 ** 4 + 6 * m1 * m2 - 4 * m1 * m3 + m4) / np.sqrt(m2 - m1 ** 2) ** 4
        if "spectralFlatness" in need:  # spectralFlatness.js:1-10
            out["spectralFlatness"] = np.exp(_seqsum(np.log(amp64)) / n) * n / _seqsum(amp64)
        if "spectralSlope" in need:  # spectralSlope.js:1-18
            freq = np.arange(n, dtype=f64) * sr / N
            amp_sum = _seqsum(amp64)
            freq_sum = np.cumsum(freq)[-1] if n else 0.0
            pow_freq_sum = np.cumsum(freq * freq)[-1] if n else 0.0
            amp_freq_sum = _seqsum(freq * amp64)
            out["spectralSlope"] = (n * amp_freq_sum - freq_sum * amp_sum) / (amp_sum * (pow_freq_sum - freq_sum ** 2))
        if "spectralRolloff" in need:  # spectralRolloff.js:1-16
            out["spectralRolloff"] = _rolloff(amp64, sr, rolloff_fraction)
        if need & {"loudness", "perceptualSpread", "perceptualSharpness"}:
            bb = bark_band_limits(bark_scale(N, sr), n, nb)
            specific = np.zeros((F, nb), dtype=f32)
            for b in range(nb):  # loudness.js:47-66
                specific[:, b] = np.power(_seqsum(amp64[:, bb[b]:bb[b + 1]]), 0.23).astype(f32)
            total = _seqsum(specific)
            if "loudness" in need:
                out["loudness"] = {"specific": specific, "total": total}
            if "perceptualSpread" in need:  # perceptualSpread.js:1-14
                mx = np.fmax.reduce(specific.astype(f64), axis=1, initial=0.0)  # NaN never compares greater
                out["perceptualSpread"] = ((total - mx) / total) ** 2
            if "perceptualSharpness" in need:  # perceptualSharpness.js:1-16
                acc = np.zeros(F, dtype=f64)
                for i in range(nb):
                    if i < 15:  # spec[i + 1] past the end is `undefined`: NaN (fewer than 16 bands)
                        acc = acc + (i + 1) * (specific[:, i + 1].astype(f64) if i + 1 < nb else np.nan)
                    else:
                        acc = acc + 0.066 * np.exp(0.171 * (i + 1))
                out["perceptualSharpness"] = acc * (0.11 / total)
        if "mfcc" in need:
            out["mfcc"] = _mfcc(power, N, sr, nf, nc)
    return {k: out[k] for k in feats if k in out}


def _rolloff(amp64: np.ndarray, sr: float, fraction: float = 0.99) -> np.ndarray:
    F, n = amp64.shape
    nyq_bin = sr / (2 * (n - 1))
    ec = _seqsum(amp64)
    thr = fraction * ec
    # ec after subtracting amp[n-1], ..., amp[m] in that order (sequential float64)
    seq = np.cumsum(np.concatenate([ec[:, None], -amp64[:, ::-1]], axis=1), axis=1)  # [F, n+1]
    # seq[:, t] = ec after t subtractions; loop stops at first t with !(seq > thr) or t == n
    cont = seq > thr[:, None]
    t = np.where(cont.all(axis=1), n, np.argmin(cont, axis=1))
    return (n - t).astype(f64) * nyq_bin  # (q + 1) with q = n - 1 - t


def _mfcc(power: np.ndarray, N: int, sr: float, num_filters: int = NUM_MEL_FILTERS,
          num_coeffs: int = NUM_MFCC) -> np.ndarray:
    """src/extractors/mfcc.js:53-93 with the float32 running sum."""
    F, n = power.shape
    fb = mel_filterbank(N, sr, num_filters)
    p64 = power.astype(f64)
    logged = np.zeros((F, num_filters), dtype=f32)
    nonfinite = ~np.isfinite(p64).all(axis=1)
    for i in range(num_filters):
        acc = np.zeros(F, dtype=f32)
        nz = np.nonzero(fb[i, :n])[0]
        if len(nz):
            for j in range(nz[0], nz[-1] + 1):  # zero-weight terms add +0 and are skipped
                acc = (acc.astype(f64) + fb[i, j] * p64[:, j]).astype(f32)
        acc = np.where(nonfinite, f32(np.nan), acc)  # 0 * inf / NaN anywhere poisons every band
        logged[:, i] = np.log(acc.astype(f64)).astype(f32)
    dct = dct_matrix(num_filters, num_coeffs)
    out = np.zeros((F, num_coeffs), dtype=f32)
    for k in range(num_coeffs):
        v = np.zeros(F, dtype=f64)
        for m in range(num_filters):
            v = v + f64(dct[k + m * num_coeffs]) * logged[:, m].astype(f64)
        out[:, k] = (v / num_coeffs).astype(f32)
    return out


def extract(signal: np.ndarray, bufferSize: int, hop: int | None = None, sr: float = 44100.0,
            window: str = "hanning", features=None, chunk: int = 4096, fft: str = "jsfft",
            params: dict | None = None) -> dict:
    """Frame a clip ([f*hop, f*hop+N), no padding) and extract, in chunks."""
    hop = bufferSize if hop is None else hop
    frames = frame_signal(signal, bufferSize, hop)
    parts = [extract_frames(frames[i:i + chunk], sr, window, features, fft, params)
             for i in range(0, len(frames), chunk)]
    if not parts:
        parts = [extract_frames(np.zeros((0, bufferSize), dtype=f32), sr, window, features, fft, params)]
    return _concat(parts)


def _concat(parts):
    out = {}
    for k in parts[0]:
        if isinstance(parts[0][k], dict):
            out[k] = {s: np.concatenate([p[k][s] for p in parts]) for s in parts[0][k]}
        else:
            out[k] = np.concatenate([p[k] for p in parts])
    return out


def synth_clip(clip_index: int, length: int, sr: float = 44100.0, seed: int = 0x4D455944) -> np.ndarray:
    """Deterministic synthetic clip for parity subsets: white noise (amp 0.25)
    plus three sines (amp 0.2 each) at log-uniform 55..15000 Hz.  |x| < 0.85."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[clip_index, 0, 0, 0]))
    fp = np.exp(rng.uniform(np.log(55.0), np.log(15000.0), 3))
    ph = rng.uniform(0, 2 * np.pi, 3)
    t = np.arange(length, dtype=f64) / sr
    x = 0.25 * (rng.random(length) - 0.5) * 2
    for f, p in zip(fp, ph):
        x = x + 0.2 * np.sin(2 * np.pi * f * t + p)
    return x.astype(f32)
