"""Parity of the CUDA path (through the C ABI) against the oracle.  These are
the parity tests proper; they need a B200 (`-m gpu`)."""
import os

import numpy as np
import pytest

import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import c_oracle, meyda_oracle as mo
from tests import parity

pytestmark = pytest.mark.gpu
SR = 44100.0
EXACT = _capi.MB_FLAG_EXACT_FFT
FLAG_VARIANTS = [pytest.param(0, id="fast"), pytest.param(_capi.MB_FLAG_GENERIC_KERNEL, id="generic"),
                 pytest.param(_capi.MB_FLAG_NO_REFINE, id="fast-norefine"),  # (float32 FFT alone: may use the noise band)
                 pytest.param(EXACT, id="exact"),  # (the warp-per-frame exact kernel at 512 / 1024 / 2048)
                 pytest.param(EXACT | _capi.MB_FLAG_GENERIC_KERNEL, id="exact-generic"),  # (block per frame)
                 pytest.param(EXACT | _capi.MB_FLAG_CLUSTER_FFT, id="exact-cluster")]


def run_gpu(clips, N, hop=None, window="hanning", features=mb.FEATURES, flags=0):
    data, off, ln = mb.meyda._normalize_clips(clips)
    plan = mb.Plan(N, hop, SR, window, features, flags=flags)
    try:
        out, per = plan.extract_host(data, off, ln)
    finally:
        plan.close()
    return out, per


def oracle_concat(clips, N, hop, window="hanning", impl=c_oracle, **kw):
    parts = [impl.extract(c, N, hop, SR, window, **kw) for c in clips]
    return mo._concat(parts)


NO_REFINE = _capi.MB_FLAG_NO_REFINE


def is_adaptive(flags, N=None):
    """Every float32 plan (warp, multi-frame warp, generic and multi-warp-per-frame kernels alike) flags the frames
    whose features sit in the reference's own rounding noise and redoes them with the exact FFT (mb_adaptive.cuh):
    it must meet the flat 1e-3 tolerance with NO noise band.  Only MB_FLAG_NO_REFINE plans may use the band."""
    return not (flags & (EXACT | NO_REFINE))


def verify(out, clips, N, hop, window="hanning", flags=0, max_banded_frac=0.10, adaptive=None):
    """CUDA result vs the oracle.  The default (adaptive) plan must meet BASELINE.json's flat tolerances; the
    non-adaptive float32 kernels (MB_FLAG_GENERIC_KERNEL, MB_FLAG_NO_REFINE, sizes without a warp kernel) may use
    the reference's own rounding-noise band (tests/parity.py) on a bounded fraction of values, which is printed;
    the exact-FFT mode gets bit-exact spectra and 5e-6 on numbers."""
    if isinstance(clips, np.ndarray) and clips.ndim == 1:
        clips = [clips]
    clips = [c for c in clips if len(c) >= N]
    ref = oracle_concat(clips, N, hop, window)
    exact = bool(flags & EXACT)
    adaptive = is_adaptive(flags, N) if adaptive is None else adaptive
    noise = None if (exact or adaptive) else mo._concat([mo.noise_band(c, N, hop, SR, window) for c in clips])
    banded = parity.compare_all(out, ref, N, noise_band=noise, exact=exact)
    assert not (adaptive and banded), banded
    frames = max(1, len(ref["rms"]))
    if banded:
        print("noise-banded values (of %d frames): %s" % (frames, banded))
    for k, v in banded.items():
        per_frame = {"mfcc": 13, "loudness.specific": 24}.get(k, 1)
        assert v <= max_banded_frac * frames * per_frame, (k, v, frames)
    return banded


# ---- BASELINE config 5's shape: bufferSize 32768, hop 8192, amplitudeSpectrum + rolloff / flatness / slope over several
# channels of 319 frames, in the default (adaptive multi-warp-per-frame kernel) and the exact-cluster mode
@pytest.mark.parametrize("flags", [pytest.param(0, id="fast"), pytest.param(EXACT, id="exact-cluster")])
def test_config5_shape(flags):
    N, hop, frames = 32768, 8192, 319
    L = N + hop * (frames - 1)  # 2,637,824 samples: 60 s at 44.1 kHz holds 319 frames
    feats = ["amplitudeSpectrum", "spectralRolloff", "spectralFlatness", "spectralSlope"]
    clips = [mo.synth_clip(70 + c, L) for c in range(4)]
    t = np.arange(L) / SR
    clips[3] = (0.4 * np.sin(2 * np.pi * 997.0 * t)).astype(np.float32)  # one tonal channel: its frames are redone exactly
    data, off, ln = mb.meyda._normalize_clips(clips)
    plan = mb.Plan(N, hop, SR, "hanning", feats, flags=flags)
    try:
        out, per = plan.extract_host(data, off, ln)
        name, refined = plan.kernel_name, plan.refined_frames
    finally:
        plan.close()
    assert per.tolist() == [frames] * 4 and name == ("exact-cluster2" if flags else "big32768")
    if not flags:
        assert frames <= refined < 2 * frames, refined  # the tonal channel, and hardly anything of the noisy ones
    pick = np.r_[0:3, frames - 2:frames + 2, 2 * frames - 1, 3 * frames:3 * frames + 3, 4 * frames - 1]  # a few frames of every channel
    for g in pick:
        c, f = divmod(int(g), frames)
        ref = c_oracle.extract(clips[c][f * hop:f * hop + N], N, N, SR)
        one = {k: v[g:g + 1] for k, v in out.items()}
        assert parity.compare_all(one, ref, N, exact=bool(flags & EXACT)) == {}, (c, f)


# ---- BASELINE config 1: sound1.wav, N=512, five features
@pytest.mark.parametrize("flags", FLAG_VARIANTS)
def test_config1_sound1_512(golden_audio, flags):
    feats = ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"]
    out, per = run_gpu(golden_audio["sound1"], 512, features=feats, flags=flags)
    assert per.tolist() == [325] and set(out) == {"rms", "energy", "zcr", "amplitude_spectrum", "spectral_centroid"}
    verify(out, golden_audio["sound1"], 512, 512, flags=flags)


# ---- BASELINE config 2: all 18 extractors x 3 clips x 4 buffer sizes
@pytest.mark.parametrize("flags", FLAG_VARIANTS)
@pytest.mark.parametrize("N", [256, 512, 1024, 2048])
@pytest.mark.parametrize("clip", ["sound1", "sound2", "sound3"])
def test_config2_all_features(golden_audio, clip, N, flags):
    x = golden_audio[clip]
    out, per = run_gpu(x, N, flags=flags)
    assert per[0] == mo.num_frames(len(x), N, N)
    # sound3 is a sine sweep and sound2 band-limited noise: where a band holds
    # nothing but FFT rounding noise the reference's own float32 per-stage
    # rounding sets its value; see verify().
    verify(out, x, N, N, flags=flags, max_banded_frac=0.25)


def test_golden_fixture(golden_audio, golden_features):
    """CUDA output against the committed golden vectors (not the live oracle)."""
    names = [str(s) for s in golden_features["scalar_names"]]
    out, _ = run_gpu(golden_audio["sound1"], 1024)
    g = golden_features["sound1/1024/scalars"]
    ref = {n: g[:, i] for i, n in enumerate(names)}
    assert np.array_equal(out["zcr"], ref["zcr"].astype(np.int32))
    parity.assert_numbers("rms", out["rms"], ref["rms"])
    parity.assert_numbers("centroid", out["spectral_centroid"], ref["spectralCentroid"])
    parity.assert_numbers("mfcc", out["mfcc"], golden_features["sound1/1024/mfcc"])
    parity.assert_numbers("specific", out["loudness_specific"], golden_features["sound1/1024/specific"])
    pick = golden_features["sound1/1024/frames"]
    parity.assert_spectrum("amp", out["amplitude_spectrum"][pick], golden_features["sound1/1024/amp"])


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
@pytest.mark.parametrize("N,hop", [(2048, 512), (1024, 100), (512, 1), (256, 700), (2048, 2047)])
def test_hop_reuse_equivalence(N, hop, flags):
    """hop != N: result equals per-frame extraction of explicit slices."""
    x = mo.synth_clip(3, N + hop * 37 + 11)
    out, per = run_gpu(x, N, hop, flags=flags)
    assert per[0] == mo.num_frames(len(x), N, hop) and per[0] >= 38
    verify(out, x, N, hop, flags=flags)


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
def test_hamming_window(golden_audio, flags):
    x = golden_audio["sound2"][:40000]
    out, _ = run_gpu(x, 1024, 512, window="hamming", flags=flags)
    verify(out, x, 1024, 512, window="hamming", flags=flags)


@pytest.mark.parametrize("N,hop,flags", [(2048, 512, 0), (512, 512, 0), (1024, 256, EXACT)])
def test_blackman_window(golden_audio, N, hop, flags):
    """src/meyda.js:140-156 (commented out in the reference)."""
    x = golden_audio["sound1"][:50000]
    out, _ = run_gpu(x, N, hop, window="blackman", flags=flags)
    verify(out, x, N, hop, window="blackman", flags=flags)
    plan = mb.Plan(N, hop, SR, "blackman")
    assert np.array_equal(plan.tables()["window"], mo.blackman(N))
    plan.close()


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
def test_ragged_and_empty_clips(flags):
    N, hop = 512, 128
    lens = [0, 511, 512, 513, 5000, 640, 12345, 1]
    clips = [mo.synth_clip(10 + i, L) for i, L in enumerate(lens)]
    out, per = run_gpu(clips, N, hop, flags=flags)
    assert per.tolist() == [mo.num_frames(L, N, hop) for L in lens]
    verify(out, clips, N, hop, flags=flags)
    # no clip at all / only too-short clips
    out, per = run_gpu([], N, hop, flags=flags)
    assert len(per) == 0 and out["rms"].shape == (0,)
    out, per = run_gpu([np.zeros(100, np.float32)], N, hop, flags=flags)
    assert per.tolist() == [0] and out["mfcc"].shape == (0, 13)


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
@pytest.mark.parametrize("N", [16, 64, 128, 4096, 8192, 16384, 32768])
def test_other_buffer_sizes(N, flags):
    if (flags & EXACT) and N > 16384:  # does not fit one CTA: the 2-CTA cluster kernel takes over
        plan = mb.Plan(N, N, SR, flags=flags)
        assert plan.kernel_name == "exact-cluster2"
        plan.close()
    if flags == 0:  # bufferSize / 2048 warps per frame from 4096 up
        plan = mb.Plan(N, N // 4, SR)
        assert plan.kernel_name == ("big%d" % N if N >= 4096 else "generic")
        plan.close()
    x = mo.synth_clip(N, N * 3 + 5)
    hop = N // 4
    out, per = run_gpu(x, N, hop, flags=flags)
    verify(out, x, N, hop, flags=flags)


@pytest.mark.parametrize("N", [4096, 8192, 16384])
def test_multi_warp_frame_kernels(N):
    """The bufferSize-32768 kernel compiled for 2 / 4 / 8 warps per frame: many frames (several CTAs per SM, the
    grid-stride loop, the prefetch of the next frame), ragged clips, feature subsets, and a misaligned clip that
    must fall back to the generic kernel with results inside the same tolerances."""
    lens = [N * 6 + 4 * 37, N - 4, N, 0, N * 9, N * 2 + 8]
    clips = [mo.synth_clip(700 + i, L) for i, L in enumerate(lens)]
    hop = N // 4
    out, per = run_gpu(clips, N, hop)
    assert per.tolist() == [mo.num_frames(L, N, hop) for L in lens]
    verify(out, clips, N, hop)
    sub, _ = run_gpu(clips, N, hop, features=["amplitudeSpectrum", "spectralRolloff", "spectralFlatness", "spectralSlope"])
    for k in sub:
        assert np.array_equal(sub[k], out[k], equal_nan=True), k
    # many frames through few CTAs: every frame equals its single-clip run bit for bit
    x = mo.synth_clip(901, N + hop * 700)
    big, per = run_gpu(x, N, hop, features=["rms", "spectralCentroid", "mfcc", "loudness"])
    assert per[0] == 701
    pick = [0, 1, 147, 148, 295, 296, 592, 699, 700]
    for f in pick:
        one, _ = run_gpu(x[f * hop:f * hop + N], N, hop, features=["rms", "spectralCentroid", "mfcc", "loudness"])
        for k in one:
            assert np.array_equal(one[k][0], big[k][f], equal_nan=True), (k, f)
    # frames off the 16-byte grid (odd clip lengths in front, hop not a multiple of 4) are read by the lanes inside
    # the same kernel: same bits as the aligned single-frame run; and streaming in odd blocks == the batch call
    odd = [mo.synth_clip(77, 3), mo.synth_clip(78, N * 2 + 13), x[:N + 5 * (hop + 2)]]
    feats = ["rms", "zcr", "spectralCentroid", "spectralKurtosis", "mfcc", "loudness", "amplitudeSpectrum"]
    out2, per2 = run_gpu(odd, N, hop + 2, features=feats)
    assert per2.tolist() == [0, mo.num_frames(N * 2 + 13, N, hop + 2), 6]
    f0 = int(per2[1])
    for f in range(6):
        a = f * (hop + 2)
        one, _ = run_gpu(x[a:a + N].copy(), N, hop + 2, features=feats)
        for k in one:
            assert np.array_equal(one[k][0], out2[k][f0 + f], equal_nan=True), (k, f)
    plan = mb.Plan(N, hop + 2, SR, "hanning", feats)
    assert plan.kernel_name == "big%d" % N
    st = mb.Stream(plan)
    got, pos, rng = [], 0, np.random.default_rng(N)
    sig = odd[2]
    while pos < len(sig):
        step = int(rng.integers(1, 3 * N))
        got.append(st.push(sig[pos:pos + step]).arrays)
        pos += step
    st.close()
    plan.close()
    for k in got[0]:
        cat = np.concatenate([g[k] for g in got])
        assert np.array_equal(cat, out2[k][f0:], equal_nan=True), k


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
@pytest.mark.parametrize("N", [256, 512, 2048])
def test_degenerate_frames(N, flags):
    """silence, DC, impulse, bin-centred tone, full-scale
    square, denormal-level noise, NaN sample -- special values are results."""
    t = np.arange(N)
    rng = np.random.default_rng(5)
    frames = {
        "silence": np.zeros(N),
        "dc": np.full(N, 0.5),
        "impulse": np.eye(1, N, 7)[0],
        "tone": 0.8 * np.sin(2 * np.pi * 32 * t / N),
        "square": np.where((t // 16) % 2 == 0, 1.0, -1.0),
        "tiny": rng.standard_normal(N) * 1e-30,
        "negzero": np.where(t % 2 == 0, -0.0, 0.0),
    }
    clips = [v.astype(np.float32) for v in frames.values()]
    nanclip = mo.synth_clip(2, N).copy()
    nanclip[N // 3] = np.nan
    clips.append(nanclip)
    out, per = run_gpu(clips, N, N, flags=flags)
    assert per.tolist() == [1] * len(clips)
    names = list(frames) + ["nan"]
    for i, nm in enumerate(names):
        one = {k: v[i:i + 1] for k, v in out.items()}
        try:
            verify(one, clips[i], N, N, flags=flags, max_banded_frac=1.0)
        except AssertionError as e:
            raise AssertionError("frame %r (N=%d): %s" % (nm, N, e))


def test_feature_subsets_match_full_run(golden_audio):
    """A feature costs nothing when not requested, and does not change the others: the float32 kernels' bits do not
    depend on the feature set (MB_FLAG_NO_REFINE isolates them).  In the default, adaptive mode WHICH frames are
    redone with the exact FFT depends on the features asked for (mb_adaptive.cuh bounds only those), so a subset run
    agrees with the full run bit for bit on the frames neither redid and within the parity tolerance everywhere."""
    NR = _capi.MB_FLAG_NO_REFINE
    x = golden_audio["sound1"][:30000]
    full, _ = run_gpu(x, 2048, 512, flags=NR)
    full_a, _ = run_gpu(x, 2048, 512)
    ref = oracle_concat([x], 2048, 512)
    for feats in (["mfcc"], ["zcr", "buffer"], ["spectralRolloff", "loudness"], ["complexSpectrum"],
                  ["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"],  # config 3
                  [f for f in mb.FEATURES if f not in ("buffer", "complexSpectrum", "amplitudeSpectrum", "powerSpectrum")],
                  ["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis", "rms"]):
        # (the warp kernel is instantiated for three fixed feature sets -- all, config 3, all but the big arrays --
        # and takes every other set at run time: all of them must agree bit for bit)
        sub, _ = run_gpu(x, 2048, 512, features=feats, flags=NR)
        for k, v in sub.items():
            assert np.array_equal(v, full[k], equal_nan=True), (feats, k)
        sub, _ = run_gpu(x, 2048, 512, features=feats)
        assert not parity.compare_all(sub, ref, 2048), feats  # adaptive: flat tolerance, no noise band
    assert not parity.compare_all(full_a, ref, 2048)
    for N in (256, 512, 1024):  # the multi-frame kernel has fixed sets too
        full, _ = run_gpu(x, N, N // 2, flags=NR)
        ref = oracle_concat([x], N, N // 2)
        for feats in (["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"],
                      [f for f in mb.FEATURES if f not in ("buffer", "complexSpectrum", "amplitudeSpectrum", "powerSpectrum")],
                      ["mfcc", "zcr"], ["complexSpectrum", "spectralRolloff"]):
            sub, _ = run_gpu(x, N, N // 2, features=feats, flags=NR)
            for k, v in sub.items():
                assert np.array_equal(v, full[k], equal_nan=True), (N, feats, k)
            sub, _ = run_gpu(x, N, N // 2, features=feats)
            assert not parity.compare_all(sub, ref, N), (N, feats)


def test_device_memory_call_matches_host_call(golden_audio):
    import torch
    x = golden_audio["sound2"][:100000]
    N, hop = 2048, 512
    host, per = run_gpu(x, N, hop)
    plan = mb.Plan(N, hop, SR)
    nf = int(per[0])
    dx = torch.from_numpy(x).cuda()
    outs = {k: torch.zeros(s, dtype=torch.int32 if d == np.int32 else torch.float32, device="cuda")
            for k, (s, d) in plan.output_shapes(nf).items()}
    plan.set_stream(torch.cuda.current_stream().cuda_stream)
    plan.extract_device(dx.data_ptr(), dx.numel(), np.array([0]), np.array([len(x)]),
                        {k: v.data_ptr() for k, v in outs.items()}, sync=False)
    torch.cuda.synchronize()
    assert plan.launch_count >= 1
    for k, v in outs.items():
        assert np.array_equal(v.cpu().numpy(), host[k], equal_nan=True), k
    plan.close()


def test_host_produced_rows_equal_device_produced_rows(golden_audio):
    """`buffer` and powerSpectrum written by the host threads of a host-memory call (mb_set_host_rows 1, the default),
    and the mirrored half of complexSpectrum as well (mode 2), are bit for bit what the device writes (mode 0).  sound3
    is tonal: many of its frames are redone by the exact kernel, whose upper half is NOT the mirror image of its lower
    half to the last bit -- those rows must come from the device in every mode."""
    try:
        for name, N, hop in (("sound1", 2048, 512), ("sound1", 512, 512), ("sound3", 1024, 256), ("sound3", 256, 256),
                             ("sound1", 4096, 1024)):
            x = golden_audio[name][:120000]
            res = {}
            for mode in (0, 1, 2):
                mb.set_host_rows(mode)
                res[mode], _ = run_gpu(x, N, hop)
            for mode in (1, 2):
                assert set(res[mode]) == set(res[0])
                for k in res[0]:
                    assert np.array_equal(res[mode][k], res[0][k], equal_nan=True), (name, N, mode, k)
    finally:
        mb.set_host_rows(-1)


def test_host_mirrored_rows_with_special_frames():
    """mode 2 on frames with NaN, -0 and silence (the sign of a mirrored zero and the bits of a mirrored NaN must be the
    device's), on plans that never refine (MB_FLAG_NO_REFINE: every row is mirrored by the host) and in every float32
    kernel family."""
    rng = np.random.default_rng(5)
    GEN = _capi.MB_FLAG_GENERIC_KERNEL
    try:
        for N, flags in ((2048, 0), (2048, NO_REFINE), (512, NO_REFINE), (256, NO_REFINE), (2048, NO_REFINE | GEN), (4096, NO_REFINE),
                         (32768, NO_REFINE)):
            x = (0.3 * rng.standard_normal(N * 24)).astype(np.float32)
            x[N * 3: N * 4] = 0.0
            x[N * 5: N * 6] = -0.0
            x[N * 7 + 11] = np.nan
            x[N * 9 + 5] = -np.nan
            res = {}
            for mode in (0, 2):
                mb.set_host_rows(mode)
                plan = mb.Plan(N, N, 44100.0, "hanning", ["complexSpectrum", "amplitudeSpectrum", "rms"], device=0, flags=flags)
                out, _ = plan.extract_host(x, np.array([0], np.int64), np.array([len(x)], np.int64))
                plan.close()
                res[mode] = out
            for k in res[0]:
                a, b = np.asarray(res[0][k]), np.asarray(res[2][k])
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (N, flags, k, np.argwhere(a.view(np.uint32) != b.view(np.uint32))[:4])
    finally:
        mb.set_host_rows(-1)


def test_host_mirrored_rows_over_several_chunks():
    """mode 2 on a call of several pipeline chunks whose frames are partly redone by the exact kernel (white noise: a few
    per cent): every chunk's redone rows come from the device, the others from the host threads -- same bits as mode 0."""
    rng = np.random.default_rng(11)
    clips = [(0.3 * rng.standard_normal(441000)).astype(np.float32) for _ in range(6)]
    data, off, ln = mb.meyda._normalize_clips(clips)
    feats = mb.FEATURES
    res, refined = {}, {}
    try:
        for mode in (0, 2):
            mb.set_host_rows(mode)
            plan = mb.Plan(2048, 512, SR, "hanning", feats, device=0)
            res[mode], _ = plan.extract_host(data, off, ln)
            refined[mode] = plan.refined_frames
            plan.close()
    finally:
        mb.set_host_rows(-1)
    assert refined[0] == refined[2] and refined[0] > 0
    assert len(res[0]["complex_real"]) * (33 << 10) > (128 << 20)  # more than two 64 MiB chunks
    for k in res[0]:
        a, b = np.asarray(res[0][k]), np.asarray(res[2][k])
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (k, np.argwhere(a.view(np.uint32) != b.view(np.uint32))[:4])


def test_clip_sharding_is_bit_identical():
    """mb_extract_multi over two plans on the same device == single call."""
    clips = [mo.synth_clip(40 + i, 3000 + 977 * i) for i in range(9)]
    data, off, ln = mb.meyda._normalize_clips(clips)
    single, per = run_gpu(clips, 1024, 256)
    plans = [mb.Plan(1024, 256, SR), mb.Plan(1024, 256, SR), mb.Plan(1024, 256, SR)]
    multi, per2 = mb.extract_multi(plans, data, off, ln)
    assert per.tolist() == per2.tolist()
    for k in single:
        assert np.array_equal(single[k], multi[k], equal_nan=True), k
    for p in plans:
        p.close()


def test_streaming_matches_batch():
    x = mo.synth_clip(77, 20000)
    for N, hop in ((512, 128), (1024, 1024), (256, 600)):
        batch, per = run_gpu(x, N, hop)
        plan = mb.Plan(N, hop, SR)
        st = mb.Stream(plan)
        got = []
        pos = 0
        rng = np.random.default_rng(N)
        while pos < len(x):
            step = int(rng.integers(1, 3000))
            got.append(st.push(x[pos:pos + step]).arrays)
            pos += step
        st.close()
        plan.close()
        cat = {k: np.concatenate([g[k] for g in got]) for k in batch}
        assert len(cat["rms"]) == per[0]
        for k in batch:
            assert np.array_equal(cat[k], batch[k], equal_nan=True), (N, hop, k)


def test_streaming_fixed_buffers_replay_a_cuda_graph():
    """One buffer per push (the onaudioprocess cadence): the push shape repeats, so it is captured once and
    replayed as a CUDA graph; results stay bit-identical to the batch call, also across a reset."""
    x = mo.synth_clip(78, 40000)
    for N, hop, block in ((2048, 512, 512), (512, 512, 512), (1024, 256, 768)):
        batch, per = run_gpu(x, N, hop)
        plan = mb.Plan(N, hop, SR)
        st = mb.Stream(plan)
        for _round in range(2):
            got = [st.push(x[pos:pos + block]).arrays for pos in range(0, len(x), block)]
            cat = {k: np.concatenate([g[k] for g in got]) for k in batch}
            assert len(cat["rms"]) == per[0]
            for k in batch:
                assert np.array_equal(cat[k], batch[k], equal_nan=True), (N, hop, k, _round)
            st.reset()
            st.push(x[:0])
        assert st.graph_launches > 0.8 * 2 * (len(x) // block - N // block - 4), st.graph_launches
        st.close()
        plan.close()


def test_meyda_class_get_and_callback(golden_audio):
    """The reference's usage: construct, start(features), per-buffer callback
    with the get([...]) object; get('name') returns the bare value."""
    x = golden_audio["sound1"]
    seen = []
    m = mb.Meyda(mb.AudioContext(44100), x, 512, callback=seen.append)
    m.start(["rms", "zcr", "loudness", "mfcc", "complexSpectrum"])
    assert m.process() == 325 and len(seen) == 325
    ref = c_oracle.extract(x, 512, 512, SR)
    f = seen[100]
    assert set(f) == {"rms", "zcr", "loudness", "mfcc", "complexSpectrum"}
    assert isinstance(f["rms"], float) and f["zcr"] == int(ref["zcr"][100])
    assert set(f["loudness"]) == {"specific", "total"} and f["loudness"]["specific"].shape == (24,)
    assert set(f["complexSpectrum"]) == {"real", "imag"} and f["complexSpectrum"]["real"].shape == (512,)
    assert abs(f["rms"] - ref["rms"][100]) <= 1e-3 * ref["rms"][100]
    assert isinstance(m.get("spectralCentroid"), float)
    assert np.array_equal(m.get("buffer"), x[324 * 512:325 * 512])
    m.stop()
    m.setSource(x[:5000])
    m.windowingFunction = "hamming"
    assert m.process() == 9 and len(seen) == 325  # stopped: no more callbacks
    got = m.get(["rms", "bogus"])
    assert list(got) == ["rms"]


def test_batched_extract_callback_order(golden_audio):
    x = golden_audio["sound3"][:8192]
    order = []
    res = mb.extract([x, x[:3000]], 1024, 512, features=["zcr", "spectralRolloff"], callback=order.append)
    assert res.total_frames == 15 + 4 and len(order) == 19
    assert order[3]["zcr"] == res.value(3, "zcr") and list(res.clip(1)) == list(range(15, 19))


def test_errors_through_the_abi():
    plan = mb.Plan(512, 512, SR, features=["rms"])
    data = np.zeros(1000, np.float32)
    with pytest.raises(mb.MeydaNativeError) as ei:
        plan.extract_host(data, np.array([600], np.int64), np.array([600], np.int64))
    assert ei.value.status == _capi.MB_ERR_OUT_OF_RANGE
    o = _capi.Outputs()
    import ctypes as C
    i64p = C.POINTER(C.c_int64)
    off, ln = np.array([0], np.int64), np.array([1000], np.int64)
    st = plan._L.mb_extract(plan.handle, data.ctypes.data, 1000, off.ctypes.data_as(i64p), ln.ctypes.data_as(i64p), 1,
                            C.byref(o), 0)
    assert st == _capi.MB_ERR_MISSING_OUTPUT and b"rms" in plan._L.mb_last_error()
    plan.close()


def test_plan_tables_equal_oracle_tables():
    for N in (256, 512, 2048, 32768):
        plan = mb.Plan(N, N, SR, "hamming")
        t = plan.tables()
        plan.close()
        assert np.array_equal(t["window"], mo.hamming(N))
        assert np.array_equal(t["bbLimits"], mo.bark_band_limits(mo.bark_scale(N, SR), N // 2))
        assert np.array_equal(t["melBins"], mo.mel_bins(N, SR).astype(np.int32))


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
def test_full_size_properties(flags):
    """BASELINE sizes through size-independent properties: rms^2*N == energy,
    power == amp^2, loudness.total == sum(specific), clip order independence,
    and oracle parity on the first and last clip of the batch."""
    import torch
    N, hop, n_clips, L = 2048, 512, 64, 441000
    g = torch.Generator(device="cuda").manual_seed(1234)
    x = (torch.rand(n_clips, L, device="cuda", generator=g) - 0.5) * 0.5
    t = torch.arange(L, device="cuda", dtype=torch.float32) / SR
    x += 0.3 * torch.sin(2 * np.pi * (110.0 * (1 + torch.arange(n_clips, device="cuda"))[:, None]) * t[None, :])
    feats = ["rms", "energy", "zcr", "amplitudeSpectrum", "powerSpectrum", "loudness", "spectralCentroid", "mfcc"]
    plan = mb.Plan(N, hop, SR, features=feats, flags=flags)
    plan.set_stream(torch.cuda.current_stream().cuda_stream)  # same stream as the torch.zeros fills below
    nf_clip = mo.num_frames(L, N, hop)
    nf = nf_clip * n_clips
    outs = {k: torch.zeros(s, dtype=torch.int32 if d == np.int32 else torch.float32, device="cuda")
            for k, (s, d) in plan.output_shapes(nf).items()}
    off = np.arange(n_clips, dtype=np.int64) * L
    ln = np.full(n_clips, L, np.int64)
    plan.extract_device(x.data_ptr(), x.numel(), off, ln, {k: v.data_ptr() for k, v in outs.items()})
    assert torch.allclose(outs["rms"] ** 2 * N, outs["energy"], rtol=1e-5)
    assert torch.equal(outs["power_spectrum"], outs["amplitude_spectrum"] * outs["amplitude_spectrum"])
    assert torch.allclose(outs["loudness_total"], outs["loudness_specific"].sum(1), rtol=1e-5)
    # reversed clip order gives the reversed result, bit for bit
    outs2 = {k: torch.zeros_like(v) for k, v in outs.items()}
    plan.extract_device(x.data_ptr(), x.numel(), off[::-1].copy(), ln, {k: v.data_ptr() for k, v in outs2.items()})
    for k in ("zcr", "mfcc", "spectral_centroid"):
        a = outs[k].reshape(n_clips, nf_clip, -1)
        b = outs2[k].reshape(n_clips, nf_clip, -1).flip(0)
        assert torch.equal(a, b) or torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), k
    plan.close()
    for c in (0, n_clips - 1):
        sl = slice(c * nf_clip, (c + 1) * nf_clip)
        verify({k: v[sl].cpu().numpy() for k, v in outs.items()}, x[c].cpu().numpy(), N, hop, flags=flags)


def test_ragged_clips_stay_on_the_warp_kernel():
    """Odd clip lengths put most frames off 16-byte alignment: the bufferSize-2048 kernel loads
    those by lanes instead of TMA rather than falling back to the generic kernel."""
    N, hop = 2048, 512
    lens = [2048 + 512 * 3 + 1, 5000, 2047, 2048, 7777, 4099]
    clips = [mo.synth_clip(60 + i, L) for i, L in enumerate(lens)]
    data, off, ln = mb.meyda._normalize_clips(clips)
    assert any(int(o) % 4 for o in off)
    plan = mb.Plan(N, hop, SR)
    out, per = plan.extract_host(data, off, ln)
    assert plan.kernel_name == "warp2048"
    plan.close()
    assert per.tolist() == [mo.num_frames(L, N, hop) for L in lens]
    verify(out, clips, N, hop)
    # and an odd hop
    out, per = run_gpu(clips[1], N, 333)
    verify(out, clips[1], N, 333)


@pytest.mark.parametrize("N", [256, 512, 1024])
def test_multi_frame_warp_kernel_groups(N):
    """bufferSize 256 / 512 / 1024 run 8 / 4 / 2 frames per warp at a time: a frame's bits must not depend on its
    neighbours in the group (partial last groups, groups that straddle clips, frames off the 16-byte grid,
    one rescaled or NaN frame next to ordinary ones), and the kernel must agree with the oracle."""
    hop = N // 4
    lens = [N + hop * 6 + 1, N, N - 1, N + hop, 3 * N + 5, 0, N + 2 * hop + 3]
    clips = [mo.synth_clip(80 + i, L) for i, L in enumerate(lens)]
    clips[4] = clips[4].copy()
    clips[4][N:N + N // 2] *= np.float32(1e-30)   # a stretch the float32 squares would underflow on
    clips[6] = clips[6].copy()
    clips[6][N // 2 + 1] = np.nan  # (not on a frame's first sample: see DESIGN.md section 5 on NaN x 0)
    data, off, ln = mb.meyda._normalize_clips(clips)
    plan = mb.Plan(N, hop, SR)
    out, per = plan.extract_host(data, off, ln)
    assert plan.kernel_name == "warpmf%d" % N
    assert per.tolist() == [mo.num_frames(L, N, hop) for L in lens]
    # every frame on its own (a launch of one frame = a group of one valid frame) gives the same bits
    row = 0
    for c, L in zip(clips, lens):
        for f in range(mo.num_frames(L, N, hop)):
            one, _ = plan.extract_host(c[f * hop:f * hop + N].copy(), np.zeros(1, np.int64), np.array([N], np.int64))
            for k in out:
                assert np.array_equal(one[k][0], out[k][row], equal_nan=True), (N, k, row)
            row += 1
    plan.close()
    verify(out, [c for c in clips if len(c) >= N], N, hop, max_banded_frac=1.0)


# ---- the parameters the reference keeps as constants (mb_plan_create_ex)
PARAM_SETS = [
    pytest.param(dict(numBarkBands=30, numMelFilters=40, numMfccCoefficients=20, rolloffFraction=0.85), id="30-40-20-0.85"),
    pytest.param(dict(numBarkBands=12), id="12-bands"),  # perceptualSharpness.js:8 reads past the end: NaN
    pytest.param(dict(numBarkBands=64, numMelFilters=128, numMfccCoefficients=128, rolloffFraction=1.0), id="maxima"),
    pytest.param(dict(numBarkBands=1, numMelFilters=1, numMfccCoefficients=1, rolloffFraction=0.5), id="minima"),
    pytest.param(dict(numMfccCoefficients=5), id="5-coefficients"),
]


@pytest.mark.parametrize("params", PARAM_SETS)
@pytest.mark.parametrize("flags", FLAG_VARIANTS)
@pytest.mark.parametrize("N,hop", [(512, 512), (2048, 512), (256, 128), (4096, 4096)])
def test_plan_parameters(golden_audio, N, hop, flags, params):
    """Oracle: the numpy restatement with the same parameters, itself pinned to the reference's JavaScript run with
    them (tests/test_js_pin.py::test_oracle_parameters_reproduce_the_reference_javascript).  sound1 needs no noise band."""
    x = golden_audio["sound1"][:40 * N if N <= 512 else 70000]
    data, off, ln = mb.meyda._normalize_clips(x)
    plan = mb.Plan(N, hop, SR, "hanning", flags=flags, **params)
    try:
        # never a warp kernel (built around 24 / 26 / 13 / 0.99); the multi-warp-per-frame kernels share the generic epilogue
        assert plan.kernel_name.startswith(("generic", "exact", "big")), plan.kernel_name
        nb, nc = params.get("numBarkBands", 24), params.get("numMfccCoefficients", 13)
        assert (plan.numBarkBands, plan.numMfccCoefficients) == (nb, nc)
        assert plan.rolloffFraction == params.get("rolloffFraction", 0.99)
        t = plan.tables()
        assert np.array_equal(t["bbLimits"], mo.bark_band_limits(mo.bark_scale(N, SR), N // 2, nb))
        assert np.array_equal(t["melBins"], np.minimum(mo.mel_bins(N, SR, plan.numMelFilters), N // 2).astype(np.int32))
        out, per = plan.extract_host(data, off, ln)
        _, lay = plan.query(ln)
        assert (lay.num_bark_bands, lay.num_mfcc) == (nb, nc)
    finally:
        plan.close()
    assert out["loudness_specific"].shape == (per[0], nb) and out["mfcc"].shape == (per[0], nc)
    ref = mo.extract(x, N, hop, SR, "hanning", params=params)
    assert ref["mfcc"].shape == out["mfcc"].shape
    banded = parity.compare_all(out, ref, N, noise_band=None, exact=bool(flags & EXACT))
    assert not any(banded.values()), banded
    if nb < 16:
        assert np.isnan(out["perceptual_sharpness"]).all()


def test_default_parameters_keep_the_tuned_kernels():
    for N, name in [(2048, "warp2048"), (512, "warpmf512")]:
        plan = mb.Plan(N, N // 4, SR, "hanning", numBarkBands=24, numMelFilters=26, numMfccCoefficients=13, rolloffFraction=0.99)
        assert plan.kernel_name == name
        plan.close()
        plan = mb.Plan(N, N // 4, SR, "hanning", rolloffFraction=0.95)
        assert plan.kernel_name == "generic"
        plan.close()


def test_streaming_and_sharding_with_parameters(golden_audio):
    """The packed streaming outputs and the multi-device split use the plan's row widths."""
    params = dict(numBarkBands=30, numMelFilters=40, numMfccCoefficients=20)
    x = golden_audio["sound2"][:30000]
    N, hop = 512, 256
    plan = mb.Plan(N, hop, SR, "hanning", **params)
    try:
        data, off, ln = mb.meyda._normalize_clips(x)
        batch, _ = plan.extract_host(data, off, ln)
        st = mb.Stream(plan)
        parts = [st.push(x[i:i + 1000]).arrays for i in range(0, len(x), 1000)]
        st.close()
    finally:
        plan.close()
    for k in ("mfcc", "loudness_specific", "spectral_rolloff"):
        got = np.concatenate([p[k] for p in parts])
        assert got.shape == batch[k].shape and np.array_equal(got, batch[k], equal_nan=True), k
    res = mb.extract([x, x[:5000]], N, hop, features=["mfcc", "loudness"], devices=[0, 0], **params)
    assert np.array_equal(res["mfcc"][:len(batch["mfcc"])], batch["mfcc"], equal_nan=True)
    assert res["loudness"]["specific"].shape[1] == 30


# ---- the CUDA path against the reference's own JavaScript outputs (no oracle in between)
def _js_case_ref(js, ci):
    """One case of tests/golden/js_reference_vectors.npz (what the unmodified .js files returned under minijs,
    tools/make_js_golden.py) in the oracle's result layout."""
    g = lambda k: np.asarray(js["%d/%s" % (ci, k)])  # noqa: E731
    ref = {k: g(k).reshape(1).astype(np.float64) for k in parity.NUMBER_FIELDS.values() if "%d/%s" % (ci, k) in js}
    ref["spectralRolloff"] = g("spectralRolloff").reshape(1).astype(np.float64)
    ref["zcr"] = g("zcr").reshape(1)
    for k in ("buffer", "amplitudeSpectrum", "powerSpectrum", "mfcc"):
        ref[k] = g(k)[None].astype(np.float32)
    ref["complexSpectrum"] = {"real": g("complexSpectrum.real")[None].astype(np.float32),
                              "imag": g("complexSpectrum.imag")[None].astype(np.float32)}
    ref["loudness"] = {"specific": g("loudness.specific")[None].astype(np.float32),
                       "total": g("loudness.total").reshape(1).astype(np.float64)}
    return ref


@pytest.mark.parametrize("flags", FLAG_VARIANTS)
def test_cuda_against_the_reference_javascript_vectors(golden_audio, flags):
    js = np.load(os.path.join(os.path.dirname(__file__), "golden", "js_reference_vectors.npz"))
    cases = [str(c).split("/") for c in js["cases"]]
    exact = bool(flags & EXACT)
    checked = 0
    for ci, (clip, N, f, window) in enumerate(cases):
        N, f = int(N), int(f)
        adaptive = is_adaptive(flags, N)
        if clip == "nan" and not (exact or adaptive):
            continue  # (a NaN on a window zero; only the exact arithmetic keeps the reference's finite imag plane)
        sig = golden_audio[clip][f * N:(f + 1) * N] if clip in golden_audio else mo.degenerate_frame(clip, N)
        out, per = run_gpu(sig, N, window=window, flags=flags)
        assert per.tolist() == [1]
        noise = None if (exact or adaptive) else mo.noise_band(sig, N, N, SR, window)
        if clip == "square" and not (exact or adaptive):
            # DESIGN.md section 5, second corner: the DC bin of this frame cancels to an exact 0 in the reference's
            # arithmetic, so TWO mel filters are empty there (ln 0 = -Infinity twice, opposite DCT signs: NaN in
            # coefficients 8..12); a float32 FFT leaves ~1e-9 of rounding noise in that bin and gets -+Infinity
            m = out.pop("mfcc")[0]
            assert np.all(m[:8] == -np.inf) and not np.isfinite(m[8:]).any()
        try:
            parity.compare_all(out, _js_case_ref(js, ci), N, noise_band=noise, exact=exact)
        except AssertionError as e:
            raise AssertionError("case %d %s N=%d frame %d %s: %s" % (ci, clip, N, f, window, e)) from None
        checked += 1
    assert checked == (17 if (exact or is_adaptive(flags)) else 16)


@pytest.mark.parametrize("flags", [pytest.param(_capi.MB_FLAG_GENERIC_KERNEL, id="generic"), pytest.param(EXACT, id="exact")])
def test_cuda_parameters_against_the_reference_javascript_vectors(golden_audio, flags):
    """tests/golden/js_reference_params.npz: NUM_BARK_BANDS 30 / 12 / 40, 40 filters, 20 coefficients, fraction 0.85."""
    js = np.load(os.path.join(os.path.dirname(__file__), "golden", "js_reference_params.npz"))
    for ci, (clip, N, f, window, nb, edited) in enumerate(str(c).split("/") for c in js["cases"]):
        N, f, nb = int(N), int(f), int(nb)
        if clip == "sound3" and not flags & EXACT:
            continue  # a near-pure tone: bands that hold only FFT rounding noise need the noise band (see verify())
        params = dict(numBarkBands=nb)
        if int(edited):
            params.update(numMelFilters=40, numMfccCoefficients=20, rolloffFraction=0.85)
        sig = golden_audio[clip][f * N:(f + 1) * N]
        data, off, ln = mb.meyda._normalize_clips(sig)
        feats = ["loudness", "perceptualSpread", "perceptualSharpness", "mfcc", "spectralRolloff", "spectralCentroid"]
        plan = mb.Plan(N, N, SR, window, feats, flags=flags, **params)
        try:
            out, _ = plan.extract_host(data, off, ln)
        finally:
            plan.close()
        tol = 5e-6 if flags & EXACT else 1e-3
        for field, key in (("loudness_total", "loudness.total"), ("perceptual_spread", "perceptualSpread"),
                           ("perceptual_sharpness", "perceptualSharpness"), ("spectral_centroid", "spectralCentroid"),
                           ("spectral_rolloff", "spectralRolloff")):
            parity.assert_numbers(key, out[field], np.asarray(js["%d/%s" % (ci, key)]).reshape(1), tol=tol)
        parity.assert_numbers("loudness.specific", out["loudness_specific"], js["%d/loudness.specific" % ci][None], tol=tol)
        parity.assert_numbers("mfcc", out["mfcc"], js["%d/mfcc" % ci][None], tol=tol)
