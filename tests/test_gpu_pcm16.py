"""16-bit PCM ingest on the device: the int16 -> float32 step of decodeAudioData
and the channel pick (lib/bufferLoader.js:13-44, src/meyda.js:72) happen inside the kernels' frame load.
s / 32768 is exact in float32, so the PCM path must equal the float path on the converted samples BIT FOR BIT,
and through it the oracle."""
import numpy as np
import pytest

import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import meyda_oracle as mo
from tests import parity
from tests.test_gpu_parity import SR, run_gpu, verify
from tests.test_wav_cpu import make_wav

pytestmark = pytest.mark.gpu
EXACT = _capi.MB_FLAG_EXACT_FFT
CASES = [pytest.param(2048, 512, 0, id="warp2048"), pytest.param(2048, 2048, 0, id="warp2048-hopN"),
         pytest.param(512, 128, 0, id="generic512"), pytest.param(256, 256, 0, id="generic256"),
         pytest.param(2048, 1024, EXACT, id="exact2048"),
         pytest.param(1024, 512, EXACT | _capi.MB_FLAG_CLUSTER_FFT, id="exact-cluster1024"),
         pytest.param(2048, 512, _capi.MB_FLAG_GENERIC_KERNEL, id="generic2048")]


def golden_pcm():
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "audio_pcm16.npz"))
    return {k: z[k] for k in z.files}


def run_pcm(pcm, offsets, lengths, N, hop, flags=0, channel=0, features=mb.FEATURES):
    plan = mb.Plan(N, hop, SR, "hanning", features, flags=flags)
    try:
        return plan.extract_pcm16_host(pcm, np.asarray(offsets, np.int64), np.asarray(lengths, np.int64), channel=channel)
    finally:
        plan.close()


def assert_same_bits(a: dict, b: dict):
    assert set(a) == set(b)
    for k in a:
        parity.assert_bits(k, a[k], b[k])


@pytest.mark.parametrize("N,hop,flags", CASES)
def test_pcm_path_equals_float_path_bit_for_bit(N, hop, flags):
    pcm = golden_pcm()["sound3"][:60000]
    got, per = run_pcm(pcm, [0], [len(pcm)], N, hop, flags)
    ref, per_f = run_gpu(mo.pcm16_to_float(pcm), N, hop, flags=flags)
    assert per.tolist() == per_f.tolist() and per[0] == mo.num_frames(len(pcm), N, hop)
    assert_same_bits(got, ref)


@pytest.mark.parametrize("N,hop", [(2048, 512), (512, 512)])
def test_pcm_ragged_clips_odd_offsets_and_extremes(N, hop):
    """Clips that start off the 16-byte grid (no TMA: lanes copy the frame), a clip shorter than a buffer, an
    empty one, and the int16 extremes (-32768 -> -1.0 exactly, 32767)."""
    rng = np.random.default_rng(11)
    pcm = rng.integers(-32768, 32768, size=40000, dtype=np.int64).astype(np.int16)
    pcm[1000:1000 + N] = -32768
    pcm[5001:5001 + N:2] = 32767
    offsets = [0, 1001, 5001, 9003, 20000, 26007]
    lengths = [N + 3 * hop, N, N + hop + 1, N - 1, 0, 3 * N]
    got, per = run_pcm(pcm, offsets, lengths, N, hop)
    x = mo.pcm16_to_float(pcm)
    ref, per_f = run_gpu([x[o:o + l] for o, l in zip(offsets, lengths)], N, hop)
    assert per.tolist() == per_f.tolist() == [mo.num_frames(l, N, hop) for l in lengths]
    assert_same_bits(got, ref)
    verify(got, [x[o:o + l] for o, l in zip(offsets, lengths)], N, hop)  # and against the oracle itself


@pytest.mark.parametrize("N,hop", [(2048, 512), (1024, 1024)])
def test_interleaved_channels(N, hop):
    """getChannelData(c) of a stereo / 5-channel file: the kernel strides over the interleaved frames."""
    g = golden_pcm()
    n = 30000
    for channels in (2, 5):
        pcm = np.zeros((n, channels), np.int16)
        for c in range(channels):
            pcm[:, c] = np.roll(g["sound1"][:n], 977 * c) // (c + 1)
        for c in (0, channels - 1):
            got, per = run_pcm(pcm, [0, 4000], [n - 4000, 9000], N, hop, channel=c)
            x = mo.pcm16_to_float(pcm[:, c])
            ref, _ = run_gpu([x[0:n - 4000], x[4000:13000]], N, hop)
            assert_same_bits(got, ref)


def test_device_memory_call():
    torch = pytest.importorskip("torch")
    N, hop = 2048, 512
    pcm = golden_pcm()["sound2"][:50000]
    plan = mb.Plan(N, hop, SR, "hanning", mb.FEATURES)
    try:
        lengths = np.array([len(pcm)], np.int64)
        per, lay = plan.query(lengths)
        d_pcm = torch.from_numpy(pcm.copy()).cuda()
        outs = {k: torch.zeros(s, dtype=torch.int32 if d == np.int32 else torch.float32, device="cuda")
                for k, (s, d) in plan.output_shapes(int(lay.total_frames)).items()}
        torch.cuda.synchronize()
        plan.extract_pcm16_device(d_pcm.data_ptr(), len(pcm), 1, 0, np.zeros(1, np.int64), lengths,
                                  {k: v.data_ptr() for k, v in outs.items()})
        assert plan.kernel_name == "warp2048" and plan.launch_count == 2  # the float32 kernel + the (here empty) exact-FFT pass over the frames it flagged
        ref, _ = run_gpu(mo.pcm16_to_float(pcm), N, hop)
        assert_same_bits({k: v.cpu().numpy() for k, v in outs.items()}, ref)
    finally:
        plan.close()


def test_extract_wav_against_the_oracle():
    """WAV bytes in, `get`-shaped features out: config 1 of BASELINE.json straight from the file format."""
    g = golden_pcm()
    feats = ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid", "mfcc", "loudness"]
    res = mb.extract_wav([make_wav(g["sound1"]), make_wav(g["sound2"][:30001])], 512, features=feats)
    assert res.frames_per_clip.tolist() == [325, 58]
    ref = mo._concat([mo.extract(mo.pcm16_to_float(g["sound1"]), 512, 512, SR, "hanning", feats),
                      mo.extract(mo.pcm16_to_float(g["sound2"][:30001]), 512, 512, SR, "hanning", feats)])
    noise = mo._concat([mo.noise_band(mo.pcm16_to_float(g["sound1"]), 512, 512, SR, "hanning"),
                        mo.noise_band(mo.pcm16_to_float(g["sound2"][:30001]), 512, 512, SR, "hanning")])
    parity.compare_all(res.arrays, ref, 512, noise_band=noise)
    f0 = res.frame(0)
    assert set(f0) == set(feats) and f0["zcr"] == 26 and abs(f0["rms"] - 0.0050815644) < 1e-9  # SURVEY probe values
    stereo = np.stack([g["sound1"], g["sound1"][::-1]], axis=1)
    r2 = mb.extract_wav(make_wav(stereo), 512, features=feats, channel=0)
    for k in res.arrays:
        parity.assert_bits(k, r2.arrays[k], res.arrays[k][:325])


def _wav_bytes(fmt_tag, channels, rate, bits, payload: bytes) -> bytes:
    import struct
    block = channels * bits // 8
    fmt = struct.pack("<HHIIHH", fmt_tag, channels, rate, rate * block, block, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"data" + struct.pack("<I", len(payload)) + payload
    return b"RIFF" + struct.pack("<I", len(body)) + body


@pytest.mark.parametrize("N,hop", [(2048, 512), (512, 512), (1024, 300)])
def test_24bit_and_float_wav_payloads(N, hop):
    """WAV format 1 / 24 bits (s / 8388608, exact in float32) and format 3 / 32-bit float, mono and interleaved stereo:
    converted inside the generic kernels' frame load, bit-identical to the float path on the decoded samples."""
    rng = np.random.default_rng(24)
    n = 30000
    v24 = rng.integers(-(1 << 23), 1 << 23, size=(n, 2), dtype=np.int64)
    v24[100:100 + N, 0] = -(1 << 23)
    v24[5000:5000 + N:3, 1] = (1 << 23) - 1
    f32 = (rng.standard_normal((n, 2)) * 0.3).astype(np.float32)
    packed = (v24.astype("<i4").view(np.uint8).reshape(n, 2, 4)[:, :, :3]).copy()  # little-endian, low 3 bytes
    feats = mb.FEATURES
    for ch in (1, 2):
        for c in range(ch):
            # 24-bit
            payload = packed[:, :ch].tobytes() if ch == 2 else packed[:, 0].tobytes()
            res = mb.extract_wav(_wav_bytes(1, ch, 44100, 24, payload), N, hop=hop, features=feats, channel=c)
            x = (v24[:, c].astype(np.float32) / np.float32(8388608.0)).astype(np.float32)
            ref, per = run_gpu(x, N, hop, flags=_capi.MB_FLAG_GENERIC_KERNEL)
            assert res.frames_per_clip.tolist() == per.tolist()
            assert_same_bits(res.arrays, ref)
            # 32-bit float
            payload = f32[:, :ch].tobytes() if ch == 2 else f32[:, 0].tobytes()
            res = mb.extract_wav(_wav_bytes(3, ch, 44100, 32, payload), N, hop=hop, features=feats, channel=c)
            ref, _ = run_gpu(f32[:, c].copy(), N, hop, flags=_capi.MB_FLAG_GENERIC_KERNEL)
            assert_same_bits(res.arrays, ref)
    # and against the oracle itself on one of them
    res = mb.extract_wav(_wav_bytes(1, 1, 44100, 24, packed[:, 0].tobytes()), N, hop=hop, features=feats)
    verify(res.arrays, (v24[:, 0].astype(np.float32) / np.float32(8388608.0)).astype(np.float32), N, hop)


def test_pcm_errors():
    plan = mb.Plan(512, 512, SR, "hanning", ["rms"])
    try:
        pcm = np.zeros((1000, 2), np.int16)
        with pytest.raises(mb.MeydaNativeError) as e:
            plan.extract_pcm16_host(pcm, np.zeros(1, np.int64), np.array([1000], np.int64), channel=2)
        assert e.value.status == _capi.MB_ERR_INVALID_ARG
        with pytest.raises(mb.MeydaNativeError) as e:
            plan.extract_pcm16_host(pcm, np.array([600], np.int64), np.array([512], np.int64))
        assert e.value.status == _capi.MB_ERR_OUT_OF_RANGE
    finally:
        plan.close()
    with pytest.raises(mb.MeydaError):
        mb.extract_wav(make_wav(np.zeros(100, np.int16)), 16, features=["nope"])
    with pytest.raises(mb.MeydaNativeError) as e:
        plan2 = mb.Plan(512, 512, SR, "hanning", ["rms"])
        try:
            plan2.extract_pcm_host(np.zeros(2048, np.int16), 7, 1, np.zeros(1, np.int64), np.array([1024], np.int64))
        finally:
            plan2.close()
    assert e.value.status == _capi.MB_ERR_INVALID_ARG
    with pytest.raises(mb.MeydaError):  # 8-bit file
        import io, wave
        b = io.BytesIO(); w = wave.open(b, "wb"); w.setnchannels(1); w.setsampwidth(1); w.setframerate(8000)
        w.writeframes(bytes(64)); w.close()
        mb.extract_wav(b.getvalue(), 16, features=["rms"])


@pytest.mark.parametrize("N,hop,channels", [(2048, 512, 1), (512, 512, 1), (1024, 300, 2), (256, 64, 1)])
def test_pcm16_streaming_matches_batch(N, hop, channels):
    """Buffer-by-buffer pushes of int16 blocks (one per onaudioprocess call): bit-identical to the batch PCM call,
    which is bit-identical to the float path; repeating push shapes replay as a CUDA graph."""
    g = golden_pcm()["sound2"][:24000]
    pcm = np.stack([np.roll(g, 555 * c) for c in range(channels)], axis=1) if channels > 1 else g
    ch = channels - 1
    batch, per = run_pcm(pcm, [0], [len(g)], N, hop, channel=ch)
    plan = mb.Plan(N, hop, SR)
    st = mb.Stream(plan, pcm16_channels=channels, channel=ch)
    block = 512
    got = [st.push(pcm[pos:pos + block]).arrays for pos in range(0, len(g), block)]
    cat = {k: np.concatenate([x[k] for x in got]) for k in batch}
    assert len(cat["rms"]) == per[0]
    assert_same_bits(cat, batch)
    if block % hop == 0:  # the push shape repeats: captured once, then replayed
        assert st.graph_launches > 10
    with pytest.raises(mb.MeydaNativeError):  # a float block on a PCM stream
        o = plan.alloc_host_outputs(4)
        st.push_into(np.zeros(64, np.float32), o)
    st.close()
    plan.close()
