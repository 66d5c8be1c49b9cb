"""No-GPU checks of the WAV front end (mb_wav_parse / meyda_b200.wav_info): the part of
lib/bufferLoader.js:13-44 + decodeAudioData that is host work."""
import io
import os
import struct
import wave

import numpy as np
import pytest

import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import meyda_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_AUDIO = "/root/reference/audio"


@pytest.fixture(scope="module", autouse=True)
def _built():
    from meyda_b200.build import build
    build()


def make_wav(pcm: np.ndarray, rate=44100) -> bytes:
    b = io.BytesIO()
    w = wave.open(b, "wb")
    w.setnchannels(1 if pcm.ndim == 1 else pcm.shape[1])
    w.setsampwidth(2)
    w.setframerate(rate)
    w.writeframes(np.ascontiguousarray(pcm, dtype="<i2").tobytes())
    w.close()
    return b.getvalue()


def test_parse_plain_files():
    mono = make_wav(np.arange(100, dtype=np.int16))
    assert mb.wav_info(mono) == {"format": 1, "channels": 1, "sampleRate": 44100, "bitsPerSample": 16,
                                 "dataOffset": 44, "sampleFrames": 100}
    stereo = make_wav(np.arange(60, dtype=np.int16).reshape(30, 2), rate=22050)
    i = mb.wav_info(stereo)
    assert (i["channels"], i["sampleRate"], i["sampleFrames"], i["dataOffset"]) == (2, 22050, 30, 44)


def test_parse_skips_other_chunks_and_odd_padding():
    data = np.arange(10, dtype="<i2").tobytes()
    fmt = struct.pack("<HHIIHH", 1, 1, 8000, 16000, 2, 16)
    body = (b"WAVE" + b"LIST" + struct.pack("<I", 3) + b"abc" + b"\0"      # odd-sized chunk, padded to even
            + b"fmt " + struct.pack("<I", len(fmt)) + fmt
            + b"fact" + struct.pack("<I", 4) + struct.pack("<I", 10)
            + b"data" + struct.pack("<I", len(data)) + data)
    blob = b"RIFF" + struct.pack("<I", len(body)) + body
    i = mb.wav_info(blob)
    assert i["sampleFrames"] == 10 and i["sampleRate"] == 8000
    assert blob[i["dataOffset"]:i["dataOffset"] + 20] == data


def test_parse_extensible_and_truncated_data():
    # WAVE_FORMAT_EXTENSIBLE (0xFFFE): the real format is the first two bytes of the sub-format GUID
    ext = struct.pack("<HHIIHH", 0xFFFE, 2, 48000, 192000, 4, 16) + struct.pack("<HHI", 22, 16, 3) + \
        struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
    data = b"\1\0" * 16
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(ext)) + ext + b"data" + struct.pack("<I", 1000) + data
    i = mb.wav_info(b"RIFF" + struct.pack("<I", len(body)) + body)
    assert i["format"] == 1 and i["channels"] == 2
    assert i["sampleFrames"] == 8  # the data chunk claims 1000 bytes, the file holds 32: 8 stereo frames


def test_parse_errors():
    for blob in (b"", b"RIFF\0\0\0\0WAVX", b"RIFF\x04\0\0\0WAVE", make_wav(np.zeros(4, np.int16))[:36]):
        with pytest.raises(mb.MeydaNativeError) as e:
            mb.wav_info(blob)
        assert e.value.status == _capi.MB_ERR_INVALID_ARG


@pytest.mark.skipif(not os.path.isdir(REF_AUDIO), reason="reference checkout not mounted")
def test_reference_fixtures_parse_like_the_oracle_reader():
    """audio/sound{1,2,3}.wav: same samples as the oracle's reader and as the committed PCM fixture."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "audio_pcm16.npz"))
    for name in ("sound1", "sound2", "sound3"):
        blob = open(os.path.join(REF_AUDIO, name + ".wav"), "rb").read()
        i = mb.wav_info(blob)
        assert (i["format"], i["channels"], i["bitsPerSample"], i["sampleRate"]) == (1, 1, 16, 44100)
        pcm = np.frombuffer(blob, dtype="<i2", count=i["sampleFrames"], offset=i["dataOffset"])
        ref_float, rate = mo.read_wav_pcm16(os.path.join(REF_AUDIO, name + ".wav"))
        assert rate == 44100 and np.array_equal(pcm, z[name])
        assert np.array_equal(mo.pcm16_to_float(pcm), ref_float)


def test_parser_never_reads_outside_the_file():
    """Truncated and bit-flipped files either parse to something that fits inside the bytes given or are rejected."""
    from hypothesis import given, settings, strategies as st
    good = [make_wav(np.arange(64, dtype=np.int16)), make_wav(np.arange(120, dtype=np.int16).reshape(40, 3), rate=8000)]

    @settings(max_examples=300, deadline=None)
    @given(st.integers(0, 1), st.integers(0, 300), st.lists(st.tuples(st.integers(0, 299), st.integers(0, 255)), max_size=6),
           st.binary(max_size=40))
    def check(which, cut, flips, tail):
        b = bytearray(good[which])
        for pos, val in flips:
            if pos < len(b):
                b[pos] = val
        blob = bytes(b[:max(0, len(b) - cut)]) + tail
        try:
            i = mb.wav_info(blob)
        except mb.MeydaNativeError as e:
            assert e.status == _capi.MB_ERR_INVALID_ARG
            return
        assert i["channels"] >= 1 and i["bitsPerSample"] % 8 == 0 and i["sampleFrames"] >= 0
        assert 0 <= i["dataOffset"] <= len(blob)
        assert i["dataOffset"] + i["sampleFrames"] * i["channels"] * (i["bitsPerSample"] // 8) <= len(blob)

    check()
