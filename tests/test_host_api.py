"""Host-side mirror of the reference interface: names, shapes, error behaviour
(no GPU needed for these)."""
import numpy as np
import pytest

import meyda_b200 as mb
from meyda_b200.meyda import _normalize_clips, _split_features
from oracle import meyda_oracle as mo


def test_feature_info_is_the_reference_table():
    assert mb.featureInfo == mo.FEATURE_INFO
    assert len(mb.FEATURES) == 18 and mb.FEATURES[0] == "buffer" and mb.FEATURES[-1] == "mfcc"
    types = [v["type"] for v in mb.featureInfo.values()]
    assert types.count("number") == 12 and types.count("array") == 4 and types.count("multipleArrays") == 2
    assert mb.featureInfo["loudness"]["arrayNames"] == {"1": "total", "2": "specific"}


def test_is_power_of_two_like_utils_js():
    for v in (1, 2, 256, 32768):
        assert mb.isPowerOfTwo(v)
    for v in (0, 3, 600, -4, 2.5, None, "x", float("nan")):
        assert not mb.isPowerOfTwo(v)


def test_constructor_errors_match_reference_messages():
    ctx = mb.AudioContext(44100)
    with pytest.raises(mb.MeydaError, match="Buffer size is not a power of two: Meyda will not run."):
        mb.Meyda(ctx, np.zeros(4096, np.float32), 600)
    with pytest.raises(mb.MeydaError, match="Buffer size is not a power of two"):
        mb.Meyda(ctx, np.zeros(4096, np.float32), None)  # `bufSize || 256` default is dead code
    with pytest.raises(mb.MeydaError, match="AudioContext wasn't specified: Meyda will not run."):
        mb.Meyda(None, np.zeros(4096, np.float32), 512)
    m = mb.Meyda(ctx, np.zeros(4096, np.float32), 512)
    assert m.windowingFunction == "hanning" and m.featureInfo is mb.featureInfo
    with pytest.raises(mb.MeydaError, match="Invalid Feature Format"):
        m.get(42)
    with pytest.raises(TypeError):
        m.get("notAFeature")


def test_split_features_list_drops_unknown(capsys):
    feats, single = _split_features(["rms", "bogus", "zcr", "rms"])
    assert feats == ["rms", "zcr"] and not single
    assert "bogus" in capsys.readouterr().err
    assert _split_features("mfcc") == (["mfcc"], True)
    with pytest.raises(mb.MeydaError, match="Invalid Feature Format"):
        _split_features(3.5)


def test_normalize_clips_forms():
    a = np.arange(5, dtype=np.float32)
    b = np.arange(3, dtype=np.float32)
    d, off, ln = _normalize_clips([a, b])
    assert d.tolist() == [0, 1, 2, 3, 4, 0, 1, 2] and off.tolist() == [0, 5] and ln.tolist() == [5, 3]
    d, off, ln = _normalize_clips(np.zeros((3, 7), np.float32))
    assert off.tolist() == [0, 7, 14] and ln.tolist() == [7, 7, 7]
    d, off, ln = _normalize_clips(a)
    assert off.tolist() == [0] and ln.tolist() == [5]
    d, off, ln = _normalize_clips({"data": a, "offsets": [1], "lengths": [3]})
    assert off.tolist() == [1] and ln.tolist() == [3]
    d, off, ln = _normalize_clips([])
    assert d.size == 0 and len(off) == 0


def test_feature_mask_bits():
    assert mb.feature_mask(["buffer"]) == 1 and mb.feature_mask(["mfcc"]) == 1 << 17
    assert mb.feature_mask(mb.FEATURES) == (1 << 18) - 1


def test_no_gpu_means_loud_failure():
    import ctypes as C
    from meyda_b200 import _capi
    n = C.c_int(0)
    _capi.lib().mb_device_count(C.byref(n))
    if n.value:
        pytest.skip("a GPU is present")
    with pytest.raises(mb.MeydaNativeError, match="no CPU fallback"):
        mb.extract(np.zeros(4096, np.float32), 512, features=["rms"])
