"""No-GPU checks of the C-ABI library: it loads, exports every symbol the
header declares, and its device-independent entry points behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from meyda_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from meyda_b200.build import build
    build()
    return _capi.lib()


def _header_functions():
    src = open(os.path.join(ROOT, "include", "meyda_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_header_symbol(lib):
    names = _header_functions()
    assert len(names) >= 20
    assert sorted(names) == sorted(_capi.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n


def test_version_and_feature_names(lib):
    assert lib.mb_version() == 100
    from meyda_b200 import FEATURES
    assert _capi.feature_names() == FEATURES  # key order of src/feature-info.js
    for i, n in enumerate(FEATURES):
        assert lib.mb_feature_from_name(n.encode()) == i
    assert lib.mb_feature_from_name(b"nope") == -1
    assert lib.mb_feature_name(99) is None


def test_num_frames_rule(lib):
    assert lib.mb_num_frames(166400, 512, 512) == 325
    assert lib.mb_num_frames(441001, 2048, 2048) == 215
    assert lib.mb_num_frames(441000, 2048, 512) == 858
    assert lib.mb_num_frames(1323000, 2048, 512) == 2580
    assert lib.mb_num_frames(2646000, 32768, 8192) == 319
    assert lib.mb_num_frames(2047, 2048, 512) == 0
    assert lib.mb_num_frames(2048, 2048, 512) == 1


def test_plan_create_errors_without_device(lib):
    h = C.c_void_p()
    st = lib.mb_plan_create(C.byref(h), 0, 600, 600, 44100.0, 0, 1, 0)
    assert st == _capi.MB_ERR_NOT_POWER_OF_TWO
    assert lib.mb_last_error().decode() == "Buffer size is not a power of two: Meyda will not run."
    assert lib.mb_plan_create(C.byref(h), 0, 8, 8, 44100.0, 0, 1, 0) == _capi.MB_ERR_UNSUPPORTED
    assert lib.mb_plan_create(C.byref(h), 0, 512, 0, 44100.0, 0, 1, 0) == _capi.MB_ERR_INVALID_ARG
    assert lib.mb_plan_create(C.byref(h), 0, 512, 512, 44100.0, 7, 1, 0) == _capi.MB_ERR_INVALID_ARG
    assert lib.mb_plan_create(C.byref(h), 0, 512, 512, 44100.0, 0, 0, 0) == _capi.MB_ERR_INVALID_ARG
    # mb_plan_create_ex: the parameters are validated before any device is touched
    bad = [_capi.Params(65, 0, 0, 0, 0.0), _capi.Params(-1, 0, 0, 0, 0.0), _capi.Params(0, 129, 0, 0, 0.0),
           _capi.Params(0, 20, 21, 0, 0.0), _capi.Params(0, 0, 27, 0, 0.0), _capi.Params(0, 0, 0, 1, 0.0),
           _capi.Params(0, 0, 0, 0, 1.5), _capi.Params(0, 0, 0, 0, -0.1), _capi.Params(0, 0, 0, 0, float("nan"))]
    for prm in bad:
        st = lib.mb_plan_create_ex(C.byref(h), 0, 512, 512, 44100.0, 0, 1, 0, C.byref(prm))
        assert st == _capi.MB_ERR_INVALID_ARG and not h.value, (prm.num_bark_bands, prm.num_mel_filters, prm.num_mfcc)
    assert b"rolloff_fraction" in lib.mb_last_error()
    n = C.c_int(-1)
    lib.mb_device_count(C.byref(n))
    if n.value == 0:  # no GPU here: the product path must fail loudly, not fall back
        st = lib.mb_plan_create(C.byref(h), 0, 512, 512, 44100.0, 0, 1, 0)
        assert st == _capi.MB_ERR_NO_DEVICE and b"no CPU fallback" in lib.mb_last_error()
        assert not h.value


def test_struct_layouts_match_header():
    assert C.sizeof(_capi.Outputs) == 20 * C.sizeof(C.c_void_p)
    assert C.sizeof(_capi.Layout) == 8 + 4 + 4 + 4 + 4 + 8 + 8 + 4 + 4
    assert C.sizeof(_capi.Params) == 4 * 4 + 8 and _capi.Params.rolloff_fraction.offset == 16
    assert [f for f, _, _ in _capi.OUTPUT_FIELDS][:4] == ["buffer", "rms", "energy", "zcr"]


def test_product_package_never_imports_oracle():
    """The product path must not route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "meyda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/ (or", ""), os.path.join(dirpath, f)


def test_napi_addon_source_type_checks_against_the_c_abi():
    """Node and its headers are absent here, so js/addon.cc cannot be built; it is at least type-checked against
    include/meyda_b200.h and a declaration-only stand-in for node_api.h (tools/napi_stub, documented signatures)."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    r = subprocess.run([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tools", "napi_stub"),
                        os.path.join(ROOT, "js", "addon.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(ROOT, "js", "addon.cc")).read()
    for name in ("mb_plan_create_ex", "mb_extract", "mb_extract_pcm16", "mb_query_output", "mb_wav_parse", "napi_queue_async_work",
                 "mb_stream_create", "mb_stream_push", "mb_stream_reset", "mb_stream_frames_after", "mb_stream_destroy",
                 "mb_extract_multi", "mb_plan_get_params", "mb_plan_refined_frames", "mb_host_alloc", "mb_host_free",
                 "napi_create_external_arraybuffer"):
        assert name in src
    # every native entry the JavaScript facade calls is one the addon registers
    import re
    facade = open(os.path.join(ROOT, "js", "meyda_b200.js")).read()
    used = set(re.findall(r"native\.(\w+)", facade))
    registered = set(re.findall(r'\{"(\w+)", NULL, \w+, NULL, NULL, NULL, napi_default, NULL\}', src))
    assert used and used <= registered, used - registered
    # and the facade carries the reference's class surface (src/meyda.js:229-261)
    for name in ("class Meyda", "setSource (", "start (", "stop (", "get (", "windowingFunction", "featureInfo"):
        assert name in facade, name


def test_set_host_threads_validates_its_argument():
    """mb_set_host_threads touches no device: range-checked, 0 restores the automatic choice."""
    from meyda_b200 import _capi
    L = _capi.lib()
    assert L.mb_set_host_threads(4) == 0
    assert L.mb_set_host_threads(0) == 0
    assert L.mb_set_host_threads(-1) != 0
    assert b"host thread count" in L.mb_last_error()


def test_set_host_rows_validates_its_argument():
    from meyda_b200 import _capi
    L = _capi.lib()
    for mode in (2, 1, 0, -1):
        assert L.mb_set_host_rows(mode) == 0
    assert L.mb_set_host_rows(3) != 0
    assert b"host rows mode" in L.mb_last_error()
