"""Pins the oracle against the reference's OWN JavaScript.

tests/golden/js_reference_vectors.npz holds what the unmodified files
lib/jsfft/{complex_array,fft}.js, src/utils.js, src/extractors/*.js and the
compute* method bodies of src/meyda.js return on seventeen frames when executed by
oracle/minijs.py (an ES5-subset interpreter written for this purpose, since no
JavaScript engine exists in the image; tools/make_js_golden.py is the generator).
Both oracle restatements must reproduce them: Float32Array results bit for bit,
Numbers to 1e-12 (relative; 1e-13 absolute where a value is pure cancellation noise).  Where /root/reference is mounted, two cases are re-executed
live to show the vectors are reproducible from the reference sources."""
import os

import numpy as np
import pytest

from oracle import c_oracle, meyda_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 44100.0
NUMBERS = ["rms", "energy", "zcr", "spectralCentroid", "spectralFlatness", "spectralSlope", "spectralRolloff",
           "spectralSpread", "spectralSkewness", "spectralKurtosis", "perceptualSpread", "perceptualSharpness"]


@pytest.fixture(scope="module")
def js():
    return np.load(os.path.join(ROOT, "tests", "golden", "js_reference_vectors.npz"))


def _signal(golden_audio, clip, N, f):
    if clip not in golden_audio:
        return mo.degenerate_frame(clip, N)
    return golden_audio[clip][f * N:(f + 1) * N]


def _bits(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and bool(((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))).all())


def _num(a, b):
    a, b = float(a), float(b)
    if np.isnan(a) or np.isnan(b):
        return np.isnan(a) and np.isnan(b)
    if np.isinf(a) or np.isinf(b):
        return a == b
    return abs(a - b) <= 1e-12 * max(abs(a), abs(b)) + 1e-13  # 1e-13 absolute: values that are pure cancellation noise


def _check(js, ci, r, who):
    for k in NUMBERS:
        assert _num(r[k][0], js["%d/%s" % (ci, k)]), (who, ci, k, r[k][0], float(js["%d/%s" % (ci, k)]))
    assert _num(r["loudness"]["total"][0], js["%d/loudness.total" % ci]), (who, ci, "loudness.total")
    assert _bits(r["loudness"]["specific"][0], js["%d/loudness.specific" % ci]), (who, ci, "loudness.specific")
    assert _bits(r["mfcc"][0], js["%d/mfcc" % ci]), (who, ci, "mfcc", r["mfcc"][0], js["%d/mfcc" % ci])
    assert _bits(r["amplitudeSpectrum"][0], js["%d/amplitudeSpectrum" % ci]), (who, ci, "amplitudeSpectrum")
    assert _bits(r["powerSpectrum"][0], js["%d/powerSpectrum" % ci]), (who, ci, "powerSpectrum")
    assert _bits(r["complexSpectrum"]["real"][0], js["%d/complexSpectrum.real" % ci]), (who, ci, "real")
    assert _bits(r["complexSpectrum"]["imag"][0], js["%d/complexSpectrum.imag" % ci]), (who, ci, "imag")
    assert _bits(r["buffer"][0], js["%d/buffer" % ci]), (who, ci, "buffer")


def test_oracles_reproduce_the_reference_javascript(js, golden_audio):
    cases = [str(c).split("/") for c in js["cases"]]
    assert len(cases) == 17
    for ci, (clip, N, f, window) in enumerate(cases):
        N, f = int(N), int(f)
        sig = _signal(golden_audio, clip, N, f)
        _check(js, ci, c_oracle.extract(sig, N, N, SR, window), "C oracle")
        _check(js, ci, mo.extract(sig, N, N, SR, window), "numpy oracle")
        # plan-time tables computed by the reference's own constructor code
        t = c_oracle.plan_tables(N, SR)
        assert _bits(t["hanning"], js["%d/hanning" % ci]) and _bits(t["hamming"], js["%d/hamming" % ci])
        assert _bits(t["bark"], js["%d/barkScale" % ci])
        assert np.array_equal(t["bbLimits"], js["%d/bbLimits" % ci].astype(np.int32))
    got = [mo.is_power_of_two(v) for v in js["isPowerOfTwo/in"]]
    assert got == [bool(v) for v in js["isPowerOfTwo/out"]]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/extractors"), reason="reference checkout not mounted")
def test_vectors_are_reproducible_from_the_reference_sources(js, golden_audio):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_js_golden as gen
    it = gen.build()
    cases = [str(c).split("/") for c in js["cases"]]
    for ci in (2, 7):  # sound2 N=256 (the +-Infinity mfcc case) and the impulse
        clip, N, f, window = cases[ci]
        r = gen.run_frame(it, _signal(golden_audio, clip, int(N), int(f)), SR, window)
        assert _bits(r["mfcc"], js["%d/mfcc" % ci]) and _bits(r["amplitudeSpectrum"], js["%d/amplitudeSpectrum" % ci])
        assert _num(r["spectralKurtosis"], js["%d/spectralKurtosis" % ci])
        assert _bits(r["complexSpectrum"]["imag"], js["%d/complexSpectrum.imag" % ci])


def test_minijs_semantics():
    """The interpreter itself on the constructs the reference relies on."""
    from oracle.minijs import Interpreter
    it = Interpreter(ROOT)
    run = lambda src: it.run_source("return (function(){" + src + "})()")  # noqa: E731
    assert run("var a = 1\nvar b = 2\nreturn a + b") == 3.0  # ASI
    assert run("var x = 5; x >>= 1; x <<= 3; return (x & 24) + (x | 1)") == 33.0
    assert run("var f = {}; f[3] = f[1.5] = true; return f.hasOwnProperty(3) && f.hasOwnProperty('1.5') && !f.hasOwnProperty(2)")
    assert run("var a = new Float32Array(2); a[0] = 0.1; a[5] = 7; return a[0]") == float(np.float32(0.1))
    assert run("var z = Array.apply(null, new Array(4)).map(Number.prototype.valueOf, 0); return z.length + z[3]") == 4.0
    assert run("function P(){ if (!(this instanceof P)) return new P(); this.v = 4 } P.prototype.g = function(){ return this.v }; return P().g()") == 4.0
    assert run("var o = {v: 2, f: function(){ return this.v }}; var g = o.f.bind({v: 9}); return g() + o.f()") == 11.0
    assert run("return typeof nope === 'undefined' && typeof Math.pow === 'function'")
    assert np.isnan(run("return 0/0")) and run("return 1/0") == float("inf") and run("return Math.pow(0, 0.23)") == 0.0
    assert run("var i = 0, s = 0; while (i < 5) { i++; if (i == 2) continue; s += i } return s") == 13.0
    assert run("return (7 % 2 === 1) && (-7 % 2 === -1) && ('a' + 1 === 'a1')")


def _node_baseline_core():
    src = open(os.path.join(ROOT, "baseline", "node", "ref_worker.js"), encoding="utf-8").read()
    return src.split("// ---- BEGIN ES5 CORE")[1].split("// ---- END ES5 CORE")[0]


def test_node_baseline_core_is_plain_es5():
    """baseline/node/ref_worker.js cannot run here (no Node); its numeric core at least parses as the ES5 subset."""
    from oracle.minijs import Parser
    assert len(Parser(_node_baseline_core()).program()[1]) == 3  # makeReferencePath, framesOf, runClips


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/extractors"), reason="reference checkout not mounted")
def test_node_baseline_core_drives_the_reference_sources(js, golden_audio):
    """The per-frame sequence of the Node baseline, executed under minijs on the reference's own files, returns the
    golden vectors (so what that script times is the path the oracle restates), and frames clips like the C ABI."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_js_golden as gen
    from oracle.minijs import JSArray, JSObject, JSTypedArray
    it = gen.build()
    G = it.global_env.vars
    it.run_source(_node_baseline_core())
    env = JSObject(it.object_proto)
    for k in ["ComplexArray", "extractors"] + gen.METHODS:
        env.put(k, G[k])
    G["env"] = env
    G["audioContext"] = JSObject(it.object_proto)
    G["audioContext"].put("sampleRate", SR)
    G["names"] = JSArray(it, ["buffer"] + gen.EXTRACTORS + ["loudness"])
    ci = 2
    clip, N, f, window = [str(c).split("/") for c in js["cases"]][ci]
    N, f = int(N), int(f)
    G["clip"] = JSTypedArray(it, "Float32Array", np.asarray(golden_audio[clip][:(f + 1) * N], np.float32).copy())
    it.run_source("var P = makeReferencePath(env, %d, %r, '%s', names); var one = P.frame(clip.subarray(%d, %d));"
                  % (N, SR, window, f * N, (f + 1) * N))
    r = {k: gen.to_py(v) for k, v in G["one"].props.items()}  # copied now: the reference returns aliases of reused buffers
    it.run_source("var rc = runClips(P, [clip.subarray(0, 1024), clip.subarray(0, 700), clip.subarray(0, 100)], %d, 128);" % N)
    for k in NUMBERS:
        assert _num(r[k], js["%d/%s" % (ci, k)]), k
    for k in ("mfcc", "amplitudeSpectrum", "powerSpectrum", "buffer"):
        assert _bits(r[k], js["%d/%s" % (ci, k)]), k
    assert _bits(r["complexSpectrum"]["real"], js["%d/complexSpectrum.real" % ci])
    assert _bits(r["loudness"]["specific"], js["%d/loudness.specific" % ci])
    want = sum(mo.num_frames(n, N, 128) for n in (1024, 700, 100))
    assert gen.to_py(G["rc"])["frames"] == want == 11


def test_oracle_parameters_reproduce_the_reference_javascript(golden_audio):
    """numBarkBands through the reference's own NUM_BARK_BANDS option; filter / coefficient counts and the rolloff
    fraction through the reference's module text with that one constant replaced (tools/make_js_golden_params.py)."""
    js = np.load(os.path.join(ROOT, "tests", "golden", "js_reference_params.npz"))
    cases = [str(c).split("/") for c in js["cases"]]
    assert len(cases) == 4
    for ci, (clip, N, f, window, nb, edited) in enumerate(cases):
        N, f, nb = int(N), int(f), int(nb)
        params = {"numBarkBands": nb}
        if int(edited):
            params.update(numMelFilters=40, numMfccCoefficients=20, rolloffFraction=0.85)
        r = mo.extract(_signal(golden_audio, clip, N, f), N, N, SR, window, params=params)
        assert np.array_equal(mo.bark_band_limits(mo.bark_scale(N, SR), N // 2, nb), js["%d/bbLimits" % ci].astype(np.int32))
        assert r["loudness"]["specific"].shape == (1, nb) and r["mfcc"].shape == (1, 20 if int(edited) else 13)
        assert _bits(r["loudness"]["specific"][0], js["%d/loudness.specific" % ci]), (ci, "specific")
        assert _bits(r["mfcc"][0], js["%d/mfcc" % ci]), (ci, "mfcc")
        for k in ("perceptualSpread", "perceptualSharpness", "spectralRolloff", "spectralCentroid"):
            assert _num(r[k][0], js["%d/%s" % (ci, k)]), (ci, k, r[k][0], float(js["%d/%s" % (ci, k)]))
        assert _num(r["loudness"]["total"][0], js["%d/loudness.total" % ci])
    assert np.isnan(js["1/perceptualSharpness"])  # 12 bands: spec[12] is undefined (perceptualSharpness.js:8)
