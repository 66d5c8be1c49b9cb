#!/usr/bin/env python
"""Runs the reference's OWN JavaScript (unmodified files under /root/reference:
lib/jsfft/complex_array.js, lib/jsfft/fft.js, src/utils.js, every
src/extractors/*.js, and the compute* method bodies of src/meyda.js) under
oracle/minijs.py on a handful of frames, and commits what they return as
tests/golden/js_reference_vectors.npz.  tests/test_js_pin.py pins both oracle
restatements against these vectors.

No JavaScript engine exists in this image; minijs is a ~900-line ES5-subset
interpreter written for this purpose (float64 Numbers, Float32Array rounding,
prototype chains, ASI).  The driver below is the only glue: it performs the
intended per-buffer sequence of src/meyda.js:69-91 with the wiring fixes of
SURVEY.md 2.3 (a fresh zero-imaginary ComplexArray transformed per frame;
perceptual* reaching loudness through m.featureExtractors.loudness; mfcc's free
`audioContext`; the free global `µ`).  Everything numeric is the reference's code.
"""
import os
import re
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import meyda_oracle as mo  # noqa: E402  (only for the PCM fixture reader)
from oracle.minijs import (Interpreter, JSArray, JSObject, JSTypedArray, Parser, undefined)  # noqa: E402

REF = "/root/reference"
EXTRACTORS = ["rms", "energy", "zcr", "complexSpectrum", "amplitudeSpectrum", "powerSpectrum", "spectralCentroid",
              "spectralFlatness", "spectralSlope", "spectralRolloff", "spectralSpread", "spectralSkewness",
              "spectralKurtosis", "perceptualSpread", "perceptualSharpness", "mfcc"]
METHODS = ["computeAmplitude", "computeHamming", "computeHanning", "computeWindow", "computeBarkScale"]

DRIVER = r"""
var m = {signal: signal, audioContext: {sampleRate: SR}, featureExtractors: {}};
m.barkScale = computeBarkScale.call(m, N, SR);
m.hanning = computeHanning.call(m, N);
m.hamming = computeHamming.call(m, N);
var windowedSignal = computeWindow.call(m, signal, WINDOW);
var data = new ComplexArray(N);
data.map(function(value, i, n) { value.real = windowedSignal[i]; });
var spec = data.FFT();
m.complexSpectrum = spec;
m.ampSpectrum = new Float32Array(N / 2);
computeAmplitude.call(m, spec, m.ampSpectrum, N);
var loud = extractors.loudness({NUM_BARK_BANDS: NBARK, barkScale: m.barkScale, normalisedSpectrum: m.ampSpectrum, sampleRate: SR});
m.featureExtractors.loudness = function(bufferSize, mm) { return loud.process(); };
var results = {};
for (var x = 0; x < names.length; x++) { results[names[x]] = extractors[names[x]](N, m); }
results.loudness = loud.process();
results.buffer = m.signal;
results.hanning = m.hanning; results.hamming = m.hamming; results.barkScale = m.barkScale; results.bbLimits = loud.bbLimits;
"""


def lift_methods(interp):
    """The compute* methods of the ES6 class in src/meyda.js as plain functions (bodies verbatim)."""
    src = open(os.path.join(REF, "src/meyda.js"), encoding="utf-8").read()
    out = {}
    for name in METHODS:
        mt = re.search(r"\n\t+%s\(([^)]*)\)\s*\{" % name, src)
        i = depth = mt.end()
        depth = 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[i], 0)
            i += 1
        text = "(function(%s) {%s)" % (mt.group(1), src[mt.end():i])
        ast = Parser(text).program()
        out[name] = interp.eval(ast[1][0][1], interp.global_env, undefined)
    return out


def to_py(v):
    if isinstance(v, JSTypedArray):
        return v.data.copy()
    if isinstance(v, JSArray):
        return np.array([to_py(x) for x in v.items])
    if isinstance(v, JSObject):
        if "real" in v.props and "imag" in v.props:
            return {"real": to_py(v.props["real"]), "imag": to_py(v.props["imag"])}
        return {k: to_py(x) for k, x in v.props.items() if isinstance(x, (float, JSTypedArray))}
    return v


def build(source_edits=None):
    """source_edits: {file name: [(old, new)]} -- used only by the parameter vectors (make_js_golden_params.py),
    to replace the local constants numFilters / numCoeffs / 0.99 of mfcc.js and spectralRolloff.js."""
    it = Interpreter(REF)
    it.source_edits = source_edits or {}
    ca = it.require("lib/jsfft/complex_array")
    it.require("lib/jsfft/fft")  # decorates ComplexArray.prototype with FFT
    utils = it.require("src/utils")
    G = it.global_env.vars
    G["ComplexArray"] = ca.get("ComplexArray")
    G["µ"] = utils.get("µ")  # free global in spectralCentroid/Spread/Skewness/Kurtosis.js
    ex = JSObject(it.object_proto)
    for n in EXTRACTORS + ["loudness"]:
        ex.put(n, it.require("src/extractors/" + n))
    G["extractors"] = ex
    G.update(lift_methods(it))
    G["isPowerOfTwo"] = utils.get("isPowerOfTwo")
    return it


def run_frame(it, signal, sr, window, num_bark_bands=24):
    G = it.global_env.vars
    N = len(signal)
    G["signal"] = JSTypedArray(it, "Float32Array", np.asarray(signal, dtype=np.float32).copy())
    G["N"], G["SR"], G["WINDOW"], G["NBARK"] = float(N), float(sr), window, float(num_bark_bands)
    G["audioContext"] = JSObject(it.object_proto)  # free global in mfcc.js:20,37
    G["audioContext"].put("sampleRate", float(sr))
    G["names"] = JSArray(it, EXTRACTORS)
    it.run_source(DRIVER)
    res = G["results"]
    return {k: to_py(v) for k, v in res.props.items()}


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "audio_pcm16.npz"))
    clips = {k: mo.pcm16_to_float(z[k]) for k in z.files}
    cases = [("sound1", 512, 0, "hanning"), ("sound1", 512, 100, "hanning"), ("sound2", 256, 5, "hanning"),
             ("sound3", 1024, 200, "hanning"), ("sound2", 512, 40, "hamming"), ("sound1", 2048, 10, "hanning"),
             ("silence", 256, 0, "hanning"), ("impulse", 256, 0, "hanning"), ("square", 256, 0, "hamming"),
             # second batch: the special-value frames (NaN compares, -0, float32 underflow), a bin-centred tone, DC,
             # and more of the demo clips at other sizes / windows
             ("nan", 256, 0, "hanning"), ("negzero", 256, 0, "hanning"), ("tiny", 256, 0, "hanning"),
             ("tone", 512, 0, "hanning"), ("dc", 256, 0, "hamming"), ("sound2", 1024, 100, "hanning"),
             ("sound3", 2048, 50, "hamming"), ("sound1", 256, 300, "hamming")]
    for name, N, _f, _w in cases:
        if name not in clips:
            clips[name] = mo.degenerate_frame(name, N)
    it = build()
    out = {"cases": np.array(["%s/%d/%d/%s" % c for c in cases])}
    for ci, (clip, N, f, window) in enumerate(cases):
        t0 = time.time()
        sig = clips[clip][f * N:(f + 1) * N]
        r = run_frame(it, sig, 44100.0, window)
        for k, v in r.items():
            if isinstance(v, dict):
                for s, a in v.items():
                    out["%d/%s.%s" % (ci, k, s)] = np.asarray(a)
            else:
                out["%d/%s" % (ci, k)] = np.asarray(v)
        print("case %d %s N=%d frame %d %s: %.1f s  rms=%.10g centroid=%.10g mfcc0=%.8g" % (
            ci, clip, N, f, window, time.time() - t0, r["rms"], r["spectralCentroid"], r["mfcc"][0]), flush=True)
    # the power-of-two gate of src/utils.js:13-19
    ipo = it.global_env.vars["isPowerOfTwo"]
    out["isPowerOfTwo/in"] = np.array([1, 2, 3, 256, 600, 0, 32768], np.float64)
    out["isPowerOfTwo/out"] = np.array([bool(ipo.call(undefined, [float(v)])) for v in out["isPowerOfTwo/in"]])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "js_reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
