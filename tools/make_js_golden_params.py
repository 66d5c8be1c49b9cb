#!/usr/bin/env python
"""Golden vectors for the parameters the reference keeps as constants (mb_plan_create_ex), produced by the
reference's own JavaScript under oracle/minijs.py like tools/make_js_golden.py:

  * NUM_BARK_BANDS is an option of the reference's Loudness constructor (src/extractors/loudness.js:14): the
    unmodified sources are run with 30 and with 12 bands (12: perceptualSharpness reads spec[i + 1] past the
    end and returns NaN);
  * numFilters / numCoeffs (src/extractors/mfcc.js:15,71) and the 0.99 of spectralRolloff.js:9 are local
    constants: the module text is loaded with exactly that constant replaced (40 filters, 20 coefficients,
    fraction 0.85), everything else verbatim.

Writes tests/golden/js_reference_params.npz; tests/test_js_pin.py pins the numpy oracle's `params` against it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_js_golden as gen  # noqa: E402
from oracle import meyda_oracle as mo  # noqa: E402

EDITS = {"mfcc.js": [("var numFilters = 26;", "var numFilters = 40;"), ("var numCoeffs = 13;", "var numCoeffs = 20;")],
         "spectralRolloff.js": [("0.99 * ec", "0.85 * ec")]}
# (clip, N, frame, window, NUM_BARK_BANDS, edited constants?)
CASES = [("sound1", 512, 0, "hanning", 30, False), ("sound3", 512, 300, "hanning", 12, False),
         ("sound2", 1024, 10, "hamming", 24, True), ("sound1", 256, 100, "hanning", 40, True)]
KEEP = ["loudness", "perceptualSpread", "perceptualSharpness", "mfcc", "spectralRolloff", "bbLimits", "spectralCentroid"]


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "audio_pcm16.npz"))
    clips = {k: mo.pcm16_to_float(z[k]) for k in z.files}
    its = {False: gen.build(), True: gen.build(EDITS)}
    out = {"cases": np.array(["%s/%d/%d/%s/%d/%d" % c for c in CASES])}
    for ci, (clip, N, f, window, nb, edited) in enumerate(CASES):
        r = gen.run_frame(its[edited], clips[clip][f * N:(f + 1) * N], 44100.0, window, nb)
        for k in KEEP:
            v = r[k]
            if isinstance(v, dict):
                for s, a in v.items():
                    out["%d/%s.%s" % (ci, k, s)] = np.asarray(a)
            else:
                out["%d/%s" % (ci, k)] = np.asarray(v)
        print(ci, clip, N, nb, edited, "sharpness", r["perceptualSharpness"], "rolloff", r["spectralRolloff"],
              "mfcc", len(r["mfcc"]), flush=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "js_reference_params.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
