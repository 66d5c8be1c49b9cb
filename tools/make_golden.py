#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.  Run in the dev container (needs /root/reference).

  audio_pcm16.npz   the int16 PCM payload of the reference's three demo clips
                    (audio/sound1-3.wav: "Vowels", "White Noise", "Sine Sweep",
                    index.html:41-49) -- parity INPUTS for BASELINE configs 1-2.
  golden_features.npz  the C oracle's outputs (oracle/meyda_oracle.c) on those
                    clips at bufferSize 256/512/1024/2048, hop = bufferSize,
                    hanning: every scalar feature + loudness.specific + mfcc for
                    every frame, and the array features for three frames each.

The reference has no golden vectors of its own (package.json:25).  These pin the
ORACLE against drift on the full demo clips; they are not outputs of the
reference itself.  The vectors that ARE the reference's own outputs (its .js
files executed by oracle/minijs.py) are made by tools/make_js_golden.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle, meyda_oracle as mo  # noqa: E402

REF = "/root/reference/audio"
OUT = os.path.join(ROOT, "tests", "golden")
SIZES = (256, 512, 1024, 2048)


def main():
    os.makedirs(OUT, exist_ok=True)
    pcm = {}
    for i in (1, 2, 3):
        x, sr = mo.read_wav_pcm16(f"{REF}/sound{i}.wav")
        assert sr == 44100.0
        pcm[f"sound{i}"] = np.round(x * 32768).astype(np.int16)
        assert np.array_equal(mo.pcm16_to_float(pcm[f"sound{i}"]), x)
    np.savez_compressed(os.path.join(OUT, "audio_pcm16.npz"), **pcm)

    g = {}
    for name, p in pcm.items():
        x = mo.pcm16_to_float(p)
        for N in SIZES:
            r = c_oracle.extract(x, N, N, 44100.0, "hanning")
            nf = len(r["rms"])
            key = f"{name}/{N}"
            g[f"{key}/scalars"] = np.stack(
                [r[k] if k != "loudness.total" else r["loudness"]["total"] for k in c_oracle.SCALAR_NAMES], axis=1)
            g[f"{key}/specific"] = r["loudness"]["specific"]
            g[f"{key}/mfcc"] = r["mfcc"]
            pick = np.array(sorted({0, nf // 2, nf - 1}))
            g[f"{key}/frames"] = pick
            g[f"{key}/amp"] = r["amplitudeSpectrum"][pick]
            g[f"{key}/power"] = r["powerSpectrum"][pick]
            g[f"{key}/real"] = r["complexSpectrum"]["real"][pick]
            g[f"{key}/imag"] = r["complexSpectrum"]["imag"][pick]
    g["scalar_names"] = np.array(c_oracle.SCALAR_NAMES)
    np.savez_compressed(os.path.join(OUT, "golden_features.npz"), **g)
    for f in os.listdir(OUT):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
