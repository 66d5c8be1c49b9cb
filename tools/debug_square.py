"""Is each kernel's mfcc / loudness consistent with its own amplitude spectrum?  (debug aid)"""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import meyda_oracle as mo
rng = np.random.default_rng(1)
for N in (512, 1024):
    fb = mo.mel_filterbank(N, 44100.0)
    dct = mo.dct_matrix().reshape(26, 13).T.astype(np.float64)
    bb = mo.bark_band_limits(mo.bark_scale(N, 44100.0), N // 2)
    for name in ("square", "noise", "tone"):
        x = mo.degenerate_frame(name, N) if name != "noise" else (rng.standard_normal(N) * 0.1).astype(np.float32)
        x4 = np.concatenate([x, x[::-1].copy(), (x * 0.5).astype(np.float32), x])  # four frames: does the position in the group matter?
        for nm, flags in (("fast", 0), ("generic", _capi.MB_FLAG_GENERIC_KERNEL)):
            p = mb.Plan(N, N, 44100.0, "hanning", ["amplitudeSpectrum", "mfcc", "loudness"], flags=flags)
            out, _ = p.extract_host(x4, np.array([0], np.int64), np.array([4 * N], np.int64))
            kn = p.kernel_name
            p.close()
            for fr in (0, 3):
                amp = out["amplitude_spectrum"][fr]
                pw = (amp ** 2).astype(np.float32).astype(np.float64)
                lg = np.log(fb[:, :N // 2] @ pw)
                mf = dct @ lg / 13
                spec = np.array([amp[bb[b]:bb[b + 1]].astype(np.float64).sum() ** 0.23 for b in range(24)])
                print(N, name, kn, "frame", fr, "max |mfcc - mfcc(own amp)| %.2e" % np.abs(out["mfcc"][fr] - mf).max(),
                      "max |specific - own| %.2e" % np.abs(out["loudness_specific"][fr] - spec).max(), flush=True)
