"""RMS(|Z - Z_exact|)/peak of single synthetic frames (square wave, tone ...) for every FFT path."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import meyda_oracle as mo

for N in (512, 1024, 2048):
    for name in ("square", "tone", "dc", "impulse"):
        x = mo.degenerate_frame(name, N)
        w = (x.astype(np.float64) * mo.hanning(N).astype(np.float64)).astype(np.float32)
        Zx = np.conj(np.fft.fft(w.astype(np.float64))) / np.sqrt(N)
        rr, ri = mo.fft_jsfft(w[None, :])
        pk = np.abs(Zx).max()
        row = {"ref": np.sqrt((np.abs(rr[0].astype(np.float64) + 1j * ri[0] - Zx) ** 2).mean()) / pk}
        for nm, flags in (("fast", 0), ("generic", _capi.MB_FLAG_GENERIC_KERNEL)):
            p = mb.Plan(N, N, 44100.0, "hanning", ["complexSpectrum"], flags=flags)
            out, _ = p.extract_host(x, np.array([0], np.int64), np.array([N], np.int64))
            kn = p.kernel_name
            p.close()
            Zg = out["complex_real"][0].astype(np.float64) + 1j * out["complex_imag"][0]
            row[nm + ":" + kn] = np.sqrt((np.abs(Zg - Zx) ** 2).mean()) / pk
        print(N, name, {k: "%.3g" % v for k, v in row.items()}, flush=True)
