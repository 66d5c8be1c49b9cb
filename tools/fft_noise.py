"""How noisy is each FFT relative to the mathematically exact transform?
Prints RMS(|Z - Z_exact|)/peak for the reference arithmetic (oracle jsfft), the
GPU float32 kernels and the GPU exact mode, on frames of the demo clips."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import meyda_oracle as mo

z = np.load(os.path.join(ROOT, "tests", "golden", "audio_pcm16.npz"))
for clip in ("sound1", "sound2", "sound3"):
    x = mo.pcm16_to_float(z[clip])[:2048 * 60]
    for N in (512, 2048):
        fr = mo.frame_signal(x, N, N)
        w = (fr.astype(np.float64) * mo.hanning(N).astype(np.float64)).astype(np.float32)
        Zx = np.conj(np.fft.fft(w.astype(np.float64), axis=1)) / np.sqrt(N)
        rr, ri = mo.fft_jsfft(w)
        Zr = rr.astype(np.float64) + 1j * ri
        pk = np.abs(Zx).max(axis=1, keepdims=True)
        row = {"ref": np.sqrt((np.abs(Zr - Zx) ** 2).mean(axis=1, keepdims=True)) / pk}
        for name, flags in (("gpu_fast", 0), ("gpu_generic", _capi.MB_FLAG_GENERIC_KERNEL), ("gpu_exact", _capi.MB_FLAG_EXACT_FFT)):
            p = mb.Plan(N, N, 44100.0, "hanning", ["complexSpectrum"], flags=flags)
            out, _ = p.extract_host(x, np.array([0], np.int64), np.array([len(x)], np.int64))
            kn = p.kernel_name
            p.close()
            Zg = out["complex_real"].astype(np.float64) + 1j * out["complex_imag"]
            row[name + ":" + kn] = np.sqrt((np.abs(Zg - Zx) ** 2).mean(axis=1, keepdims=True)) / pk
        print(clip, N, {k: "%.3g" % float(np.median(v)) for k, v in row.items()})
