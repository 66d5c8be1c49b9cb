"""Device-resident throughput of the other BASELINE configs and modes (not the
driver's bench): config 3 (mfcc + moments, N=2048), config 5 (N=32768 spectrum +
rolloff/flatness/slope), exact-FFT mode, other buffer sizes, 16-bit PCM input."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb
from meyda_b200 import _capi

SR = 44100.0
dev = torch.device("cuda", 0)
PER = {"buffer": lambda N: N, "complexSpectrum": lambda N: 2 * N, "amplitudeSpectrum": lambda N: N // 2,
       "powerSpectrum": lambda N: N // 2, "loudness": lambda N: 25, "mfcc": lambda N: 13}

C3 = ["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"]
C5 = ["amplitudeSpectrum", "spectralRolloff", "spectralFlatness", "spectralSlope"]

def run(name, N, hop, n_clips, clip_len, feats, flags=0, steps=3, pcm16=False):
    g = torch.Generator(device=dev).manual_seed(7)
    x = (torch.rand(n_clips, clip_len, device=dev, generator=g) - 0.5) * 0.5
    t = torch.arange(clip_len, device=dev, dtype=torch.float32) / SR
    x += 0.3 * torch.sin(2 * np.pi * 440.0 * t)[None, :]
    if pcm16:  # the same clips as 16-bit PCM (what a WAV file holds): half the input bytes
        x = (x * 32767.0).round().to(torch.int16)
    plan = mb.Plan(N, hop, SR, "hanning", feats, flags=flags)
    fpc = (clip_len - N) // hop + 1
    nf = fpc * n_clips
    outs = {k: torch.empty(s, dtype=torch.int32 if d == np.int32 else torch.float32, device=dev)
            for k, (s, d) in plan.output_shapes(nf).items()}
    st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); plan.set_stream(st.cuda_stream)
    off = np.arange(n_clips, dtype=np.int64) * clip_len; ln = np.full(n_clips, clip_len, np.int64)
    ptrs = {k: v.data_ptr() for k, v in outs.items()}
    def call():
        if pcm16: plan.extract_pcm16_device(x.data_ptr(), x.numel(), 1, 0, off, ln, ptrs, sync=False)
        else: plan.extract_device(x.data_ptr(), x.numel(), off, ln, ptrs, sync=False)
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps): call()
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    bpf = (2 if pcm16 else 4) * hop + 4 * sum(PER.get(f, lambda N: 1)(N) for f in feats)
    fps = nf / (ms * 1e-3)
    print(json.dumps({"config": name, "kernel": plan.kernel_name, "N": N, "hop": hop, "frames": nf, "refined": plan.refined_frames, "ms": round(ms, 3),
                      "frames_per_s": round(fps), "alg_bytes_per_frame": bpf, "alg_GBps": round(fps * bpf / 1e9, 1),
                      "hbm_frac_of_measured": round(fps * bpf / 1e9 / 6542.7, 4)}), flush=True)
    plan.close(); del x, outs; torch.cuda.empty_cache()

which = sys.argv[1:] or ["c3", "c5", "exact", "sizes", "pcm"]
if "pcm" in which:
    run("full set, float32 input", 2048, 512, 1500, 441000, mb.FEATURES)
    run("full set, 16-bit PCM input", 2048, 512, 1500, 441000, mb.FEATURES, pcm16=True)
    run("config3, 16-bit PCM input", 2048, 512, 4000, 441000, C3, pcm16=True)
if "c3" in which:
    run("config3 mfcc+moments", 2048, 512, 4000, 441000, C3)
    run("config3 generic", 2048, 512, 1000, 441000, C3, flags=_capi.MB_FLAG_GENERIC_KERNEL)
if "c5" in which:
    run("config5 N=32768 hop=8192", 32768, 8192, 256, 2646000, C5)
if "c5ab" in which:  # the adaptive statistics' cost in the multi-warp-per-frame kernel
    run("config5 N=32768 hop=8192", 32768, 8192, 256, 2646000, C5)
    run("config5 N=32768 hop=8192, MB_FLAG_NO_REFINE", 32768, 8192, 256, 2646000, C5, flags=_capi.MB_FLAG_NO_REFINE)
    run("config5 features N=8192, adaptive", 8192, 2048, 256, 2646000, C5)
    run("config5 features N=8192, MB_FLAG_NO_REFINE", 8192, 2048, 256, 2646000, C5, flags=_capi.MB_FLAG_NO_REFINE)
if "largeab" in which:  # the same, full set, at every large size
    for N in (4096, 8192, 16384, 32768):
        run("full set N=%d hop=N/4, adaptive" % N, N, N // 4, 64, 2646000, mb.FEATURES)
        run("full set N=%d hop=N/4, MB_FLAG_NO_REFINE" % N, N, N // 4, 64, 2646000, mb.FEATURES, flags=_capi.MB_FLAG_NO_REFINE)
if "bigab" in which:  # (short form of largeab + c5ab for variant runs)
    for fl, nm in ((0, "adaptive"), (_capi.MB_FLAG_NO_REFINE, "MB_FLAG_NO_REFINE")):
        run("full set N=4096 hop=N/4, " + nm, 4096, 1024, 64, 2646000, mb.FEATURES, flags=fl)
        run("full set N=8192 hop=N/4, " + nm, 8192, 2048, 64, 2646000, mb.FEATURES, flags=fl)
        run("config5 N=32768 hop=8192, " + nm, 32768, 8192, 256, 2646000, C5, flags=fl)
if "exact" in which:
    run("full set exact-FFT", 2048, 512, 200, 441000, mb.FEATURES, flags=_capi.MB_FLAG_EXACT_FFT)
    run("config3 exact-FFT", 2048, 512, 200, 441000, C3, flags=_capi.MB_FLAG_EXACT_FFT)
if "exactone" in which:  # (profiling target)
    run("full set exact-FFT (warp per frame) N=2048", 2048, 512, 400, 441000, mb.FEATURES, flags=_capi.MB_FLAG_EXACT_FFT, steps=1)
if "exactcmp" in which:  # the warp-per-frame exact kernel against the block-per-frame one
    EX, GEN = _capi.MB_FLAG_EXACT_FFT, _capi.MB_FLAG_GENERIC_KERNEL
    for N, hop in ((2048, 512), (1024, 1024), (512, 512)):
        run("full set exact-FFT (warp per frame) N=%d" % N, N, hop, 400, 441000, mb.FEATURES, flags=EX)
        run("full set exact-FFT (block per frame) N=%d" % N, N, hop, 200, 441000, mb.FEATURES, flags=EX | GEN)
    run("config3 exact-FFT (warp per frame)", 2048, 512, 400, 441000, C3, flags=EX)
    run("config3 exact-FFT (block per frame)", 2048, 512, 200, 441000, C3, flags=EX | GEN)
if "small" in which:  # the reference's own cadence: back-to-back buffers (hop = bufferSize)
    for N in (256, 512, 1024):
        run("full set N=%d hop=N" % N, N, N, 800, 441000, mb.FEATURES)
    run("config-1 features N=512 hop=N", 512, 512, 2000, 441000, ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"])
    run("full set exact-FFT N=512 hop=N", 512, 512, 200, 441000, mb.FEATURES, flags=_capi.MB_FLAG_EXACT_FFT)
    run("config-3 features (mfcc + moments) N=512 hop=N", 512, 512, 3000, 441000, C3)
    run("config-3 features (mfcc + moments) N=1024 hop=N", 1024, 1024, 3000, 441000, C3)
    run("full set N=512 hop=N, 3000 clips", 512, 512, 3000, 441000, mb.FEATURES)
    run("full set N=1024 hop=N, 3000 clips", 1024, 1024, 3000, 441000, mb.FEATURES)
if "smallab" in which:  # what the adaptive second pass costs at the reference's own sizes
    for N in (256, 512, 1024):
        run("full set N=%d hop=N, adaptive" % N, N, N, 800, 441000, mb.FEATURES)
        run("full set N=%d hop=N, MB_FLAG_NO_REFINE" % N, N, N, 800, 441000, mb.FEATURES, flags=_capi.MB_FLAG_NO_REFINE)
if "c1" in which:  # BASELINE config 1's feature list at the reference's default bufferSize
    run("config-1 features N=512 hop=N, 3000 clips", 512, 512, 3000, 441000, ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"])
    run("config-1 features N=512 hop=N/4", 512, 128, 800, 441000, ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"])
if "large" in which:
    for N in (4096, 8192, 16384, 32768):
        run("full set N=%d hop=N/4" % N, N, N // 4, 64, 2646000, mb.FEATURES)
        run("config-5 features N=%d hop=N/4" % N, N, N // 4, 256, 2646000, C5)
if "sizes" in which:
    for N in (256, 512, 1024, 4096):
        run("full set N=%d hop=N/4" % N, N, N // 4, 400, 441000, mb.FEATURES)
