"""Which CTA size should the generic block-per-frame kernel use at each bufferSize?  (MB_GENERIC_CTA override.)"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb
from meyda_b200 import _capi
SR = 44100.0
dev = torch.device("cuda", 0)

def run(N, hop, feats, flags, cta, n_clips=300, clip_len=441000):
    os.environ["MB_GENERIC_CTA"] = str(cta)
    g = torch.Generator(device=dev).manual_seed(7)
    x = (torch.rand(n_clips, clip_len, device=dev, generator=g) - 0.5) * 0.5
    plan = mb.Plan(N, hop, SR, "hanning", feats, flags=flags | _capi.MB_FLAG_GENERIC_KERNEL)
    nf = ((clip_len - N) // hop + 1) * n_clips
    outs = {k: torch.empty(s, dtype=torch.int32 if d == np.int32 else torch.float32, device=dev)
            for k, (s, d) in plan.output_shapes(nf).items()}
    st = torch.cuda.Stream(device=dev); torch.cuda.set_stream(st); plan.set_stream(st.cuda_stream)
    off = np.arange(n_clips, dtype=np.int64) * clip_len; ln = np.full(n_clips, clip_len, np.int64)
    ptrs = {k: v.data_ptr() for k, v in outs.items()}
    for _ in range(2): plan.extract_device(x.data_ptr(), x.numel(), off, ln, ptrs, sync=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(2): plan.extract_device(x.data_ptr(), x.numel(), off, ln, ptrs, sync=False)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    plan.close(); del x, outs; torch.cuda.empty_cache()
    return nf / (ms * 1e-3) / 1e6

C1 = ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"]
res = {}
for label, feats, flags in (("full", mb.FEATURES, 0), ("config1", C1, 0), ("exact", mb.FEATURES, _capi.MB_FLAG_EXACT_FFT)):
    for N in ((8192, 16384, 32768) if "big" in sys.argv else (64, 256, 512, 1024, 2048, 4096)):
        row = {}
        for cta in ((128, 256, 512, 1024) if "big" in sys.argv else (32, 64, 128, 256, 512)):
            if cta > N or (N >= 2048 and cta < 64) or (N <= 256 and cta > 256): continue
            try:
                row[cta] = round(run(N, N, feats, flags, cta, n_clips=(150 if label == "exact" else 300) // (4 if "big" in sys.argv else 1)), 1)
            except Exception as e:
                row[cta] = str(e)[:60]
        res["%s N=%d" % (label, N)] = row
        print(label, N, row, flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep_generic_cta.json"), "w"), indent=1)
