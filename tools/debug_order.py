import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meyda_b200 as mb
from meyda_b200 import _capi
SR=44100.0
N, hop, n_clips, L = 2048, 512, 64, 441000
g = torch.Generator(device="cuda").manual_seed(1234)
x = (torch.rand(n_clips, L, device="cuda", generator=g) - 0.5) * 0.5
feats = ["rms", "energy", "zcr", "amplitudeSpectrum", "powerSpectrum", "loudness", "spectralCentroid", "mfcc"]
for flags in (0, _capi.MB_FLAG_GENERIC_KERNEL, _capi.MB_FLAG_EXACT_FFT):
    plan = mb.Plan(N, hop, SR, features=feats, flags=flags)
    nf_clip = (L - N)//hop + 1; nf = nf_clip*n_clips
    mk = lambda: {k: torch.zeros(s, dtype=torch.int32 if d == np.int32 else torch.float32, device="cuda") for k,(s,d) in plan.output_shapes(nf).items()}
    off = np.arange(n_clips, dtype=np.int64)*L; ln = np.full(n_clips, L, np.int64)
    runs=[]
    for rep in range(3):
        o = mk(); plan.extract_device(x.data_ptr(), x.numel(), off, ln, {k:v.data_ptr() for k,v in o.items()}); runs.append(o)
    o2 = mk(); plan.extract_device(x.data_ptr(), x.numel(), off[::-1].copy(), ln, {k:v.data_ptr() for k,v in o2.items()})
    for k in runs[0]:
        a = runs[0][k].reshape(n_clips, nf_clip, -1).float()
        d01 = (a - runs[1][k].reshape(n_clips, nf_clip, -1).float()).abs().nan_to_num().max().item()
        d02 = (a - runs[2][k].reshape(n_clips, nf_clip, -1).float()).abs().nan_to_num().max().item()
        b = o2[k].reshape(n_clips, nf_clip, -1).flip(0).float()
        dr = (a-b).abs().nan_to_num(); 
        print(plan.kernel_name, k, 'repeat diffs', d01, d02, 'reversed-order diff max', dr.max().item(), 'count', int((dr>0).sum().item()))
    plan.close()

print("---- locate")
plan = mb.Plan(N, hop, SR, features=feats, flags=_capi.MB_FLAG_GENERIC_KERNEL)
nf_clip = (L - N)//hop + 1; nf = nf_clip*n_clips
off = np.arange(n_clips, dtype=np.int64)*L; ln = np.full(n_clips, L, np.int64)
ref=None
for rep in range(8):
    o = {k: torch.zeros(s, dtype=torch.int32 if d == np.int32 else torch.float32, device="cuda") for k,(s,d) in plan.output_shapes(nf).items()}
    plan.extract_device(x.data_ptr(), x.numel(), off, ln, {k:v.data_ptr() for k,v in o.items()})
    if ref is None: ref=o; continue
    for k in o:
        d=(o[k].float()-ref[k].float()).abs().nan_to_num().reshape(nf,-1).max(1).values
        idx=torch.nonzero(d>0).flatten().tolist()
        if idx: print('rep',rep,k,'frames differing',idx[:10], 'vals', o[k].reshape(nf,-1)[idx[0]][:4].tolist(), ref[k].reshape(nf,-1)[idx[0]][:4].tolist())
