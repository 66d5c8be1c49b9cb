"""End-to-end rate of a host-memory call on material whose frames the exact kernel redoes (a pure tone: all of them;
white noise: a few per cent), host-rows modes 1 and 2 in turn.  usage (GPU box): PYTHONPATH=$PWD python tools/e2e_material.py"""
import time, numpy as np, torch, meyda_b200 as mb
sr = 44100; n = sr * 30
t = np.arange(n) / sr
clips = 12
tone = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)
rng = np.random.default_rng(1)
noise = (0.3 * rng.standard_normal(n)).astype(np.float32)
for name, sig in (("tone", tone), ("noise", noise)):
    x = torch.empty(clips, n, dtype=torch.float32, pin_memory=True); x.copy_(torch.from_numpy(np.tile(sig, (clips, 1))))
    hx = x.numpy().reshape(-1); off = np.arange(clips, dtype=np.int64) * n; ln = np.full(clips, n, np.int64)
    import os
    for mode in [int(m) for m in os.environ.get("MODES", "1,2,1,2").split(",")]:
        mb.set_host_rows(mode)
        plan = mb.Plan(2048, 512, 44100.0, "hanning", mb.FEATURES, device=0)
        nf = clips * ((n - 2048) // 512 + 1)
        ho = plan.alloc_host_outputs(nf, pinned=True)
        for _ in range(2): plan.extract_host(hx, off, ln, out=ho)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): plan.extract_host(hx, off, ln, out=ho)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print("%s mode %d: %.3f M frames/s, refined %d of %d" % (name, mode, nf / dt / 1e6, plan.refined_frames, nf), flush=True)
        plan.close()
