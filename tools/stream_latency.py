"""Per-buffer latency of the streaming path: one mb_stream_push per buffer, host memory
in and out, as `onaudioprocess` delivers buffers (src/meyda.js:69-91).  Wall clock around the blocking call."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import meyda_b200 as mb

SR = 44100.0
rng = np.random.default_rng(3)

def run(name, N, hop, feats, pushes=3000):
    plan = mb.Plan(N, hop, SR, "hanning", feats)
    st = mb.Stream(plan)
    out = plan.alloc_host_outputs(1)
    x = (rng.standard_normal(hop * (pushes + 8)) * 0.2).astype(np.float32)
    st.push_into(np.ascontiguousarray(x[:N - hop]), out) if N > hop else None  # prime the overlap tail
    t = np.zeros(pushes)
    for i in range(pushes):
        blk = x[i * hop:(i + 1) * hop]
        t0 = time.perf_counter()
        nf = st.push_into(blk, out)
        t[i] = time.perf_counter() - t0
        assert nf == 1
    t = t[200:] * 1e6
    print(json.dumps({"case": name, "bufferSize": N, "hop": hop, "features": len(feats), "kernel": plan.kernel_name,
                      "pushes": pushes, "graph_replays": st.graph_launches, "us_median": round(float(np.median(t)), 1),
                      "us_mean": round(float(t.mean()), 1), "us_p99": round(float(np.percentile(t, 99)), 1),
                      "buffer_period_us": round(hop / SR * 1e6, 1)}), flush=True)
    st.close(); plan.close()

run("reference usage: bufferSize 512, back-to-back, config-1 features", 512, 512,
    ["rms", "energy", "zcr", "amplitudeSpectrum", "spectralCentroid"])
run("bufferSize 512, back-to-back, all 18 features", 512, 512, mb.FEATURES)
run("bufferSize 2048, hop 512, all 18 features", 2048, 512, mb.FEATURES)
run("bufferSize 2048, hop 512, mfcc + moments", 2048, 512,
    ["mfcc", "spectralCentroid", "spectralSpread", "spectralSkewness", "spectralKurtosis"])
