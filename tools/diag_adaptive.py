"""GPU diagnostic: refined-frame counts and noise-band-free parity for config 2 (tools use only)."""
import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import meyda_b200 as mb
from meyda_b200 import _capi
from oracle import c_oracle, meyda_oracle as mo
from tests import parity
g = np.load('/root/repo/tests/golden/audio_pcm16.npz')
SR = 44100.0
for clip in ['sound1', 'sound2', 'sound3']:
    x = mo.pcm16_to_float(g[clip])
    for N in [256, 512, 1024, 2048]:
        data, off, ln = mb.meyda._normalize_clips(x)
        plan = mb.Plan(N, N, SR, 'hanning', mb.FEATURES)
        t0 = time.time(); out, per = plan.extract_host(data, off, ln); dt = time.time() - t0
        refined = plan.refined_frames; name = plan.kernel_name
        plan.close()
        ref = c_oracle.extract(x, N, N, SR)
        try:
            banded = parity.compare_all(out, ref, N, noise_band=None)
            res = 'OK'
        except AssertionError as e:
            res = 'FAIL ' + str(e)[:300]
        print('%s N=%d frames=%d refined=%d kernel=%s %s' % (clip, N, per[0], refined, name, res), flush=True)
