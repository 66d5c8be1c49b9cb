"""mb_extract_multi on REAL devices: one process, one plan and one host thread per GPU, clips sharded on frame count
(the call the N-API addon makes).  Needs at least two GPUs (`gpurun --gpus 2 ...`); skipped otherwise."""
import numpy as np
import pytest

import meyda_b200 as mb
from meyda_b200 import _capi
from meyda_b200.sharding import shard_by_frames
from oracle import meyda_oracle as mo

pytestmark = pytest.mark.gpu
SR = 44100.0


def _n_devices():
    import ctypes as C
    n = C.c_int(0)
    _capi.check(_capi.lib().mb_device_count(C.byref(n)))
    return n.value


@pytest.mark.parametrize("N,hop,feats", [(2048, 512, None), (512, 512, None),
                                         (2048, 512, ["mfcc", "spectralCentroid", "spectralKurtosis", "zcr"])])
def test_extract_multi_on_real_devices_matches_the_single_device_call(N, hop, feats):
    nd = min(_n_devices(), 4)
    if nd < 2:
        pytest.skip("needs at least two GPUs")
    feats = mb.FEATURES if feats is None else feats
    rng = np.random.default_rng(11)
    # ragged clips: empty, shorter than a frame, one frame, long, tonal (refined by the adaptive plans), noisy
    lens = [0, N - 1, N, 40 * N + 17, 3 * N, 25 * N + 1, 8 * N, 60 * N, N + hop, 12 * N]
    clips = [mo.synth_clip(20 + i, L) for i, L in enumerate(lens)]
    t = np.arange(25 * N + 1) / SR
    clips[5] = (0.5 * np.sin(2 * np.pi * 440.0 * t)).astype(np.float32)  # a pure tone: every frame goes to the exact pass
    data, off, ln = mb.meyda._normalize_clips(clips)
    single = mb.Plan(N, hop, SR, "hanning", feats, device=0)
    try:
        want, per = single.extract_host(data, off, ln)
        refined_single = single.refined_frames
    finally:
        single.close()
    plans = [mb.Plan(N, hop, SR, "hanning", feats, device=d) for d in range(nd)]
    try:
        got, per2 = mb.meyda.extract_multi(plans, data, off, ln)
        launches = [p.launch_count for p in plans]
        refined = sum(p.refined_frames for p in plans)
    finally:
        for p in plans:
            p.close()
    assert per.tolist() == per2.tolist() == [mo.num_frames(L, N, hop) for L in lens]
    for k in want:
        assert np.array_equal(got[k], want[k], equal_nan=True), k  # bit for bit, whichever device computed a frame
    assert refined == refined_single and refined >= mo.num_frames(25 * N + 1, N, hop) - 2
    # every device with a non-empty shard launched, and the shards are the frame-balanced contiguous ranges
    shards = shard_by_frames(per, nd)
    for d, (c0, c1) in enumerate(shards):
        assert (launches[d] > 0) == (int(per[c0:c1].sum()) > 0), (d, launches, shards)
    frames = [int(per[c0:c1].sum()) for c0, c1 in shards]
    assert max(frames) - min(frames) <= int(per.max()), (frames, "contiguous clip ranges balance to within one clip")
