"""Seeded random batches through the tuned kernels: ragged clips, arbitrary hops (aligned or not), frame counts that
are not multiples of a warp's chunk or group.  Frame bookkeeping, every feature against the oracle at the flat
tolerances (no noise band), and a batch equals its clips extracted one by one, bit for bit (a frame's bits depend on
nothing but its samples and the feature list)."""
import numpy as np
import pytest

import meyda_b200 as mb
from oracle import c_oracle, meyda_oracle as mo
from tests import parity
from tests.test_gpu_parity import SR

pytestmark = pytest.mark.gpu


def _cases():
    import os
    rng = np.random.default_rng(int(os.environ.get("MEYDA_FUZZ_SEED", "20261018")))
    out = []
    for i in range(int(os.environ.get("MEYDA_FUZZ_CASES", "36"))):  # (a longer soak: MEYDA_FUZZ_CASES=400 MEYDA_FUZZ_SEED=...)
        N = int(rng.choice([256, 512, 1024, 2048]))
        hop = int(rng.choice([N, N // 2, N // 4, int(rng.integers(1, 2 * N)), 4 * int(rng.integers(1, N // 2))]))
        lens = [int(rng.integers(0, 6 * N)) for _ in range(int(rng.integers(1, 7)))]
        out.append(pytest.param(N, hop, lens, i, id="%d-N%d-hop%d-%dclips" % (i, N, hop, len(lens))))
    return out


@pytest.mark.parametrize("N,hop,lens,seed", _cases())
def test_random_batches(N, hop, lens, seed):
    clips = [mo.synth_clip(1000 + 17 * seed + j, L) if L else np.zeros(0, np.float32) for j, L in enumerate(lens)]
    data, off, ln = mb.meyda._normalize_clips(clips)
    plan = mb.Plan(N, hop, SR)
    try:
        out, per = plan.extract_host(data, off, ln)
        assert plan.kernel_name in ("warp2048", "warpmf256", "warpmf512", "warpmf1024")
        assert per.tolist() == [mo.num_frames(L, N, hop) for L in lens]
        row = 0
        for c, L in zip(clips, lens):  # the batch equals its clips one by one
            nf = mo.num_frames(L, N, hop)
            if nf == 0:
                continue
            one, _ = plan.extract_host(c, np.zeros(1, np.int64), np.array([L], np.int64))
            for k in out:
                assert np.array_equal(one[k], out[k][row:row + nf], equal_nan=True), (k, row)
            row += nf
    finally:
        plan.close()
    kept = [c for c in clips if len(c) >= N]
    if not kept:
        assert all(len(v) == 0 for v in out.values())
        return
    ref = mo._concat([c_oracle.extract(c, N, hop, SR, "hanning") for c in kept])
    # every feature, spectralSkewness / spectralKurtosis included, at BASELINE.json's flat tolerances and with NO noise
    # band: the plans are adaptive (frames whose values the reference's own rounding decides are redone exactly)
    assert parity.compare_all(out, ref, N, noise_band=None) == {}
