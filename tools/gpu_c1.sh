#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --clips 2000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
tail -5 gpurun_out/bench_small.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
print('value %.1fM frac %.3f parity %s banded %s refined %s launches %s' % (d['value']/1e6, d['roofline']['frac'], d['parity'], d['parity_banded'], d['refined_frames_last_wave'], d['gpu_launches']))
print('e2e', json.dumps(d['e2e'])[:900])
for s in d['secondary']: print(s['config']['workload'][:40], '%.1fM'%(s['value']/1e6), s['kernel'], json.dumps(s['roofline'])[:300], s['parity'], s['clocks'])
print('cpu', d['cpu_baseline'])
PY
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
python tools/bench_configs.py exactcmp > gpurun_out/bench_exactcmp.json 2>&1; cut -c1-160 gpurun_out/bench_exactcmp.json
