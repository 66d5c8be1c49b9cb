#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -40; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --clips 2000 --no-cpu-baseline --no-e2e > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
print('value %.1fM frac %.3f parity %s banded %s' % (d['value']/1e6, d['roofline']['frac'], d['parity'][:30], d['parity_banded']))
for s in d['secondary']: print(s['config']['bufferSize'], '%.2fM'%(s['value']/1e6), s['kernel'], 'frac %.3f'%s['roofline']['frac'], s['parity'][:40], s['gpu_launches'])
PY
