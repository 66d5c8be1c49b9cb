#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nproc
python bench.py --steps 5 --warmup 3 --clips 1200 --no-cpu-baseline --no-secondary > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_small.json').read().strip().splitlines()[-1])
print('value %.1fM frac %.3f' % (d['value']/1e6, d['roofline']['frac']))
e=d['e2e']; print('e2e %.3fM ceiling %.3fM frac %.3f pageable %.3fM d2h %d' % (e['value']/1e6, e['pcie_ceiling']['value']/1e6, e['frac_of_pcie_ceiling'], e['pageable']['value']/1e6, e['d2h_bytes_per_step']))
PY
python -m pytest tests -m gpu -q -x -k "host or device_memory or config1 or ragged or pcm" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
