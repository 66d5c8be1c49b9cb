#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "exact" > gpurun_out/pytest_exact.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_exact.log
tail -3 gpurun_out/pytest_exact.log
python tools/bench_configs.py exactcmp > gpurun_out/bench_exactcmp.json 2>&1
cut -c1-200 gpurun_out/bench_exactcmp.json
