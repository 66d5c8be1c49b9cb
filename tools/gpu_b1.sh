#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "exact" > gpurun_out/pytest_exact.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_exact.log
tail -12 gpurun_out/pytest_exact.log
python tools/bench_configs.py exactcmp > gpurun_out/bench_exactcmp.json 2>&1
cat gpurun_out/bench_exactcmp.json | cut -c1-250
python tools/diag_adaptive.py > gpurun_out/diag_adaptive.log 2>&1; grep -c OK gpurun_out/diag_adaptive.log
