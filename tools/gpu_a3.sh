#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
CLIPS=2400 bash tools/gpu_variants.sh > gpurun_out/variants.log 2>&1
python tools/diag_adaptive.py > gpurun_out/diag_adaptive.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
cat gpurun_out/variants.log; grep -c OK gpurun_out/diag_adaptive.log; grep refined gpurun_out/diag_adaptive.log | awk '{print $1,$2,$4}' | tr '\n' ';'; tail -4 gpurun_out/pytest_gpu.log
