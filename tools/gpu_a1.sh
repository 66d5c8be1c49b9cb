#!/bin/bash
# first GPU pass of round 2: adaptive exactness
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python tools/diag_adaptive.py > gpurun_out/diag_adaptive.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --clips 4000 --no-cpu-baseline > gpurun_out/bench_adaptive.json 2> gpurun_out/bench_adaptive.err
python bench.py --steps 3 --warmup 3 --clips 4000 --no-cpu-baseline --no-refine > gpurun_out/bench_norefine.json 2> gpurun_out/bench_norefine.err
tail -3 gpurun_out/diag_adaptive.log; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_adaptive.json gpurun_out/bench_norefine.json | cut -c1-600
