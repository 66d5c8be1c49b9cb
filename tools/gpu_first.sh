#!/bin/bash
# First GPU pass: smoke, parity tests, microbench, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 120 tools/_bin/microbench > gpurun_out/microbench.json 2>&1; cat gpurun_out/microbench.json
timeout 900 python bench.py --clips 2000 --steps 2 --warmup 3 > gpurun_out/bench_2000.log 2>&1; echo "bench exit $?"
tail -3 gpurun_out/bench_2000.log
