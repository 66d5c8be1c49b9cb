#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py small > gpurun_out/bench_small.log 2>&1; cat gpurun_out/bench_small.log | tail -6
