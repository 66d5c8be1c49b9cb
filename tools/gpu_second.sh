#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/fft_noise.py > gpurun_out/fft_noise.log 2>&1; cat gpurun_out/fft_noise.log
timeout 900 python bench.py --clips 2000 --steps 2 --warmup 3 > gpurun_out/bench_2000.log 2>&1; echo "bench exit $?"
tail -2 gpurun_out/bench_2000.log
timeout 900 python bench.py --clips 2000 --steps 2 --warmup 3 --generic --no-e2e --no-cpu-baseline > gpurun_out/bench_2000_generic.log 2>&1; echo "bench-generic exit $?"
tail -1 gpurun_out/bench_2000_generic.log
