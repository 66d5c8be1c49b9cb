#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python tools/bench_configs.py small sizes exact > gpurun_out/bench_sizes.log 2>&1; cat gpurun_out/bench_sizes.log | tail -14
timeout 600 python tools/stream_latency.py > gpurun_out/stream_latency.json 2> gpurun_out/stream_latency.err; cat gpurun_out/stream_latency.json
