#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "streaming or meyda_class or callback" > gpurun_out/pytest_stream.log 2>&1; echo "stream pytest exit $?"
tail -15 gpurun_out/pytest_stream.log
timeout 600 python tools/stream_latency.py > gpurun_out/stream_latency.json 2> gpurun_out/stream_latency.err; echo "latency exit $?"; cat gpurun_out/stream_latency.json; tail -3 gpurun_out/stream_latency.err
timeout 900 python tools/bench_configs.py pcm > gpurun_out/bench_pcm.log 2>&1; echo "bench pcm exit $?"; cat gpurun_out/bench_pcm.log | tail -5
