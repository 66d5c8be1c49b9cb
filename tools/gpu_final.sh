#!/bin/bash
# End-of-round checkpoint: gpu_full.sh (smoke, every GPU test, both bench arms, ncu launch list) + the secondary configurations.
bash tools/gpu_full.sh
timeout 900 python tools/bench_configs.py c3 c5 c1 small pcm large > gpurun_out/bench_configs_final.log 2>&1; echo "configs exit $?"
grep config gpurun_out/bench_configs_final.log | sed -E 's/"alg_bytes.*//; s/"N":.*"frames_per_s"/fps/' | cut -c1-160
