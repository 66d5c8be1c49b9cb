#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "other_buffer or multi_warp_frame or full_size" > gpurun_out/pytest_large.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_large.log
echo NEW; timeout 400 python tools/bench_configs.py large c5 2>&1 | grep config | tee gpurun_out/bench_large.log
echo GENERIC; timeout 400 python tools/bench_large_generic.py 2>&1 | grep config | tee gpurun_out/bench_large_generic.log
