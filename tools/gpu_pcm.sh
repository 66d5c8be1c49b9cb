#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pcm16.py -m gpu -q --maxfail=20 -p no:cacheprovider > gpurun_out/pytest_pcm.log 2>&1; echo "pcm pytest exit $?"
tail -15 gpurun_out/pytest_pcm.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider --deselect tests/test_gpu_pcm16.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
bash tools/gpu_variants.sh
