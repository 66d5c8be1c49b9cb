#!/bin/bash
# Quick checkpoint: smoke, every GPU test, then a longer seeded fuzz soak with another seed.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/pytest_gpu.log
MEYDA_FUZZ_CASES=${FUZZ_CASES:-150} MEYDA_FUZZ_SEED=${FUZZ_SEED:-777} timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_fuzz_soak.log 2>&1; echo "soak exit $?"
tail -6 gpurun_out/pytest_fuzz_soak.log
