#!/bin/bash
# Full GPU test pass on the working-tree library, a parity subset on every tuning variant, then the A/B bench.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -5 gpurun_out/pytest_gpu.log
for so in $(ls meyda_b200/_lib/variants/lib_MB_*.so 2>/dev/null); do
  MEYDA_B200_LIB=$PWD/$so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider \
    -k "fast and (2048 or hop_reuse or ragged or hamming) or full_size_properties or stay_on_the_warp" > gpurun_out/pytest_$(basename $so .so).log 2>&1
  echo "$so subset exit $?"; tail -2 gpurun_out/pytest_$(basename $so .so).log
done
bash tools/gpu_variants.sh
