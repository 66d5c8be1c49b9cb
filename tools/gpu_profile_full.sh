#!/bin/bash
# One ncu full capture of the dominant kernel (after the same command ran clean without ncu).
mkdir -p gpurun_out
CMD="python bench.py --clips 600 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $BENCH_EXTRA"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:${KERNEL_RE:-mb_warp2048} -s 3 -c 1 -f -o gpurun_out/prof_${PROF_NAME:-warp2048} $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -2 gpurun_out/ncu_full.log
