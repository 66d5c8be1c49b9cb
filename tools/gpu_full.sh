#!/bin/bash
# Round checkpoint: every GPU test, the driver's default bench (both arms), then the ncu launch list of a short bench.
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -1 gpurun_out/bench_ref.log
CMD="python bench.py --clips 600 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
