#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/runbig.py <<'PY'
import sys, os
sys.argv = ["x", "none"]
sys.path.insert(0, os.getcwd())
exec(open("tools/bench_configs.py").read().split("which = ")[0])
run("config5 N=32768 hop=8192", 32768, 8192, 128, 2646000, C5)
PY
python /tmp/runbig.py > gpurun_out/plainbig.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_big32768 -s 2 -c 1 -f -o gpurun_out/prof_big32768 python /tmp/runbig.py > gpurun_out/ncu_big.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_big.log; cat gpurun_out/plainbig.log | tail -1
