#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/runc3.py <<'PY'
import sys, os
sys.argv = ["x", "none"]
sys.path.insert(0, os.getcwd())
exec(open("tools/bench_configs.py").read().split("which = ")[0])
run("config3 mfcc+moments", 2048, 512, 1000, 441000, C3)
PY
python /tmp/runc3.py > gpurun_out/plainc3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_warp2048 -s 2 -c 1 -f -o gpurun_out/prof_c3 python /tmp/runc3.py > gpurun_out/ncu_c3.log 2>&1
echo "exit $?"; tail -1 gpurun_out/ncu_c3.log; cat gpurun_out/plainc3.log | tail -1
