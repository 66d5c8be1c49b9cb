#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/run512.py <<'PY'
import sys, os
sys.argv = ["x", "none"]
sys.path.insert(0, os.getcwd())
exec(open("tools/bench_configs.py").read().split("C3 = ")[0])
run("full set N=512 hop=N", 512, 512, 400, 441000, mb.FEATURES)
PY
python /tmp/run512.py > gpurun_out/plain512.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_generic -s 2 -c 1 -f -o gpurun_out/prof_generic512 python /tmp/run512.py > gpurun_out/ncu_512.log 2>&1
echo "exit $?"; tail -2 gpurun_out/ncu_512.log; cat gpurun_out/plain512.log | tail -1
