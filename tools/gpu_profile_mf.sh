#!/bin/bash
mkdir -p gpurun_out
cat > gpurun_out/_runmf.py <<'PY'
import sys, os
sys.argv = ["x", "none"]
sys.path.insert(0, os.getcwd())
exec(open("tools/bench_configs.py").read().split("C3 = ")[0])
run("full set N=512 hop=N", 512, 512, 3000, 441000, mb.FEATURES)
PY
python gpurun_out/_runmf.py > gpurun_out/plainmf.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mb_warpmf -s 2 -c 1 -f -o gpurun_out/prof_warpmf512 python gpurun_out/_runmf.py > gpurun_out/ncu_mf.log 2>&1
echo "exit $?"; tail -1 gpurun_out/ncu_mf.log; cat gpurun_out/plainmf.log | tail -1
