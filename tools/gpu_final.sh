#!/bin/bash
# The round's closing run on one B200 box: checks, the bench lines and the profiles that profiles/README.md cites.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
T=${1:-r02}
bash tools/gpu.sh check
bash tools/gpu.sh bench --steps 20 --warmup 5; cp gpurun_out/bench.json gpurun_out/${T}_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm exit $?"
python tools/bench_configs.py smallab small c1 c3 c5 large exactcmp exact pcm sizes > gpurun_out/${T}_configs.json 2>&1; echo "configs exit $?"
bash tools/gpu.sh adaptive > gpurun_out/${T}_adaptive.txt; tail -13 gpurun_out/${T}_adaptive.txt
LL="python bench.py --clips 600 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
bash tools/gpu.sh profile warp2048 mb_warp2048 $LL
bash tools/gpu.sh profile exactw mb_exact_warp python tools/bench_configs.py exactone
bash tools/gpu.sh profile big32768 mb_big32768 python tools/bench_configs.py c5
bash tools/gpu.sh profile mf256 mb_warpmf python tools/bench_configs.py smallab
