#!/bin/bash
# Everything this repo runs on a GPU box, one sub-command each (gpurun -- 'bash tools/gpu.sh <what> [args]'):
#   check            smoke, every GPU test, a longer seeded fuzz soak
#   tests [-k expr]  the GPU test suite (extra arguments go to pytest)
#   bench [args]     bench.py (default flags) into gpurun_out/bench.json and a one-line digest
#   configs <names>  tools/bench_configs.py <names> (c3 c5 exactcmp c5ab large small sizes pcm ...)
#   adaptive         tools/diag_adaptive.py: refined-frame counts and band-free parity of BASELINE config 2
#   variants         A/B of meyda_b200/_lib/variants/lib_*.so against the working tree (tools/build_variants.sh)
#   profile <name> <kernel regex> <command...>
#                    launch list of `python bench.py --clips 600 ...` and one `ncu --set full` capture of <command>
#                    into gpurun_out/prof_<name>.ncu-rep (each after the same command ran clean without ncu)
#   multi            two or more GPUs: tests/test_gpu_multi.py, the bench in one process and under torchrun
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
mkdir -p gpurun_out
what=$1; shift
digest() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
print('value %.1fM frames/s frac %.3f n_gpus %d parity %s banded %s refined(last wave) %s launches %s clocks %s' % (
    d['value'] / 1e6, d['roofline']['frac'], d['n_gpus'], (d['parity'] or '')[:40], d.get('parity_banded'),
    d.get('refined_frames_last_wave'), d['gpu_launches'], d['clocks']))
e = d.get('e2e')
if e:
    print('e2e %.3fM frames/s, raw pinned-copy ceiling %.3fM (%.2f), pageable %.3fM%s' % (
        e['value'] / 1e6, e['pcie_ceiling']['value'] / 1e6, e['frac_of_pcie_ceiling'], e['pageable']['value'] / 1e6,
        ', one process x %d devices %.3fM' % (e['multi_device_one_process']['devices'], e['multi_device_one_process']['value'] / 1e6)
        if 'multi_device_one_process' in e else ''))
for s in d.get('secondary', []):
    print('  N=%d %s: %.2fM frames/s, %.3f of FP32 (%.1f of %.1f TFLOP/s), %.3f of HBM, parity %s, clocks %s' % (
        s['config']['bufferSize'], s['kernel'], s['value'] / 1e6, s['roofline']['frac'], s['roofline']['achieved'],
        s['roofline']['peak'], s['roofline']['hbm_frac'], (s['parity'] or '')[:30], s['clocks']))
if d.get('cpu_baseline'):
    print('cpu baseline', d['cpu_baseline'])
PY
}
case $what in
check)
  timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
  timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
  grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -30; tail -3 gpurun_out/pytest_gpu.log
  MEYDA_FUZZ_CASES=${FUZZ_CASES:-150} MEYDA_FUZZ_SEED=${FUZZ_SEED:-777} timeout 600 python -m pytest tests/test_gpu_fuzz.py -m gpu -q --maxfail=10 -p no:cacheprovider > gpurun_out/pytest_fuzz_soak.log 2>&1; echo "soak exit $?"
  tail -3 gpurun_out/pytest_fuzz_soak.log ;;
tests)
  python -m pytest tests -m gpu -q -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
  grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head -30; tail -3 gpurun_out/pytest_gpu.log ;;
bench)
  python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -3 gpurun_out/bench.err; digest gpurun_out/bench.json ;;
configs)
  python tools/bench_configs.py "$@" > gpurun_out/bench_configs.json 2>&1
  python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("%-58s %-12s %9.2f M frames/s  refined %s" % (d["config"][:58], d["kernel"], d["frames_per_s"] / 1e6, d.get("refined")))
    else:
        print(l.rstrip()[:200])
PY
  ;;
adaptive)
  python tools/diag_adaptive.py > gpurun_out/diag_adaptive.log 2>&1; grep -v Warning gpurun_out/diag_adaptive.log | grep -v "err = " ;;
variants)
  for so in meyda_b200/_lib/libmeyda_b200.so $(ls meyda_b200/_lib/variants/lib_*.so 2>/dev/null) meyda_b200/_lib/libmeyda_b200.so $(ls meyda_b200/_lib/variants/lib_*.so 2>/dev/null); do
    MEYDA_B200_LIB=$PWD/$so timeout 600 python bench.py --clips ${CLIPS:-2400} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary $BENCH_EXTRA > gpurun_out/variant.log 2>&1
    echo -n "$(basename $so): "; digest gpurun_out/variant.log | head -1
  done ;;
profile)
  name=$1; re=$2; shift 2
  LL="python bench.py --clips 600 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
  if [ "$name" = warp2048 ]; then
    $LL > gpurun_out/plain.log 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $LL > gpurun_out/ncu_launches.log 2>&1
    echo "launch list exit $?"
  fi
  "$@" > gpurun_out/plain_$name.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$re -s ${SKIP:-3} -c 1 -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  echo "full capture exit $?"; tail -2 gpurun_out/ncu_$name.log ;;
multi)
  nvidia-smi -L
  python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -3 gpurun_out/pytest_multi.log
  n=$(nvidia-smi -L | wc -l)
  python bench.py --steps 3 --warmup 3 --clips 2000 --no-cpu-baseline --no-secondary > gpurun_out/bench_1proc.json 2> gpurun_out/bench_1proc.err; echo "one process exit $?"; digest gpurun_out/bench_1proc.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-5} --warmup 3 "$@" > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err; echo "torchrun x$n exit $?"; digest gpurun_out/bench_${n}gpu.json ;;
*) echo "unknown sub-command $what"; exit 2 ;;
esac
