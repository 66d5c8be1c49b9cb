#!/bin/bash
# two-GPU pass: mb_extract_multi on real devices, the bench under torchrun at N=2 and its one-process multi-device leg
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -5 gpurun_out/pytest_multi.log
python bench.py --steps 3 --warmup 3 --clips 2000 --no-cpu-baseline --no-secondary > gpurun_out/bench_1proc_2gpu.json 2> gpurun_out/bench_1proc_2gpu.err; echo "bench 1 proc exit $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --clips 4000 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench torchrun exit $?"
python - <<'PY'
import json
for f in ('bench_1proc_2gpu','bench_2gpu'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][-1])
        print(f, 'value %.1fM n_gpus %d e2e %s' % (d['value']/1e6, d['n_gpus'], json.dumps({k:(v if not isinstance(v,dict) else {kk:vv for kk,vv in v.items() if kk!='how' and kk!='note'}) for k,v in d['e2e'].items() if k!='batch'})))
        for s in d.get('secondary',[]): print('   ', s['config']['bufferSize'], '%.1fM'%(s['value']/1e6), 'frac %.3f'%s['roofline']['frac'], s['clocks'])
    except Exception as e:
        print(f, 'FAILED', e); print(open('gpurun_out/%s.err'%f).read()[-1500:])
PY
