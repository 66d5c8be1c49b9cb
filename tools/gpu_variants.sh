#!/bin/bash
# A/B of library variants on the headline workload (short run each).
mkdir -p gpurun_out
for so in meyda_b200/_lib/libmeyda_b200.so $(ls meyda_b200/_lib/variants/lib_*.so 2>/dev/null) meyda_b200/_lib/libmeyda_b200.so $(ls meyda_b200/_lib/variants/lib_*.so 2>/dev/null); do
  MEYDA_B200_LIB=$PWD/$so timeout 600 python bench.py --clips ${CLIPS:-1200} --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/variant.log 2>&1
  python - "$so" <<'PY'
import json,sys
for line in open('gpurun_out/variant.log'):
    if line.startswith('{'):
        d=json.loads(line); print(sys.argv[1].split('/')[-1], '%.1fM frames/s'%(d['value']/1e6), 'frac %.3f'%d['roofline']['frac'], d['parity'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
        break
else:
    print(sys.argv[1], 'FAILED'); print(open('gpurun_out/variant.log').read()[-600:])
PY
done
