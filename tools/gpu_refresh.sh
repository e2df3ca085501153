#!/bin/bash
# Refresh the per-workload bench lines on the round-end tree (device-timed, --no-cpu: the CPU leg is in the r01f/r01g lines).
set -u
mkdir -p gpurun_out
for wl in c1 c3 vol; do
  timeout 28 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${wl}_r01i.json 2>/dev/null
  python - "$wl" <<'PY'
import json,sys
wl=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_{wl}_r01i.json').read().strip().splitlines()[-1])
    print(wl, '%.1f Mpaths/s' % (d['value']/1e6), 'ms/step %.2f' % d['ms_per_step'], 'frac %.3f' % d['roofline']['frac'])
except Exception as e:
    print(wl, 'no line', e)
PY
done
