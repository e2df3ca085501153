#!/bin/bash
# Round 2, GPU call B: new tests (multi-device, validation), the rewritten bench (sub-results), occupancy variants,
# set-up breakdown, ncu counter passes.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r02b_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02b_bench.json').read().strip().splitlines()[-1])
print('C2', round(d['value']/1e6,1), 'Mpaths/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value']/1e6,1), 'cold', d.get('e2e_cold'))
print('roofline', {k: d['roofline'][k] for k in ('bound','kernel','achieved','peak','unit','frac','traffic','counters')})
print('stage_ms', d['roofline']['stage_ms'])
for k,v in (d.get('sub') or {}).items(): print('sub', k, round(v['value']/1e6,1), v['unit'], round(v['ms_per_step'],3), 'ms')
for k,v in (d.get('strong_scaling') or {}).items(): print('strong', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms')
P
{
echo "== c2"; tools/variants.sh run c2 default blocks5 blocks7 blocks8
echo "== c3"; tools/variants.sh run c3 default blocks7
} 2>&1 | tee gpurun_out/r02b_ab.txt
MSK_DEBUG_SETUP=1 python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu 2>&1 >/dev/null | grep "\[msk\]" | tee gpurun_out/r02b_setup.txt
python tools/ncu_counters.py run c1 c2
ls -la gpurun_out | tail -8
