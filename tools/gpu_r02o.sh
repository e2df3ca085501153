#!/bin/bash
# Round 2, GPU call O: full GPU suite on the current tree, then the default bench line (as the driver runs it) and the
# reference arm.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
( time python bench.py --impl reference > gpurun_out/r02o_bench_ref.json 2> gpurun_out/r02o_bench_ref.err ) 2>&1 | grep real
( time python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r02o_bench.err
python - <<P
import json
r=json.loads(open('gpurun_out/r02o_bench_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r02o_bench.json').read().strip().splitlines()[-1])
print('ref', round(r['value']/1e6,2), 'Mpaths/s', r['cpu_baseline']['cores'], 'cores', r['cpu_baseline'].get('build','')[:40])
print('ours value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ms', round(d['ms_per_step'],3), 'ratio e2e/ref', round(d['e2e']['value']/r['value'],1))
print('cold', d.get('e2e_cold'))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','bound','frac','achieved','peak','unit','traffic')})
for k,v in (d.get('sub') or {}).items(): print('sub', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms')
for k,v in (d.get('strong_scaling') or {}).items(): print('strong', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms')
print('cpu', d['cpu_baseline']['value']/1e6, d['cpu_baseline']['cores'])
P
