#!/bin/bash
# Round 2, final measurement pass, part 2: the default bench line (as the driver runs it) + reference arm, C4 and fog lines.
set -u
mkdir -p gpurun_out
( time python bench.py --impl reference > gpurun_out/r02u_bench_ref.json 2> gpurun_out/r02u_bench_ref.err ) 2>&1 | grep real
( time python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err ) 2>&1 | grep real
tail -3 gpurun_out/r02u_bench.err
python bench.py --workload vol --steps 5 --warmup 3 --no-sub > gpurun_out/r02u_bench_vol.json 2> gpurun_out/r02u_bench_vol.err
python bench.py --workload c4 --steps 2 --warmup 1 --no-sub --no-cpu > gpurun_out/r02u_bench_c4.json 2> gpurun_out/r02u_bench_c4.err
python bench.py --workload c3 --steps 3 --warmup 1 --no-sub > gpurun_out/r02u_bench_c3.json 2> gpurun_out/r02u_bench_c3.err
python - <<P
import json
r=json.loads(open('gpurun_out/r02u_bench_ref.json').read().strip().splitlines()[-1])
d=json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print('ref', round(r['value']/1e6,2), 'Mpaths/s', r['cpu_baseline']['cores'], 'cores')
print('ours value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'ms', round(d['ms_per_step'],3), 'ratio e2e/ref', round(d['e2e']['value']/r['value'],1))
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('kernel','bound','frac','achieved','peak','unit','traffic')})
print('kernels', json.dumps(d['roofline']['kernels']))
for k,v in (d.get('sub') or {}).items(): print('sub', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms', json.dumps((v.get('roofline') or {}).get('kernels', (v.get('roofline') or {}).get('issue')))[:600])
for k,v in (d.get('strong_scaling') or {}).items(): print('strong', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms')
for n in ['vol','c4','c3']:
    x=json.loads(open('gpurun_out/r02u_bench_%s.json'%n).read().strip().splitlines()[-1])
    print(n, round(x['value']/1e6,1), round(x['ms_per_step'],2), 'ms', 'cpu', (x.get('cpu_baseline') or {}).get('value'))
P
