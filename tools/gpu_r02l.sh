#!/bin/bash
# Round 2 (after the enumeration + lanes work), multi-GPU call (N = number of GPUs of the box): in-process multi-device tests on real devices, the IPC peer-film
# test, bench at N (default line: C2 weak + sub-results incl. BASELINE configs[3] strong), the host CLI on all GPUs.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r02l_gpus_n$N.txt
timeout 600 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_peer_film.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r02l_pytest_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02l_bench_n$N.json 2> gpurun_out/r02l_bench_n$N.err
tail -c 300 gpurun_out/r02l_bench_n$N.err
python - <<P
import json
d=json.loads(open('gpurun_out/r02l_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N C2 weak', round(d['value']/1e6,1), 'Mpaths/s', round(d['ms_per_step'],3), 'ms; e2e', round(d['e2e']['value']/1e6,1), d['config']['film_reduce'])
for k,v in (d.get('sub') or {}).items(): print('sub', k, round(v['value']/1e6,1), round(v['ms_per_step'],3), 'ms')
for k,v in (d.get('strong_scaling') or {}).items(): print('strong', k, round(v['value']/1e6,1), 'Mpaths/s', round(v['ms_per_step'],3), 'ms', v['config']['job_spp'], v['config']['spp_per_gpu'])
P
# the host front-end's CLI on every GPU of the box: one process, msk_gpu_render_multi
( nvidia-smi --query-gpu=index,utilization.gpu --format=csv,noheader -lms 250 > gpurun_out/r02l_cli_util_n$N.csv & echo $! > /tmp/smi.pid )
MSK_DEVICES=all timeout 300 misaki_render_b200/lib/misaki_b200 assets/scenes/cbox.xml -D w=1920 -D h=1080 -D spp=1024 -D depth=5 -o gpurun_out/r02l_cbox_n$N.exr 2>&1 | grep -E "GPU scene|Rendering finished" | tee gpurun_out/r02l_cli_n$N.txt
kill $(cat /tmp/smi.pid)
python - <<P
import collections
act=collections.defaultdict(int)
for l in open('gpurun_out/r02l_cli_util_n$N.csv'):
    i,u=l.split(','); act[int(i)]=max(act[int(i)], int(u.strip().split()[0]))
print('max utilisation per GPU during the CLI render:', dict(act), 'gpus_active', sum(1 for v in act.values() if v>30))
P
rm -f gpurun_out/r02l_cbox_n$N.exr
