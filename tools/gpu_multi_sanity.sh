#!/bin/bash
# N-GPU sanity: peer-film test + C2 / C1 bench lines with the peer-memory reduction
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_peer_film.py -m gpu -q -x 2>&1 | tail -2
for wl in c2 c1; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --no-cpu > gpurun_out/final_${wl}_n$N.json 2> gpurun_out/final_${wl}_n$N.err
  python -c "
import json
d=json.loads(open('gpurun_out/final_${wl}_n$N.json').read().strip().splitlines()[-1])
print('$wl n=$N: %.1f Mpaths/s %.3f ms/step e2e %.1f reduce=%s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['config']['film_reduce'][:40]))"
  grep -v "OMP_NUM\|\*\*\*" gpurun_out/final_${wl}_n$N.err | tail -2
done
