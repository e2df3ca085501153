#!/bin/bash
# Multi-GPU session: N-rank bench of C2 (weak scaling, film reduce inside the timed region) and a reduced-spp C4
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --no-cpu > gpurun_out/scale_c2_n1.json 2> gpurun_out/scale_c2_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n > gpurun_out/scale_c2_n$n.json 2> gpurun_out/scale_c2_n$n.err
  fi
  tail -c 1500 gpurun_out/scale_c2_n$n.json; tail -3 gpurun_out/scale_c2_n$n.err
done
# C4 (Cornell box 1920x1080) at 256 spp per GPU: the config's 4096 spp takes ~10 s per step
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --workload c4 --spp 256 --steps 3 --warmup 3 > gpurun_out/scale_c4_n$N.json 2> gpurun_out/scale_c4_n$N.err
tail -c 1500 gpurun_out/scale_c4_n$N.json; tail -3 gpurun_out/scale_c4_n$N.err
timeout 600 python -m pytest tests -m gpu -q -k distributed 2>&1 | tail -5
