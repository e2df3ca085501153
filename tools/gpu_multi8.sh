#!/bin/bash
# 8-GPU box: C2 weak scaling at N = 4 and 8 with the peer-memory film reduction, C1 peer vs NCCL at N = 8, C4 at N = 8
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
run() { # name N workload extra...
  name=$1; N=$2; wl=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --no-cpu "$@" > gpurun_out/scale_${name}_n$N.json 2> gpurun_out/scale_${name}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${name}_n$N.json").read().strip().splitlines()[-1])
    print("${name} n=$N: %.1f Mpaths/s  %.3f ms/step  e2e %.1f  reduce=%s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["config"].get("film_reduce")))
except Exception as e:
    print("${name}: no result", e)
PY
  tail -2 gpurun_out/scale_${name}_n$N.err | grep -v OMP_NUM\|\*\*\*\*
}
run c2_peer 4 c2 --reduce peer
run c2_peer 8 c2 --reduce peer
run c2_nccl 8 c2 --reduce nccl
run c1_peer 8 c1 --reduce peer --steps 20
run c1_nccl 8 c1 --reduce nccl --steps 20
run c4_peer 8 c4 --reduce peer --spp 512 --steps 3 --warmup 3
