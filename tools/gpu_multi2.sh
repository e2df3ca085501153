#!/bin/bash
# Multi-GPU session: peer-film test, then N-rank C2 bench with the library's NVLink peer reduction vs NCCL, C1 (where
# the reduce is a large share of a 1 ms step), and a reduced-spp C4.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_peer_film.py -m gpu -q -x 2>&1 | tail -5
run() { # name workload extra...
  name=$1; wl=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload $wl --no-cpu "$@" > gpurun_out/scale_${name}_n$N.json 2> gpurun_out/scale_${name}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${name}_n$N.json").read().strip().splitlines()[-1])
    print("${name} n=$N: %.1f Mpaths/s  %.3f ms/step  e2e %.1f  reduce=%s" % (d["value"] / 1e6, d["ms_per_step"], d["e2e"]["value"] / 1e6, d["config"].get("film_reduce")))
except Exception as e:
    print("${name}: no result", e)
PY
  tail -2 gpurun_out/scale_${name}_n$N.err
}
run c2_peer c2 --reduce peer
run c2_nccl c2 --reduce nccl
run c1_peer c1 --reduce peer --steps 20
run c1_nccl c1 --reduce nccl --steps 20
run c4_peer c4 --reduce peer --spp 256 --steps 3 --warmup 3
