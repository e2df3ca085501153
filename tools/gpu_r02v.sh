#!/bin/bash
# CPU arm on the GPU box's host: checker build, speed build with and without the ORC_FAST walk (C2 and C1).
set -u
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo
for wl in c2 c1; do
python bench.py --impl reference --workload $wl --ref-kind port --steps 3 --warmup 1 --ref-budget 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl port        ', round(d['value']/1e6,2))"
python bench.py --impl reference --workload $wl --ref-kind fast --steps 3 --warmup 1 --ref-budget 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl fast+ORC_FAST', round(d['value']/1e6,2))"
FASTFLAGS="-O3 -march=native -ffp-contract=fast -funroll-loops -fno-math-errno -std=c++17 -fPIC -pthread -w" python bench.py --impl reference --workload $wl --ref-kind fast --steps 3 --warmup 1 --ref-budget 40 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl fast flags only', round(d['value']/1e6,2))"
done 2>&1 | tee gpurun_out/r02v_cpu_arm.txt
