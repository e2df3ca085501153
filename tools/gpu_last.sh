#!/bin/bash
# Last short validation of the round on one GPU: smoke, parity tests, ncu --set full of the dominant kernel, default bench line.
set -u
mkdir -p gpurun_out
timeout 60 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
timeout 100 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 70 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 1 -c 1 -f -o gpurun_out/prof_k_intersect_r01i \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_k_intersect_r01i.log 2>&1; echo "ncu rc=$?"
timeout 60 python bench.py > gpurun_out/bench_c2_r01i.json 2> gpurun_out/bench_c2_r01i.err; tail -c 300 gpurun_out/bench_c2_r01i.json
