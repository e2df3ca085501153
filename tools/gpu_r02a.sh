#!/bin/bash
# Round 2, GPU call A: parity of the v2 node layout, A/B of the traversal variants, one ncu capture.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
{
echo "== c2"; tools/variants.sh run c2 r01 default refill1 refill4 refill16 smem0 permalu tri8 tri16
echo "== c5"; tools/variants.sh run c5 r01 default refill1
echo "== c3"; tools/variants.sh run c3 r01 default
echo "== c1"; tools/variants.sh run c1 r01 default
} 2>&1 | tee gpurun_out/r02a_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 1 -c 2 -f -o gpurun_out/r02a_k_intersect \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/r02a_ncu.log 2>&1
ls -la gpurun_out | tail -5
