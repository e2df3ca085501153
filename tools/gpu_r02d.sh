#!/bin/bash
# Round 2, GPU call D: tests; leaf-size and ray-reordering experiments; C3 per-bounce queue lengths; ncu counters + full captures.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
{
echo "== c2"; tools/variants.sh run c2 default leaf1 leaf2
echo "== c2 MSK_RAY_SORT=1"; MSK_RAY_SORT=1 tools/variants.sh run c2 default
echo "== c3"; tools/variants.sh run c3 default leaf1 leaf2
echo "== c3 MSK_RAY_SORT=1"; MSK_RAY_SORT=1 tools/variants.sh run c3 default
echo "== c5"; tools/variants.sh run c5 default leaf1 leaf2
} 2>&1 | tee gpurun_out/r02d_ab.txt
MSK_DEBUG_BOUNCES=1 python bench.py --workload c3 --steps 1 --warmup 0 --no-cpu 2> gpurun_out/r02d_c3_bounces.txt > /dev/null
grep -c "ran over" gpurun_out/r02d_c3_bounces.txt
python tools/ncu_counters.py run c1 c2 c3 c5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 1 -c 1 -f -o gpurun_out/r02d_k_intersect python bench.py --one-step > gpurun_out/r02d_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 4 -c 1 -f -o gpurun_out/r02d_k_shade python bench.py --one-step > gpurun_out/r02d_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 1 -c 1 -f -o gpurun_out/r02d_k_shadow python bench.py --one-step > gpurun_out/r02d_ncu3.log 2>&1
ls -la gpurun_out | tail -12
