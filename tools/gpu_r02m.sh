#!/bin/bash
# Round 2, GPU call M: ncu counters of one C2 step + full captures of the bounce-0 / bounce-1 closest-hit launches and the
# bounce-0 shadow launch on the tree with samples-per-warp enumeration and two batches in flight.
set -u
mkdir -p gpurun_out
python tools/ncu_counters.py run c2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 0 -c 1 -f -o gpurun_out/r02m_k_intersect_b0 python bench.py --one-step > gpurun_out/r02m_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 2 -c 1 -f -o gpurun_out/r02m_k_intersect_b1 python bench.py --one-step > gpurun_out/r02m_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 0 -c 1 -f -o gpurun_out/r02m_k_shadow_b0 python bench.py --one-step > gpurun_out/r02m_ncu3.log 2>&1
tail -2 gpurun_out/r02m_ncu*.log
ls -la gpurun_out | grep r02m
