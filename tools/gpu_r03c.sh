#!/bin/bash
# Round 2, session 3, GPU call C: shared-memory stack depth under the ray pool (14-15 KB per CTA fits the 100 KB carve-out at
# 6 CTAs per SM: 128 KB of L1 instead of 96), full capture of the bounce-1 closest-hit launch with the pool.
set -u
mkdir -p gpurun_out
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "s8||" "s6||s6" "s3||s3" "s2||s2" "s8_again||"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "s8||" "s6||s6" "s3||s3" "s2||s2"
echo "== c5"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c5 "s8||" "s3||s3" "s2||s2"
} 2>&1 | tee gpurun_out/r03c_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 2 -c 1 -f -o gpurun_out/r03c_k_intersect_b1 python bench.py --one-step > gpurun_out/r03c_ncu1.log 2>&1
