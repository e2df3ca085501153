#!/bin/bash
set -u
for b in lbvh ploc; do for c in sah greedy; do echo "== c5 $b $c"; MSK_DEBUG_SETUP=1 MSK_BVH_BUILDER=$b MSK_BVH_COLLAPSE=$c python bench.py --workload c5 --one-step 2>&1 | grep "k_plan\|collapse level\|scene_create" | head -20; done; done 2>&1 | tee gpurun_out/r02r_build_times.txt
