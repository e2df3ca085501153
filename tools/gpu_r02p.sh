#!/bin/bash
# Round 2, GPU call P: shape of the wide BVH (children per node, triangles per leaf) for both builders on C2 / C3 / C5.
set -u
mkdir -p gpurun_out
for wl in c2 c3 c5; do for b in lbvh ploc; do
  echo "== $wl $b"
  MSK_DEBUG_SETUP=1 MSK_BVH_BUILDER=$b python bench.py --workload $wl --one-step 2>&1 | grep "wide BVH\|scene_create" | head -2
done; done 2>&1 | tee gpurun_out/r02p_bvh_shape.txt
