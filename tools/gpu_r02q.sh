#!/bin/bash
# Round 2, GPU call Q: SAH-optimal collapse (k_plan) -- intersection parity tests, tree shapes, A/B against the greedy collapse.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_sweep.py tests/test_gpu_baseline_sizes.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
for wl in c2 c5; do for b in lbvh ploc; do
  echo "== $wl $b sah-collapse"
  MSK_DEBUG_SETUP=1 MSK_BVH_BUILDER=$b python bench.py --workload $wl --one-step 2>&1 | grep "wide BVH\|scene_create" | head -2
done; done 2>&1 | tee gpurun_out/r02q_bvh_shape.txt
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "sah||" "greedy|MSK_BVH_COLLAPSE=greedy|" "sah_ploc|MSK_BVH_BUILDER=ploc|" "sah_prim05|MSK_BVH_CPRIM=0.5|" "sah_prim2|MSK_BVH_CPRIM=2|" "sah_ploc_prim05|MSK_BVH_BUILDER=ploc MSK_BVH_CPRIM=0.5|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "sah||" "greedy|MSK_BVH_COLLAPSE=greedy|" "sah_ploc|MSK_BVH_BUILDER=ploc|" "sah_prim05|MSK_BVH_CPRIM=0.5|"
echo "== c5"; SKIP_TESTS=1 tools/ab_knobs.sh c5 "sah||" "greedy|MSK_BVH_COLLAPSE=greedy|" "sah_ploc|MSK_BVH_BUILDER=ploc|" "sah_prim05|MSK_BVH_CPRIM=0.5|"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "sah||" "greedy|MSK_BVH_COLLAPSE=greedy|"
} 2>&1 | tee gpurun_out/r02q_ab.txt
