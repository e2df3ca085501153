#!/bin/bash
# Round 2, GPU call G: samples-per-warp path enumeration (slot_decode) -- bit-identity tests, then A/B per workload.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py -m gpu -q -x -k "enumeration or equal_seed or tail" 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "spw0|MSK_SAMPLES_PER_WARP=0|" "spw1|MSK_SAMPLES_PER_WARP=1|" "spw4|MSK_SAMPLES_PER_WARP=4|" "spw8|MSK_SAMPLES_PER_WARP=8|" "spw32|MSK_SAMPLES_PER_WARP=32|" \
   "spw32_ploc|MSK_SAMPLES_PER_WARP=32 MSK_BVH_BUILDER=ploc|" "spw0_ploc|MSK_SAMPLES_PER_WARP=0 MSK_BVH_BUILDER=ploc|" "spw32_shst2|MSK_SAMPLES_PER_WARP=32 MSK_SHADOW_STATIC_BOUNCES=2|" "spw32_shst1|MSK_SAMPLES_PER_WARP=32 MSK_SHADOW_STATIC_BOUNCES=1|"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "spw0|MSK_SAMPLES_PER_WARP=0|" "spw4|MSK_SAMPLES_PER_WARP=4|" "spw16|MSK_SAMPLES_PER_WARP=32|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "spw0|MSK_SAMPLES_PER_WARP=0|" "spw8|MSK_SAMPLES_PER_WARP=8|" "spw32|MSK_SAMPLES_PER_WARP=32|" "spw32_ploc|MSK_SAMPLES_PER_WARP=32 MSK_BVH_BUILDER=ploc|"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "spw0|MSK_SAMPLES_PER_WARP=0|" "spw32|MSK_SAMPLES_PER_WARP=32|"
} 2>&1 | tee gpurun_out/r02g_ab.txt
