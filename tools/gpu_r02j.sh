#!/bin/bash
# Round 2, GPU call J: first-bounce state elision, shared-address variants of the traversal stack / LUT, PLOC on C5.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_intersect.py tests/test_aov.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "elide1||" "elide0|MSK_FIRST_ELIDE=0|" "orig||orig" "plainptr||plainptr" "plainsmem||plainsmem" "opaque||opaque"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "elide1||" "orig||orig" "opaque||opaque" "plainsmem||plainsmem"
echo "== c5"; SKIP_TESTS=1 tools/ab_knobs.sh c5 "lbvh||" "ploc|MSK_BVH_BUILDER=ploc|" "orig||orig" "opaque||opaque"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "elide1||" "elide0|MSK_FIRST_ELIDE=0|"
} 2>&1 | tee gpurun_out/r02j_ab.txt
