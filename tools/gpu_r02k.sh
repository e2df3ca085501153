#!/bin/bash
# Round 2, GPU call K: first-bounce state elision (runtime knob), several triangles per lane and triangle phase (variants).
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_gpu_intersect.py tests/test_aov.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "elide1||" "elide0|MSK_FIRST_ELIDE=0|" "trireps2||trireps2" "trireps3||trireps3" "trireps4||trireps4"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "elide1||" "elide0|MSK_FIRST_ELIDE=0|" "trireps2||trireps2" "trireps3||trireps3" "trireps4||trireps4"
echo "== c5"; SKIP_TESTS=1 tools/ab_knobs.sh c5 "default||" "trireps2||trireps2" "trireps3||trireps3" "trireps4||trireps4"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "default||" "trireps3||trireps3"
} 2>&1 | tee gpurun_out/r02k_ab.txt
