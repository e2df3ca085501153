#!/bin/bash
# Round 2, session 3, GPU call B: refill thresholds of the shared-memory ray pool (scalar node test), per workload.
set -u
mkdir -p gpurun_out
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "base||base" "p8||p8" "p6||p6" "p4||p4" "p3||p3" "p2||p2" "p4a8||p4a8" "p8a4||p8a4" "p4t10||p4t10" "p4t14||p4t14" "base_again||base"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "base||base" "p8||p8" "p6||p6" "p4||p4" "p3||p3" "p4a8||p4a8" "p8a4||p8a4"
echo "== c5"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c5 "base||base" "p8||p8" "p6||p6" "p4||p4" "p3||p3" "p2||p2"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "base||base" "p8||p8" "p4||p4"
} 2>&1 | tee gpurun_out/r03b_ab.txt
