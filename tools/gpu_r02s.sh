#!/bin/bash
# Round 2, GPU call S: full GPU suite (develop kernel, collapse test), static closest-hit driver for the second bounce.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "static1||" "static2|MSK_STATIC_BOUNCES=2|" "static3|MSK_STATIC_BOUNCES=3|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "static1||" "static2|MSK_STATIC_BOUNCES=2|"
} 2>&1 | tee gpurun_out/r02s_ab.txt
