#!/bin/bash
# Round 2, GPU call Z: refill threshold of the any-hit lockstep driver (variants).
set -u
mkdir -p gpurun_out
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "default||" "refany4||refany4" "refany12||refany12" "refany16||refany16" "trith6||trith_any"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "default||" "refany4||refany4" "refany12||refany12"
} 2>&1 | tee gpurun_out/r02z_ab.txt
