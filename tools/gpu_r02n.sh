#!/bin/bash
# Round 2, GPU call N: root culling at emission (k_shade / root_test) -- parity suite, then A/B per workload.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "cull1||" "cull0|MSK_ROOT_CULL=0|" "nocull_build||nocull"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "cull1||" "cull0|MSK_ROOT_CULL=0|"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "cull1||" "cull0|MSK_ROOT_CULL=0|"
echo "== c4"; SKIP_TESTS=1 STEPS=1 tools/ab_knobs.sh c4 "cull1||" "cull0|MSK_ROOT_CULL=0|"
} 2>&1 | tee gpurun_out/r02n_ab.txt
python - <<P
import json
for n in ["cull1","cull0"]:
    pass
P
