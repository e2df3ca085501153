#!/bin/bash
# Round 2, GPU call W: film gather with staged weights -- bit-identity of the films against the previous build, tests, A/B.
set -u
mkdir -p gpurun_out
{ echo "== previous build"; MSK_B200_LIB=$PWD/build/variants/pregather/libmisaki_b200.so python tools/film_hash.py; echo "== this build"; python tools/film_hash.py; } 2>&1 | tee gpurun_out/r02w_film_hash.txt
timeout 1500 python -m pytest tests/test_gpu_render.py tests/test_aov.py tests/test_gpu_baseline_sizes.py tests/test_gpu_host_render.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "new||" "old||pregather"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "new||" "old||pregather"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "new||" "old||pregather"
echo "== c4"; SKIP_TESTS=1 STEPS=1 tools/ab_knobs.sh c4 "new||" "old||pregather"
} 2>&1 | tee gpurun_out/r02w_ab.txt
