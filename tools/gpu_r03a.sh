#!/bin/bash
# Round 2, session 3, GPU call A: packed-FP32 node / triangle tests (MSK_F32X2) and the shared-memory pool of prepared rays
# (MSK_RAY_POOL) with its refill thresholds: bit-identity of the films against the previous build, parity tests, A/B.
set -u
mkdir -p gpurun_out
V=$PWD/build/variants
{
echo "== base (scalar, per-lane prefetch)"; MSK_B200_LIB=$V/base/libmisaki_b200.so timeout 300 python tools/film_hash.py
echo "== f2";   MSK_B200_LIB=$V/f2/libmisaki_b200.so timeout 300 python tools/film_hash.py
echo "== this build (f32x2 + pool, refill 8)"; timeout 300 python tools/film_hash.py
echo "== pool2"; MSK_B200_LIB=$V/pool2/libmisaki_b200.so timeout 300 python tools/film_hash.py
} 2>&1 | tee gpurun_out/r03a_film_hash.txt
timeout 900 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_sweep.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "base||base" "f2||f2" "pool8||" "pool4||pool4" "pool2||pool2" "pool1||pool1" "pool4s4||pool4s4" "pool4t16||pool4t16" "pool2t16||pool2t16" "pool4nof2||pool4nof2" "base_again||base"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "base||base" "f2||f2" "pool8||" "pool4||pool4" "pool2||pool2"
echo "== c5"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c5 "base||base" "f2||f2" "pool8||" "pool4||pool4" "pool2||pool2"
} 2>&1 | tee gpurun_out/r03a_ab.txt
