#!/bin/bash
# Round 2, GPU call X: any-hit queries without front-to-back ordering (variant), parity of occlusion queries.
set -u
mkdir -p gpurun_out
MSK_B200_LIB=$PWD/build/variants/anyunord/libmisaki_b200.so timeout 900 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_sweep.py -m gpu -q -x 2>&1 | tail -3
{ echo "== default"; python tools/film_hash.py; echo "== anyunord"; MSK_B200_LIB=$PWD/build/variants/anyunord/libmisaki_b200.so python tools/film_hash.py; } 2>&1 | tee gpurun_out/r02x_film_hash.txt
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "default||" "anyunord||anyunord"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "default||" "anyunord||anyunord"
echo "== c5"; SKIP_TESTS=1 tools/ab_knobs.sh c5 "default||" "anyunord||anyunord"
} 2>&1 | tee gpurun_out/r02x_ab.txt
