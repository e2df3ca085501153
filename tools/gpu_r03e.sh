#!/bin/bash
# Round 2, session 3, GPU call E: shadow queue carries the path index in the ray record (no sh_path load in the commit);
# lockstep loop rotated to triangle -> pop -> node with an L1 prefetch of the next triangle.  Bit-identity, parity, A/B.
set -u
mkdir -p gpurun_out
V=$PWD/build/variants
{
echo "== pretag (committed tree e1ed2be)"; MSK_B200_LIB=$V/pretag/libmisaki_b200.so timeout 300 python tools/film_hash.py
echo "== this build (tag)"; timeout 300 python tools/film_hash.py
echo "== rotpf1"; MSK_B200_LIB=$V/rotpf1/libmisaki_b200.so timeout 300 python tools/film_hash.py
} 2>&1 | tee gpurun_out/r03e_film_hash.txt
timeout 900 python -m pytest tests/test_gpu_intersect.py tests/test_gpu_sweep.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "pretag||pretag" "tag||" "rot||rot" "rotpf1||rotpf1" "rotpf2||rotpf2" "pretag_again||pretag"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "pretag||pretag" "tag||" "rot||rot" "rotpf1||rotpf1" "rotpf2||rotpf2"
echo "== c5"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c5 "tag||" "rot||rot" "rotpf1||rotpf1" "rotpf2||rotpf2"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "pretag||pretag" "tag||" "rotpf1||rotpf1"
} 2>&1 | tee gpurun_out/r03e_ab.txt
