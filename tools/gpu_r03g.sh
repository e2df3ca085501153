#!/bin/bash
# Round 2, session 3, GPU call G: packet traversal of the camera-ray queue (and, as an experiment, of the bounce-0 shadow
# queue): bit-identity of the films with the knob off / on, parity tests, A/B over runtime knobs of ONE build.
set -u
mkdir -p gpurun_out
{
echo "== MSK_PACKET_CAMERA=0"; MSK_PACKET_CAMERA=0 timeout 300 python tools/film_hash.py
echo "== MSK_PACKET_CAMERA=1 (default)"; timeout 300 python tools/film_hash.py
echo "== MSK_PACKET_CAMERA=1 MSK_PACKET_SHADOW=1"; MSK_PACKET_SHADOW=1 timeout 300 python tools/film_hash.py
} 2>&1 | tee gpurun_out/r03g_film_hash.txt
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_aov.py tests/test_gpu_baseline_sizes.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "off|MSK_PACKET_CAMERA=0|" "cam||" "cam_sh|MSK_PACKET_SHADOW=1|" "off_again|MSK_PACKET_CAMERA=0|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "off|MSK_PACKET_CAMERA=0|" "cam||" "cam_sh|MSK_PACKET_SHADOW=1|"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "off|MSK_PACKET_CAMERA=0|" "cam||" "cam_sh|MSK_PACKET_SHADOW=1|"
echo "== c4"; SKIP_TESTS=1 STEPS=1 tools/ab_knobs.sh c4 "off|MSK_PACKET_CAMERA=0|" "cam||"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "off|MSK_PACKET_CAMERA=0|" "cam||"
} 2>&1 | tee gpurun_out/r03g_ab.txt
