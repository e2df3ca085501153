#!/bin/bash
# Round 2, GPU call H: film-record exchange (k_film_records), per-bounce queue lengths of VOL / C2, tail threshold on VOL.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_aov.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "spw0|MSK_SAMPLES_PER_WARP=0|" "spw8|MSK_SAMPLES_PER_WARP=8|" "spw32|MSK_SAMPLES_PER_WARP=32|" "spw32_b8m|MSK_BATCH_PATHS=8388608|" "spw32_b4m|MSK_BATCH_PATHS=4194304|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "spw32|MSK_SAMPLES_PER_WARP=32|" "spw32_lbvh_tail1m|MSK_TAIL_THRESHOLD=1048576|"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "tail256k||" "tail512k|MSK_TAIL_THRESHOLD=524288|" "tail1m|MSK_TAIL_THRESHOLD=1048576|" "tail2m|MSK_TAIL_THRESHOLD=2097152|" "b4m|MSK_BATCH_PATHS=4194304|"
} 2>&1 | tee gpurun_out/r02h_ab.txt
MSK_DEBUG_BOUNCES=1 python bench.py --workload vol --steps 1 --warmup 0 --no-cpu 2> gpurun_out/r02h_vol_bounces.txt > /dev/null
MSK_DEBUG_BOUNCES=1 python bench.py --workload c2 --steps 1 --warmup 0 --no-cpu 2> gpurun_out/r02h_c2_bounces.txt > /dev/null
grep -c "ran over" gpurun_out/r02h_vol_bounces.txt
