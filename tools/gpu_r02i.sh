#!/bin/bash
# Round 2, GPU call I: batches in flight (MSK_INFLIGHT lanes) -- full GPU suite, then A/B per workload.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; SKIP_TESTS=1 tools/ab_knobs.sh c2 "lanes1|MSK_INFLIGHT=1|" "lanes2|MSK_INFLIGHT=2|" "lanes3|MSK_INFLIGHT=3|" "lanes4|MSK_INFLIGHT=4|" "lanes4_min2m|MSK_INFLIGHT=4 MSK_SPLIT_MIN_PATHS=2097152|" "lanes2_nosplit|MSK_INFLIGHT=2 MSK_SPLIT_MIN_PATHS=1000000000|"
echo "== c3"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh c3 "lanes1|MSK_INFLIGHT=1|" "lanes2|MSK_INFLIGHT=2|" "lanes3|MSK_INFLIGHT=3|" "lanes4_b16m|MSK_INFLIGHT=4 MSK_BATCH_PATHS=16777216|" "lanes2_ploc|MSK_INFLIGHT=2 MSK_BVH_BUILDER=ploc|"
echo "== vol"; SKIP_TESTS=1 STEPS=3 tools/ab_knobs.sh vol "lanes1|MSK_INFLIGHT=1|" "lanes2|MSK_INFLIGHT=2|" "lanes4|MSK_INFLIGHT=4|" "lanes2_tail1m|MSK_INFLIGHT=2 MSK_TAIL_THRESHOLD=1048576|"
echo "== c1"; SKIP_TESTS=1 tools/ab_knobs.sh c1 "default||"
echo "== c4"; SKIP_TESTS=1 STEPS=1 tools/ab_knobs.sh c4 "lanes1|MSK_INFLIGHT=1|" "lanes2|MSK_INFLIGHT=2|"
} 2>&1 | tee gpurun_out/r02i_ab.txt
