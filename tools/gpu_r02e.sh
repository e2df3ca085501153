#!/bin/bash
# Round 2, GPU call E: graph replay / tiny-scene test, lockstep scheduling variants, shade prefetch.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render.py tests/test_gpu_intersect.py tests/test_aov.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
{
echo "== c2"; tools/variants.sh run c2 default nopref pat1 pat2 pat1t8 pat2t16 node2 node2pat1
echo "== c1"; tools/variants.sh run c1 default nopref
echo "== c1 MSK_GRAPH=0"; MSK_GRAPH=0 tools/variants.sh run c1 default
echo "== c1 MSK_STATIC_NODES=0"; MSK_STATIC_NODES=0 tools/variants.sh run c1 default
echo "== c3"; tools/variants.sh run c3 default nopref pat1 node2 node2pat1
echo "== c5"; tools/variants.sh run c5 default pat1 node2 node2pat1
} 2>&1 | tee gpurun_out/r02e_ab.txt
