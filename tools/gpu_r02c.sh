#!/bin/bash
# Round 2, GPU call C: full GPU test suite with the tight bounds, the BASELINE-size parity tests and the multi-device path.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -150 > gpurun_out/pytest_gpu.log; tail -12 gpurun_out/pytest_gpu.log
{
echo "== c2"; tools/variants.sh run c2 precise default shade5
echo "== c3"; tools/variants.sh run c3 precise default
echo "== c1"; tools/variants.sh run c1 precise default
} 2>&1 | tee gpurun_out/r02c_ab.txt
