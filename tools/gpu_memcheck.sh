#!/bin/bash
# compute-sanitizer memcheck over the smoke render (BVH build, wavefront path tracer, film gather, develop).
set -u
mkdir -p gpurun_out
timeout 36 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_smoke.txt python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
echo "rc=$?"; tail -3 gpurun_out/memcheck_smoke.txt
