#!/bin/bash
# Round 2, session 3, GPU call J: compute-sanitizer memcheck + racecheck over a render that reaches the packet kernel, the
# lockstep driver with the shared-memory ray pool, the static driver and the tail kernel.
set -u
mkdir -p gpurun_out
python tools/sanitize_job.py 2>&1 | tail -1
timeout 150 compute-sanitizer --tool memcheck --log-file gpurun_out/r03j_memcheck.txt python tools/sanitize_job.py 2>&1 | tail -1
tail -2 gpurun_out/r03j_memcheck.txt
timeout 200 compute-sanitizer --tool racecheck --log-file gpurun_out/r03j_racecheck.txt python tools/sanitize_job.py 2>&1 | tail -1
tail -2 gpurun_out/r03j_racecheck.txt
