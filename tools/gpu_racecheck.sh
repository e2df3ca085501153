#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards: k_sort, k_film_gather_tiled, BVH build) over the smoke render.
set -u
mkdir -p gpurun_out
timeout 10 compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck_smoke.txt python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
tail -2 gpurun_out/racecheck_smoke.txt
