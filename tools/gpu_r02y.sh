#!/bin/bash
# Round 2, final pass on the committed tree (after the film-gather and any-hit changes): full GPU suite, smoke, ncu
# counters of one step of C1/C2/C3/C5, the ncu launch list of the bench command, full captures of the top kernels, the
# default bench line + reference arm, C3 / C4 / fog lines.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python tools/ncu_counters.py run c1 c2 c3 c5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02y_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu > gpurun_out/r02y_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 0 -c 1 -f -o gpurun_out/r02y_k_intersect_b0 python bench.py --one-step > gpurun_out/r02y_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 2 -c 1 -f -o gpurun_out/r02y_k_intersect_b1 python bench.py --one-step > gpurun_out/r02y_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 0 -c 1 -f -o gpurun_out/r02y_k_shadow_b0 python bench.py --one-step > gpurun_out/r02y_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 1 -f -o gpurun_out/r02y_k_shade_b0 python bench.py --one-step > gpurun_out/r02y_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_film_gather -s 0 -c 1 -f -o gpurun_out/r02y_k_film_gather python bench.py --one-step > gpurun_out/r02y_ncu5.log 2>&1
