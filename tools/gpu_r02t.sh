#!/bin/bash
# Round 2, final measurement pass on the committed tree: smoke, ncu counters of one step of C1/C2/C3/C5 (bench.py's issue /
# DRAM rooflines), the ncu launch list of the default bench command, full captures of the top kernels, the default bench
# line + reference arm, and the C4 / fog lines.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python tools/ncu_counters.py run c1 c2 c3 c5
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu > gpurun_out/r02t_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 0 -c 1 -f -o gpurun_out/r02t_k_intersect_b0 python bench.py --one-step > gpurun_out/r02t_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 2 -c 1 -f -o gpurun_out/r02t_k_intersect_b1 python bench.py --one-step > gpurun_out/r02t_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 0 -c 1 -f -o gpurun_out/r02t_k_shadow_b0 python bench.py --one-step > gpurun_out/r02t_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 1 -f -o gpurun_out/r02t_k_shade_b0 python bench.py --one-step > gpurun_out/r02t_ncu4.log 2>&1
ls -la gpurun_out | grep r02t
