#!/bin/bash
# Round 2, session 3, final pass on the committed tree (ray pool, shadow tag, camera-ray packets): full GPU suite, smoke, ncu counters of one
# step of C1/C2/C3/C5, the ncu launch list of the bench command, full captures of the top kernels, the default bench line +
# the reference arm, C3 / C4 / fog lines.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python tools/ncu_counters.py run c1 c2 c3 c5
python tools/ncu_counters.py collect c1 c2 c3 c5 > gpurun_out/r03h_counters.txt 2>&1   # so that the bench lines below carry the issue / DRAM rooflines of this tree
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03h_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-sub --no-cpu > gpurun_out/r03h_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 2 -c 1 -f -o gpurun_out/r03h_k_intersect_b1 python bench.py --one-step > gpurun_out/r03h_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_intersect -s 0 -c 1 -f -o gpurun_out/r03h_k_intersect_packet python bench.py --one-step > gpurun_out/r03h_ncu0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shadow -s 0 -c 1 -f -o gpurun_out/r03h_k_shadow_b0 python bench.py --one-step > gpurun_out/r03h_ncu3.log 2>&1
timeout 900 python bench.py > gpurun_out/r03h_bench.json 2> gpurun_out/r03h_bench.err; tail -c 300 gpurun_out/r03h_bench.json; tail -2 gpurun_out/r03h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r03h_bench_reference.json 2> gpurun_out/r03h_bench_reference.err; tail -c 200 gpurun_out/r03h_bench_reference.json
for wl in c3 c4 vol; do timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/r03h_bench_$wl.json 2> gpurun_out/r03h_bench_$wl.err; head -c 200 gpurun_out/r03h_bench_$wl.json; echo; done
