#!/bin/bash
# GPU session 2: new parity tests, C5 sweep bench, DRAM-traffic passes for C2 and C5
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -c 3500 gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_c2.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/traffic_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_query --csv --log-file gpurun_out/traffic_c5.csv \
    python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu > gpurun_out/traffic_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_query_closest -s 1 -c 2 -f -o gpurun_out/prof_c5_closest \
    python bench.py --workload c5 --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_c5.log 2>&1
ls -la gpurun_out
