#!/bin/bash
# Round-end validation on one GPU: smoke, parity tests, default bench (both arms), ncu launch list of the same command.
set -u
mkdir -p gpurun_out
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 400 gpurun_out/bench_c2.json; tail -2 gpurun_out/bench_c2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tail -c 1 -f -o gpurun_out/prof_k_tail \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_k_tail.log 2>&1
ls -la gpurun_out | tail -5
