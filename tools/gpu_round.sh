#!/bin/bash
# One GPU-box session: parity tests, bench, ncu launch list, ncu full captures of the three hot kernels.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt; lscpu | grep "Model name" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 3000 gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
if [ "${1:-}" = "ncu" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_launch_bench.log 2>&1
  for k in k_intersect k_shade k_shadow; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 2 -f -o gpurun_out/prof_$k \
        python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_$k.log 2>&1
  done
  ls -la gpurun_out
fi
