#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of one bench step for a library variant
lib=$1; out=$2
MSK_B200_LIB=$lib timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/$out.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/$out.log 2>&1
