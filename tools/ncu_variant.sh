#!/bin/bash
# ncu --set full of one kernel for one library variant:  tools/ncu_variant.sh <lib.so> <kernel regex> <out name> [skip] [count]
lib=$1; k=$2; out=$3; skip=${4:-1}; cnt=${5:-1}
MSK_B200_LIB=$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -f -o gpurun_out/$out \
    python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_$out.log 2>&1
