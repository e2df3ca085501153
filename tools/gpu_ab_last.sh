#!/bin/bash
# One short A/B: default library vs launch-bounds variants of k_shade's glossy keys; parity tests on the winner.
set -u
mkdir -p gpurun_out
run() {
  MSK_B200_LIB=$2 timeout 40 python bench.py --workload c2 --steps 8 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 %.1f' % (d['value']/1e6), {k: round(v,2) for k,v in r['stage_ms'].items()})" | tee -a gpurun_out/ab_last.txt
}
: > gpurun_out/ab_last.txt
run default $PWD/misaki_render_b200/lib/libmisaki_b200.so
run glossy5 $PWD/build/variants/glossy5/libmisaki_b200.so
run glossy6 $PWD/build/variants/glossy6/libmisaki_b200.so
best=$(sort -k2 -n -r gpurun_out/ab_last.txt | head -1 | cut -d' ' -f1)
echo "best=$best"
if [ "$best" != default ]; then
  MSK_B200_LIB=$PWD/build/variants/$best/libmisaki_b200.so timeout 60 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee -a gpurun_out/ab_last.txt
fi
