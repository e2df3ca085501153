#!/bin/bash
# A/B on the GPU box: tests with the default library, then bench stage times per variant library.
#   tools/ab.sh "old shade4 shade5" [workload]
set -u
wl=${2:-c2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
run() {
  echo -n "$1  "
  MSK_B200_LIB=$2 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('M/s %.1f  ms/step %.2f  stages %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v,2) for k,v in r.get('stage_ms', r.get('launch_ms')).items()}))"
}
run default $PWD/misaki_render_b200/lib/libmisaki_b200.so
for v in $1; do run $v $PWD/build/variants/$v/libmisaki_b200.so; done
