#!/bin/bash
# A/B on the GPU box over runtime knobs (environment) and library variants of ONE snapshot:
#   tools/ab_knobs.sh [workload] "name|ENV=1 ENV2=2|variant-or-empty" ...
# GPU parity tests run first with the default library; then one bench line per configuration.
set -u
wl=${1:-c2}; shift
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
fi
for cfg in "$@"; do
  IFS='|' read -r name envs variant <<< "$cfg"
  lib=$PWD/misaki_render_b200/lib/libmisaki_b200.so
  [ -n "${variant:-}" ] && lib=$PWD/build/variants/$variant/libmisaki_b200.so
  printf '%-28s' "$name"
  env $envs MSK_B200_LIB=$lib timeout ${BENCH_TIMEOUT:-240} python bench.py --workload $wl --steps ${STEPS:-8} --warmup 3 --no-cpu 2> gpurun_out/ab_$name.err | tee gpurun_out/ab_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('M/s %.1f  ms/step %.2f  launches %d  stages %s' % (d['value']/1e6, d['ms_per_step'], d.get('gpu_launches',0), {k: round(v,2) for k,v in r.get('stage_ms', r.get('launch_ms')).items()}))" || tail -3 gpurun_out/ab_$name.err
done
