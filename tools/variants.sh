#!/bin/bash
# Build (here, no GPU needed) named variants of the CUDA library into build/variants/<name>/:
#   tools/variants.sh build name1="-DMSK_REFILL_MIN=4" name2="-DMSK_SMEM_STACK=0 -DMSK_PERM_LUT=0" ...
# and, on the GPU box, bench each of them (device-timed stage times):
#   tools/variants.sh run c2 name1 name2 ...
set -u
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    make -s -j8 -C misaki_render_b200/csrc OUT=../../build/variants/$name EXTRA="$flags" > /dev/null || { echo "build of $name failed"; cat build/variants/$name/*.log | grep -i error; exit 1; }
    echo "built $name ($flags)"
  done
else
  wl=$1; shift
  for name in "$@"; do
    lib=$PWD/build/variants/$name/libmisaki_b200.so
    [ "$name" = default ] && lib=$PWD/misaki_render_b200/lib/libmisaki_b200.so
    echo -n "$name  "
    MSK_B200_LIB=$lib timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
st=r.get('stage_ms', r.get('launch_ms'))
print('M/s %.1f  ms/step %.2f  stages %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v,2) for k,v in st.items()}))" || echo failed
  done
fi
