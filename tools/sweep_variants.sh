#!/bin/bash
# Build (here, no GPU needed) variants of the library with different traversal settings into build/variants/.
#   tools/sweep_variants.sh build "2:12:8 0:0:8"     then on the GPU box:  tools/sweep_variants.sh run "..." c2
# variant = MSK_TRAVERSAL_MODE : MSK_TRI_THRESHOLD : MSK_TRAV_MIN_BLOCKS
set -u
cd "$(dirname "$0")/.."
mode=$1; shift
variants=$1; shift
if [ "$mode" = build ]; then
  for v in $variants; do
    IFS=: read m t b c <<< "$v"
    make -s -j8 -C misaki_render_b200/csrc OUT=../../build/variants/${v//:/_} EXTRA="-DMSK_TRAVERSAL_MODE=$m -DMSK_TRI_THRESHOLD=$t -DMSK_TRAV_MIN_BLOCKS=$b -DMSK_SORT_IN_COMMIT=${c:-0}" > /dev/null || exit 1
  done
else
  wl=${1:-c2}
  for v in $variants; do
    echo -n "$v  "
    MSK_B200_LIB=$PWD/build/variants/${v//:/_}/libmisaki_b200.so python bench.py --workload $wl --steps 5 --warmup 2 --no-cpu | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('Mpaths/s %.1f  ms/step %.2f  stages %s' % (d['value']/1e6, d['ms_per_step'], {k: round(v,2) for k,v in r['stage_ms'].items()}))"
  done
fi
