#!/bin/bash
# Static SASS size of the traversal kernels of a built library: instructions, S2UR / BSSY / local-memory counts.
# usage: tools/sass_count.sh [path/to/libmisaki_b200.so]
lib=${1:-misaki_render_b200/lib/libmisaki_b200.so}
cuobjdump -sass "$lib" | awk '
/Function :/ { name=$3; next }
/^ +\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\// { n[name]++; if ($0 ~ /S2UR|S2R/) s[name]++; if ($0 ~ /BSSY/) b[name]++; if ($0 ~ /STL|LDL/) l[name]++; if ($0 ~ /PRMT/) p[name]++ }
END { for (k in n) if (k ~ /k_intersect|k_shadow|k_query|k_tail/) printf "%-70s inst %5d  S2R %3d  BSSY %3d  LDL/STL %3d PRMT %3d\n", substr(k,40,70), n[k], s[k], b[k], l[k], p[k] }' | sort
