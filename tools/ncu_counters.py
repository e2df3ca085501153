#!/usr/bin/env python3
"""Per-kernel instruction and DRAM counters of exactly one step of each workload, for bench.py's issue / DRAM rooflines.

On the GPU box (one GPU):   python tools/ncu_counters.py run c1 c2 c3 c5      -> gpurun_out/ncu_counters_<wl>.csv
Here (no GPU needed):       python tools/ncu_counters.py collect c1 c2 c3 c5  -> profiles/ncu_counters.json

Each pass is `ncu --metrics ...` (no --set full: a handful of counters, 1-3 replays per launch) around
`bench.py --workload <wl> --one-step`, which renders one step and nothing else.  The counts (warp instructions, thread
instructions, DRAM bytes per kernel and step) are a property of the code and the seeded workload, not of the clock, so
bench.py may divide them by its own live CUDA-event times; the file records the hash of the CUDA sources it was taken on
and bench.py ignores it on any other tree."""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
METRICS = "smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"


def run(wls):
    from bench import source_sha
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "ncu_counters_sha.txt").write_text(source_sha())  # the tree the passes below run on
    for wl in wls:
        cmd = ["ncu", "--metrics", METRICS, "--clock-control", "none", "--csv", "--log-file", str(out / f"ncu_counters_{wl}.csv"),
               sys.executable, str(ROOT / "bench.py"), "--workload", wl, "--one-step"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
        print(wl, "rc", r.returncode, r.stderr[-300:] if r.returncode else "")


def to_num(x, unit):
    v = float(x.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "inst": 1, "": 1}
    return v * scale.get(unit, 1)


def collect(wls):
    sha = (ROOT / "gpurun_out" / "ncu_counters_sha.txt").read_text().strip()
    result = {"source_sha": sha, "metrics": METRICS, "how": "tools/ncu_counters.py (one step per workload under ncu --metrics, --clock-control none)",
              "workloads": {}}
    for wl in wls:
        f = ROOT / "gpurun_out" / f"ncu_counters_{wl}.csv"
        rows = [r for r in csv.reader(open(f)) if len(r) > 10]
        hdr = rows[0]
        iid, ik, im, iu, iv = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
        per_launch = defaultdict(dict)
        names = {}
        for r in rows[1:]:
            per_launch[int(r[iid])][r[im]] = to_num(r[iv], r[iu])
            names[int(r[iid])] = r[ik]
        agg = defaultdict(lambda: {"launches": 0, "inst_executed": 0.0, "thread_inst_executed": 0.0, "dram_bytes": 0.0, "ms_under_ncu": 0.0})
        seen_closest = 0
        for lid in sorted(per_launch):
            m = re.search(r"(k_[a-z_0-9]+)", names[lid])
            if not m:
                continue  # cub / library kernels of the BVH build
            k = m.group(1)
            if wl == "c5" and k == "k_query_closest":  # launch order of bench.py --one-step: primary set, then secondary
                k = "k_query_closest[%s]" % ("primary" if seen_closest == 0 else "secondary")
                seen_closest += 1
            a, v = agg[k], per_launch[lid]
            a["launches"] += 1
            a["inst_executed"] += v.get("smsp__inst_executed.sum", 0.0)
            a["thread_inst_executed"] += v.get("smsp__thread_inst_executed.sum", 0.0)
            a["dram_bytes"] += v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0)
            a["ms_under_ncu"] += v.get("gpu__time_duration.sum", 0.0)
        result["workloads"][wl] = dict(agg)
        tot = sum(a["ms_under_ncu"] for a in agg.values())
        print(f"{wl}: {sum(a['launches'] for a in agg.values())} launches, {tot:.2f} ms under ncu")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms_under_ncu"])[:8]:
            print(f"   {k:28s} {a['launches']:5d} launches  {a['ms_under_ncu']:9.3f} ms ({100 * a['ms_under_ncu'] / tot:5.1f} %)  "
                  f"{a['inst_executed'] / 1e6:10.1f} M warp inst  {a['thread_inst_executed'] / max(a['inst_executed'], 1):5.2f} thr/inst  {a['dram_bytes'] / 1e6:9.1f} MB DRAM")
    (ROOT / "profiles" / "ncu_counters.json").write_text(json.dumps(result, indent=1))


if __name__ == "__main__":
    mode, wls = sys.argv[1], sys.argv[2:] or ["c1", "c2", "c3", "c5"]
    (run if mode == "run" else collect)(wls)
