#!/usr/bin/env python3
"""Per-kernel DRAM traffic and time share from an ncu CSV of ONE bench step.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/traffic_c2.csv python bench.py --steps 1 --warmup 0 --no-cpu
    python tools/traffic.py gpurun_out/traffic_c2.csv c2        # updates profiles/roofline_traffic.json, prints shares

`traffic` in bench.py's roofline object = mean(dram read + write) per launch of the dominant kernel."""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def main():
    path, key = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    iid, iname, imet, iunit, ival = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    per = defaultdict(dict)
    names = {}
    for r in rows[1:]:
        val = float(r[ival].replace(",", ""))
        unit = r[iunit].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6, "second": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6}.get(unit, 1)
        per[r[iid]][r[imet]] = val * mult
        names[r[iid]] = r[iname]
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for i, m in per.items():
        short = re.sub(r"^.*::", "", re.sub(r"\(.*$", "", names[i]).replace("void ", ""))
        short = re.sub(r"<.*$", "", short)
        a = agg[short]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    out = {}
    print(f"{'kernel':28s} {'launches':>8s} {'time ms':>10s} {'share %':>8s} {'DRAM MB/launch':>15s} {'DRAM GB/s':>10s}")
    for k, (n, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:28s} {n:8d} {ns / 1e6:10.3f} {100 * ns / total:8.2f} {by / n / 1e6:15.2f} {by / max(ns, 1):10.1f}")
        out[k] = by / n
    if key == "c5":
        # launch order of `bench.py --workload c5 --steps 1 --warmup 0`: k_query_closest = [set-up primary, primary, secondary, e2e...],
        # k_query_any = [primary, secondary, e2e...]
        order = defaultdict(list)
        for i in sorted(per, key=int):
            if "k_query_closest<false>" in names[i].replace("(bool)0", "false") or "k_query_closest<0>" in names[i]:
                order["closest"].append(per[i])
            elif "k_query_any" in names[i]:
                order["any"].append(per[i])
        tot = lambda m: m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        if len(order["closest"]) >= 3 and len(order["any"]) >= 2:
            out.update(primary_closest=tot(order["closest"][1]), secondary_closest=tot(order["closest"][2]),
                       primary_any=tot(order["any"][0]), secondary_any=tot(order["any"][1]))
            for k in ("primary_closest", "secondary_closest", "primary_any", "secondary_any"):
                print(f"{k:28s} DRAM {out[k] / 1e6:10.1f} MB")
    f = ROOT / "profiles" / "roofline_traffic.json"
    data = json.loads(f.read_text()) if f.exists() else {}
    data[key] = out
    data.setdefault("_doc", "mean DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) per launch, per kernel, from one ncu pass over one bench step (tools/traffic.py)")
    f.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
