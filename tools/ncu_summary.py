#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed) into the text kept under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_k_intersect.ncu-rep [--ops]

Per launch: duration, DRAM bytes, L1/L2 hit rates, issue / pipe utilisation, warp execution efficiency,
occupancy, stall reasons; with --ops a per-opcode table from the SASS source page (instructions executed and
average active threads), which is what shows where divergence and pipe pressure come from.
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

RAW = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads / instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle"),
    ("SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed", "pipe XU %"),
    ("TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "pipe FMA-heavy %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU (active) %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU (active) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 LSU wavefronts %"),
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return out


def raw_summary(rep):
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')][:90]}  (launch id {r[hdr.index('ID')]})")
        for key, label in RAW:
            if key in hdr:
                print(f"   {label:34s} {r[hdr.index(key)]:>16s} {units[hdr.index(key)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        if stalls:
            print("   stalled warps per issue (top):   " + ", ".join(f"{n}={v:.2f}" for v, n in stalls[:6]))


def ops_summary(rep, top=28):
    text = ncu(rep, "source")
    # the page holds one table per launch: "Kernel Name",... line then the header line starting with "Address"
    blocks, cur = [], None
    for line in text.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = []
            blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    for bi, b in enumerate(blocks[:1]):
        rows = list(csv.reader(io.StringIO("\n".join(b))))
        hdr = rows[0]
        ci, ct, cs, csamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
        agg = defaultdict(lambda: [0, 0, 0, 0])
        tot_i = tot_t = tot_s = 0
        for r in rows[1:]:
            if len(r) <= ct:
                continue
            toks = r[cs].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            op = op.split(".")[0].rstrip(";")
            try:
                ni, nt, ns = int(r[ci]), int(r[ct]), int(r[csamp])
            except ValueError:
                continue
            a = agg[op]
            a[0] += ni; a[1] += nt; a[2] += ns; a[3] += 1
            tot_i += ni; tot_t += nt; tot_s += ns
        print(f"-- SASS opcode mix of launch {bi}: {tot_i} warp instructions, {tot_t / max(tot_i, 1):.2f} threads/instruction, {tot_s} stall samples")
        print(f"   {'op':10s} {'static':>6s} {'warp inst':>12s} {'%inst':>6s} {'thr/inst':>8s} {'%samples':>8s}")
        for op, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"   {op:10s} {a[3]:6d} {a[0]:12d} {100 * a[0] / max(tot_i, 1):6.2f} {a[1] / max(a[0], 1):8.2f} {100 * a[2] / max(tot_s, 1):8.2f}")


def lines_summary(rep, top=45):
    """Per CUDA source line (needs -lineinfo and --import-source on): warp instructions, threads per instruction and
    stall samples, for the first launch in the report."""
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    fname, func, hdr, seen_funcs = None, None, None, []
    agg = {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1]
            if func not in seen_funcs:
                seen_funcs.append(func)
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and len(r) > 10:
            if len(seen_funcs) > 1 and func != seen_funcs[0]:
                continue
            ci, ct, cs = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
            try:
                key = (fname, int(r[0]))
                a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
                a[0] += int(r[ci]); a[1] += int(r[ct]); a[2] += int(r[cs])
            except ValueError:
                pass
    ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
    print(f"-- per source line ({seen_funcs[0][:80] if seen_funcs else '?'}): {ti} warp instructions, {ts} samples")
    print(f"   {'file:line':28s} {'%inst':>6s} {'thr/inst':>8s} {'%samples':>8s}  source")
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"   {f + ':' + str(ln):28s} {100 * a[0] / max(ti, 1):6.2f} {a[1] / max(a[0], 1):8.2f} {100 * a[2] / max(ts, 1):8.2f}  {a[3]}")


if __name__ == "__main__":
    rep = sys.argv[1]
    raw_summary(rep)
    if "--ops" in sys.argv:
        ops_summary(rep)
    if "--lines" in sys.argv:
        lines_summary(rep)
