"""bench.py's roofline bookkeeping (no GPU): which bound a kernel reports, and that nothing is stated without counters taken on
the current CUDA sources (the line once carried an HBM fraction for an issue-bound kernel and a 40-commit-old
traffic figure)."""
import json
import warnings
from pathlib import Path

import bench

ROOT = Path(__file__).resolve().parent.parent


def test_roofline_without_counters_states_no_fraction():
    r = bench.kernel_roofline("k_intersect", ms=5.0, launches=6, alg_bytes=25e9, counters=None, counters_note="none")
    assert r["bound"] == "issue" and r["frac"] is None and r["achieved"] is None and r["traffic"] is None
    assert r["hbm"]["dram_frac"] is None and r["issue"]["frac"] is None
    assert r["hbm"]["algorithmic_gbs"] == 25e9 / 5e-3 / 1e9  # the SURVEY 8(d) model stays available, labelled as such


def test_roofline_picks_the_larger_measured_fraction():
    hbm_peak, _, mhz = bench.peaks()
    issue_peak = 148 * bench.SM_ISSUE_SLOTS_PER_CLK * mhz * 1e6 / 1e9
    ms = 2.0
    issue_bound = {"k": {"inst_executed": 0.7 * issue_peak * 1e9 * ms * 1e-3, "thread_inst_executed": 0.7 * issue_peak * 1e9 * ms * 1e-3 * 20,
                         "dram_bytes": 0.05 * hbm_peak * 1e9 * ms * 1e-3, "launches": 3}}
    r = bench.kernel_roofline("k", ms, 3, 1e9, issue_bound, None)
    assert r["bound"] == "issue" and abs(r["frac"] - 0.7) < 1e-9 and abs(r["issue"]["threads_per_inst"] - 20) < 1e-9
    assert abs(r["hbm"]["dram_frac"] - 0.05) < 1e-9 and r["unit"] == "Gwarp-inst/s"
    dram_bound = {"k": dict(issue_bound["k"], dram_bytes=0.9 * hbm_peak * 1e9 * ms * 1e-3)}
    r = bench.kernel_roofline("k", ms, 3, 1e9, dram_bound, None)
    assert r["bound"] == "hbm" and abs(r["frac"] - 0.9) < 1e-9 and r["unit"] == "GB/s"
    assert abs(r["traffic"] - 0.9 * hbm_peak * 1e9 * ms * 1e-3 / 3) < 1.0  # DRAM bytes per launch, from the same counters


def test_counters_are_used_only_on_the_tree_they_were_taken_on(tmp_path, monkeypatch):
    d = json.loads((ROOT / "profiles" / "ncu_counters.json").read_text())
    assert set(d["workloads"]) >= {"c1", "c2", "c3", "c5"}
    for wl in ("c2", "c3"):
        assert {"k_intersect", "k_shadow", "k_shade"} <= set(d["workloads"][wl])
    if d["source_sha"] != bench.source_sha():  # legitimate while kernels are being edited; the bench line then states no fractions
        warnings.warn("profiles/ncu_counters.json was taken on other CUDA sources: run tools/ncu_counters.py on the GPU box")
        assert bench.load_counters("c2")[0] is None
    else:
        assert bench.load_counters("c2")[0] is not None
    monkeypatch.setattr(bench, "source_sha", lambda: "0" * 16)
    c, note = bench.load_counters("c2")
    assert c is None and "another source tree" in note
