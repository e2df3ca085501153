#!/usr/bin/env python3
"""bench.py -- paths/s of the B200 path-tracing backend on BASELINE.json's config, next to the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4|c5|vol]

A STEP is one complete render of the workload: every pixel x every sample of the job goes through
ray generation -> closest hit -> shade (NEE + BSDF sampling, MIS, Russian roulette) -> shadow rays -> film.
Default workload = BASELINE.json configs[1] (C2): bunny-class mesh with the Figure_2 GGX rough conductor,
512x512, 64 spp, unbounded depth with Russian roulette from depth 5 (scene synthesised by workloads/scenes.py:
the reference ships no meshes).

N > 1 (launched by torchrun, one rank per GPU): WEAK scaling -- every rank renders 64 samples per pixel of a
64*N-spp job (its own sample range, misaki_render_b200/distributed.py) and the films are summed into rank 0
with one NCCL reduce per step, inside the timed region.

Keys beyond the base contract:
  value      paths/s, device-timed (CUDA events on the library's stream), scene + BVH resident in HBM
  e2e        paths/s through msk_gpu_render with a HOST film buffer (D2H copy inside the timed region)
  roofline   the dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline  oracle (CPU restatement of the reference, all host threads) on a bounded sample of the job
  --impl reference   times that CPU restatement alone (the reference as a whole cannot be built here: DESIGN.md);
                     --ref-kind reference --workload c1 times the reference's own compiled render loop from oracle/_ref
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC, UNIT = "paths/s", "paths/s"
# algorithmic bytes per unit of work
BYTES_RAY_IN, BYTES_HIT_OUT, BYTES_OCC_OUT = 32, 20, 4
BYTES_NODE, BYTES_TRI = 80, 48  # this build's wide node / pre-gathered triangle (DESIGN.md "Data layout")
BYTES_SHADE_VERTEX = 400
BYTES_RAYGEN_SAMPLE = 112


def workload(name: str, world: int):
    """Returns (scene description, per-rank render-desc kwargs, human-readable name)."""
    from workloads import scenes
    if name == "c1":
        return scenes.cbox(256, 256), dict(spp=16, max_depth=5, rr_depth=5), "C1 Cornell box diffuse+area light 256x256 16spp depth5"
    if name == "c2":
        return (scenes.bunny(512, 512), dict(spp=64, max_depth=-1, rr_depth=5),
                "C2 bunny-class mesh (69316 tris) GGX rough conductor alpha=0.1 (Figure_2 material) 512x512 64spp max_depth=-1 rr_depth=5")
    if name == "c3":
        return (scenes.teapot(1024, 1024), dict(spp=256, max_depth=16, rr_depth=5),
                "C3 teapot-class mesh (150532 tris) GGX rough dielectric 1024x1024 256spp depth16")
    if name == "c4":
        return scenes.cbox(1920, 1080), dict(spp=4096, max_depth=5, rr_depth=5), "C4 Cornell box 1920x1080 4096spp depth5"
    if name == "vol":
        return (scenes.fog(512, 512, n=64), dict(spp=64, max_depth=-1, rr_depth=5, integrator="volpath"),
                "VOL volpath (SURVEY 8f rank 4): camera in thin fog, dielectric blob (49152 tris) with dense scattering interior, 512x512 64spp "
                "max_depth=-1 rr_depth=5")
    raise SystemExit(f"unknown workload {name}")


# --------------------------------------------------------------------------------------------- helpers
def physical_gpu_index(local_rank: int) -> str:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if local_rank < len(ids):
            return ids[local_rank]
    return str(local_rank)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons, sampled every 200 ms while the timed region runs."""
    FIELDS = ["clocks.sm", "clocks.max.sm", "power.draw", "clocks_event_reasons.hw_slowdown",
              "clocks_event_reasons.hw_thermal_slowdown", "clocks_event_reasons.sw_thermal_slowdown",
              "clocks_event_reasons.sw_power_cap"]

    def __init__(self, gpu: str):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", gpu, f"--query-gpu={','.join(self.FIELDS)}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.strip().split(",")]))

    def mark(self):
        return time.time()

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.25 and len(r) == len(self.FIELDS)] or \
               [r for _, r in self.rows if len(r) == len(self.FIELDS)]
        def num(x):
            try:
                return float(x)
            except ValueError:
                return None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def oracle_sample(sd, rd_kwargs, budget_s: float, nthreads: int = 0):
    """Time the CPU oracle on a bounded sample: the first `s` of the job's samples per pixel, s chosen from a
    1-sample calibration so the run takes about `budget_s` seconds.  Returns (paths/s, Mrays/s, info)."""
    from misaki_render_b200 import capi
    from oracle import pyoracle
    osc = pyoracle.OracleScene(sd)
    spp = rd_kwargs["spp"]
    osc.render(capi.render_desc(sample_begin=0, sample_end=1, **rd_kwargs), nthreads=nthreads)  # builds the BVH lazily
    _, st = osc.render(capi.render_desc(sample_begin=0, sample_end=1, **rd_kwargs), nthreads=nthreads)
    s = int(max(1, min(spp, round(budget_s / max(st.seconds, 1e-6)))))
    _, st = osc.render(capi.render_desc(sample_begin=0, sample_end=s, **rd_kwargs), nthreads=nthreads)
    osc.close()
    return st.paths / st.seconds, (st.rays_closest + st.rays_shadow) / st.seconds / 1e6, \
        dict(samples=s, seconds=st.seconds, threads=int(st.threads), paths=int(st.paths))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank: int):
    """The reference's CPU implementation of the path.  The reference itself needs Eigen/pugixml/TBB/Embree/
    OpenImageIO, none of which exist in this image (DESIGN.md "Reference buildability"), so this arm times
    the oracle restatement (kind "port") with every host thread, on a bounded sample per step."""
    if rank != 0:
        return
    from misaki_render_b200 import capi
    from oracle import pyoracle
    sd, rdk, wname = workload(args.workload, 1)
    if args.ref_kind == "reference":
        return run_reference_code(args, sd, rdk, wname)
    osc = pyoracle.OracleScene(sd)
    spp = rdk["spp"]
    osc.render(capi.render_desc(sample_begin=0, sample_end=1, **rdk))  # builds the oracle's BVH lazily
    _, st = osc.render(capi.render_desc(sample_begin=0, sample_end=1, **rdk))
    total_steps = args.steps + args.warmup
    s = int(max(1, min(spp, (args.ref_budget / total_steps) / max(st.seconds, 1e-6))))
    rd = capi.render_desc(sample_begin=0, sample_end=s, **rdk)
    for _ in range(args.warmup):
        osc.render(rd)
    secs, paths, rays = 0.0, 0, 0
    for _ in range(args.steps):
        _, st = osc.render(rd)
        secs += st.seconds; paths += st.paths; rays += st.rays_closest + st.rays_shadow
    value = paths / secs
    sample = f"first {s} of {spp} samples per pixel of every pixel ({sd.width}x{sd.height}), same seeds"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wname, "sample_per_step": sample},
        "mrays_per_s": rays / secs / 1e6,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": int(st.threads), "kind": "port", "sample": sample,
                         "cpu": cpu_model(),
                         "note": "CPU restatement of misaki's path (own SAH BVH + Moeller-Trumbore, std::thread over "
                                 "32x32 tiles); NOT TBB+Embree, which cannot be built in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)



def run_reference_code(args, sd, rdk, wname):
    """--ref-kind reference (C1 only): the reference's OWN compiled render loop from oracle/_ref (pyoracle.ReferenceLoop:
    integrator.cpp, path.cpp, scene.cpp, mesh / interaction, imageblock.cpp, hdrfilm.cpp, its colour textures) on the
    Cornell box of assets/cbox/scene.xml -- the one BASELINE config whose plugins the reference's build compiles.  Embree
    (brute force over the 36 triangles here), TBB (std::threads over the same tasks) and Eigen (oracle/ref_shim, scalar) are
    stand-ins, so this is the reference's CODE, not its performance with its real dependencies."""
    from oracle import pyoracle
    from workloads import scenes
    if args.workload != "c1":
        raise SystemExit("--ref-kind reference: only C1's plugins (diffuse, area, srgb) are compiled by the reference's build")
    if not pyoracle.REF_CODE_LIB.exists():
        raise SystemExit(f"{pyoracle.REF_CODE_LIB} is missing (python -c 'import __graft_entry__ as g; g.build()' where /root/reference exists)")
    loop = pyoracle.ReferenceLoop(sd, [r for _, r in scenes.CBOX_SHAPES],
                                  [(40, 40, 40) if n == "luminaire" else (-1, -1, -1) for n, _ in scenes.CBOX_SHAPES])
    threads, spp = os.cpu_count() or 1, rdk["spp"]
    _, dt1 = loop.render(1, threads)
    total_steps = args.steps + args.warmup
    s = int(max(1, min(spp, (args.ref_budget / total_steps) / max(dt1, 1e-6))))
    for _ in range(args.warmup):
        loop.render(s, threads)
    secs = sum(loop.render(s, threads)[1] for _ in range(args.steps))
    paths = sd.width * sd.height * s * args.steps
    value = paths / secs
    sample = f"{s} of {spp} samples per pixel of every pixel ({sd.width}x{sd.height}); unbounded depth, roulette from 5 (hard-wired, path.cpp:135-136)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wname, "sample_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample, "cpu": cpu_model(),
                         "note": "the reference's own integrator / path tracer / scene / film code compiled from /root/reference "
                                 "(oracle/Makefile.ref); stand-ins: brute-force intersector for Embree, std::thread for TBB, "
                                 "scalar oracle/ref_shim for Eigen"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- C5 intersection sweep
def c5_inputs(res: int, rank: int):
    """BASELINE configs[4]: the 9 998 244-triangle displaced sphere and its two ray sets."""
    from workloads import scenes
    sd = scenes.sphere10m()
    prim = scenes.primary_rays(sd, res)
    return sd, prim


def c5_oracle_sample(sd, sets, budget_s: float):
    """CPU arm of C5: the oracle's SAH BVH2 + Moeller-Trumbore (all host threads) on a strided sample of each ray
    set, closest hit and any hit.  Returns (Mrays/s, info)."""
    from oracle import pyoracle
    osc = pyoracle.OracleScene(sd)  # builds the BVH
    probe = {k: v[:: max(1, len(v) // 20000)] for k, v in sets.items()}
    t0 = time.perf_counter()
    for v in probe.values():
        osc.intersect(v)
    per_ray = (time.perf_counter() - t0) / sum(len(v) for v in probe.values())
    n_each = int(max(20000, min(min(len(v) for v in sets.values()), budget_s / (4 * per_ray))))
    rays_done, secs, detail = 0, 0.0, {}
    for k, v in sets.items():
        smp = np.ascontiguousarray(v[:: max(1, len(v) // n_each)][:n_each])
        t0 = time.perf_counter(); osc.intersect(smp); t1 = time.perf_counter(); osc.occluded(smp); t2 = time.perf_counter()
        detail[k] = {"closest_mrays_s": len(smp) / (t1 - t0) / 1e6, "any_mrays_s": len(smp) / (t2 - t1) / 1e6, "rays": len(smp)}
        rays_done += 2 * len(smp); secs += t2 - t0
    osc.close()
    return rays_done / secs / 1e6, dict(detail=detail, seconds=secs, rays=rays_done, threads=os.cpu_count())


def run_c5_reference(args, rank: int):
    if rank != 0:
        return
    sd, prim = c5_inputs(args.c5_res, 0)
    from oracle import pyoracle
    osc = pyoracle.OracleScene(sd)
    ph = osc.intersect(np.ascontiguousarray(prim[:: max(1, len(prim) // 400000)]))
    from workloads import scenes
    m = sd.meshes[0]
    sub = np.ascontiguousarray(prim[:: max(1, len(prim) // 400000)])
    sec = scenes.secondary_rays((m["verts"], m["tris"]), sub, ph, seed=0)
    osc.close()
    v, inf = c5_oracle_sample(sd, {"primary": sub, "secondary": sec}, args.ref_budget / (args.steps + args.warmup) * 1.0)
    sample = f"strided sample of {inf['rays'] // 4} rays of each set (primary from the {args.c5_res}^2 grid, secondary from their hits), closest + any hit"
    line = {"impl": "reference", "metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": inf["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": C5_NAME, "sample_per_step": sample}, "detail": inf["detail"],
            "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": inf["threads"], "kind": "port", "sample": sample, "cpu": cpu_model(),
                             "note": "oracle SAH BVH2 + Moeller-Trumbore, std::thread; NOT Embree"},
            "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


C5_NAME = "C5 10M-triangle displaced sphere (9998244 tris, one geomID): 4096^2 primary + incoherent cosine-hemisphere secondary rays, closest hit + any hit"


def run_c5(args, rank: int, local_rank: int, world: int):
    """Intersection sweep: one STEP = closest-hit and any-hit queries over both ray sets (4 launches).  Rays and
    results are resident in HBM for `value`; `e2e` goes through msk_gpu_intersect / msk_gpu_occluded with pinned
    host buffers.  N > 1: every rank traces the full ray sets against its own BVH replica (weak scaling, no
    collective -- the queries are independent)."""
    import torch
    import torch.distributed as dist
    from misaki_render_b200 import capi
    from workloads import scenes
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd, prim = c5_inputs(args.c5_res, rank)
    ctx = capi.Context(local_rank)
    t0 = time.time()
    scene = capi.Scene(ctx, sd)
    t_scene = time.time() - t0
    info = scene.accel_info()
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def to_dev(a):
        return torch.from_numpy(a.view(np.uint8).reshape(-1)).to(dev)

    d_prim = to_dev(prim)
    n_prim = len(prim)
    d_hits = torch.empty(n_prim * 20, dtype=torch.uint8, device=dev)
    d_occ = torch.empty(n_prim, dtype=torch.uint8, device=dev)
    with torch.cuda.stream(ext):
        scene.intersect_dev(d_prim.data_ptr(), d_hits.data_ptr(), n_prim)
    torch.cuda.synchronize()
    hits = d_hits.cpu().numpy().view(capi.HIT_DTYPE)
    m = sd.meshes[0]
    sec = scenes.secondary_rays((m["verts"], m["tris"]), prim, hits, seed=0)
    n_sec = len(sec)
    d_sec = to_dev(sec)
    sets = [("primary", d_prim, n_prim), ("secondary", d_sec, n_sec)]
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]

    def step(ev):
        with torch.cuda.stream(ext):
            k = 0
            for _, d_r, n in sets:
                if ev: ev[k].record(ext)
                scene.intersect_dev(d_r.data_ptr(), d_hits.data_ptr(), n); k += 1
                if ev: ev[k].record(ext)
                scene.occluded_dev(d_r.data_ptr(), d_occ.data_ptr(), n); k += 1
            if ev: ev[k].record(ext)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(None)
    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    barrier()
    tm0 = time.time()
    for i in range(args.steps):
        step(evs[i])
    barrier()
    tm1 = time.time()
    ms_total = sum(e[0].elapsed_time(e[4]) for e in evs)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(tm0, tm1) if sampler else None
    ms_step = ms_total / args.steps
    rays_step = 2 * (n_prim + n_sec) * world
    value = rays_step / (ms_step * 1e-3) / 1e6
    names = ["primary_closest", "primary_any", "secondary_closest", "secondary_any"]
    counts = [n_prim, n_prim, n_sec, n_sec]
    launch_ms = {nm: float(np.mean([e[i].elapsed_time(e[i + 1]) for e in evs])) for i, nm in enumerate(names)}
    launch_mrays = {nm: counts[i] / (launch_ms[nm] * 1e-3) / 1e6 for i, nm in enumerate(names)}

    # ---- end to end through the host-buffer entry points (H2D of the rays, D2H of the results inside)
    pin_r = torch.from_numpy(sec.view(np.uint8).reshape(-1)).pin_memory()
    pin_h = torch.empty(n_sec * 20, dtype=torch.uint8).pin_memory()
    pin_o = torch.empty(n_sec, dtype=torch.uint8).pin_memory()
    r_np, h_np, o_np = pin_r.numpy().view(capi.RAY_DTYPE), pin_h.numpy().view(capi.HIT_DTYPE), pin_o.numpy()
    lib = capi.load()

    def e2e_step():
        capi.check(lib, lib.msk_gpu_intersect(scene.handle, r_np.ctypes.data, h_np.ctypes.data, n_sec))
        capi.check(lib, lib.msk_gpu_occluded(scene.handle, r_np.ctypes.data, o_np.ctypes.data, n_sec))

    e2e_step()
    barrier()
    w0 = time.perf_counter()
    e2e_n = max(1, min(args.steps, 5))
    for _ in range(e2e_n):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - w0
    e2e = {"value": 2 * n_sec * world * e2e_n / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * n_sec * 32, "d2h_bytes_per_step": n_sec * 21,
           "ms_per_step": e2e_s / e2e_n * 1e3, "timer": "host wall clock",
           "note": f"secondary set only: msk_gpu_intersect + msk_gpu_occluded with pinned host rays/results; one-off scene upload + BVH build = {t_scene * 1e3:.0f} ms "
                   f"(BVH build {info.ms_build:.1f} ms on device)"}

    roofline = cpu = None
    if rank == 0:
        # traversal terms of the bytes model, counted by the instrumented kernel on a strided 1 Mi-ray sample of each set
        per = {}
        for nm, arr in (("primary", prim), ("secondary", sec)):
            smp = np.ascontiguousarray(arr[:: max(1, len(arr) // (1 << 20))])
            nn, nt = scene.intersect_stats(smp)
            per[nm] = {"nodes": float(nn.mean()), "tris": float(nt.mean())}
        top = max(("primary_closest", "secondary_closest"), key=lambda k: launch_ms[k])
        which = top.split("_")[0]
        n_top = n_prim if which == "primary" else n_sec
        bytes_ray = BYTES_RAY_IN + BYTES_HIT_OUT + per[which]["nodes"] * BYTES_NODE + per[which]["tris"] * BYTES_TRI
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = n_top * bytes_ray / (launch_ms[top] * 1e-3) / 1e9
        traffic = None
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get("c5", {}).get(top)
            except (ValueError, OSError):
                traffic = None
        roofline = {"bound": "hbm", "kernel": f"k_query_closest ({which} rays)", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "peak_source": peak_src, "launches_per_step": 1, "avg_launch_ms": launch_ms[top],
                    "algorithmic_bytes_per_launch": n_top * bytes_ray, "per_ray": per,
                    "bytes_model": "SURVEY 8(d): 32 B ray + 20 B hit + visited wide nodes x 80 B + tested triangles x 48 B per closest-hit ray "
                                   "(counted on a strided 1 Mi-ray sample); BVH + triangles = "
                                   f"{(info.node_bytes + info.tri_bytes) / 1e6:.0f} MB > 126 MB L2",
                    "launch_ms": launch_ms, "launch_mrays_per_s": launch_mrays}
        if world == 1 and not args.no_cpu:
            v, inf = c5_oracle_sample(sd, {"primary": prim, "secondary": sec}, args.cpu_budget)
            cpu = {"value": v, "unit": "Mrays/s", "cores": inf["threads"], "kind": "port", "cpu": cpu_model(), "detail": inf["detail"],
                   "sample": f"strided sample of {inf['rays'] // 4} rays of each set, closest + any hit ({inf['seconds']:.1f} s)",
                   "note": "oracle SAH BVH2 + Moeller-Trumbore over all host threads; NOT Embree"}
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": C5_NAME, "rays_primary": n_prim, "rays_secondary": n_sec, "tris": int(info.ntris), "wide_nodes": int(info.nnodes),
                           "bvh_bytes": int(info.node_bytes + info.tri_bytes), "bvh_build_ms": info.ms_build, "sah_cost": info.sah_cost,
                           "l2": "ray sets (0.5 GB each) and the BVH (0.6 GB) exceed the 126 MB L2; no flush needed",
                           "partition": "replicated BVH, every rank traces the full ray sets (independent queries, no collective)"},
                "e2e": e2e, "gpu_launches": 4 * args.steps, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "wall_s_timed_region": tm1 - tm0}
        print(json.dumps(line), flush=True)
    # release every torch tensor that was used on the library's stream BEFORE that stream is destroyed: the caching
    # allocator records an event on each stream a block was used on when the block is freed
    d_prim = d_sec = d_hits = d_occ = sets = pin_r = pin_h = pin_o = r_np = h_np = o_np = None
    barrier()
    if world > 1:
        dist.destroy_process_group()
    torch.cuda.synchronize()
    scene.close()
    ctx.close()


# --------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from misaki_render_b200 import capi, distributed as msk_dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this backend has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    sd, rdk, wname = workload(args.workload, world)
    if args.spp:
        rdk["spp"] = args.spp
    spp_rank = rdk["spp"]
    job = dict(rdk, spp=spp_rank * world)  # weak scaling: the job grows with N, each rank renders spp_rank samples
    rd_job = capi.render_desc(**job)
    rd_rank = msk_dist.shard_desc(rd_job, rank, world)
    npix = sd.width * sd.height
    paths_per_step = npix * spp_rank * world

    ctx = capi.Context(local_rank)
    t0 = time.time()
    scene = capi.Scene(ctx, sd)
    t_scene = time.time() - t0
    info = scene.accel_info()
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    # N > 1: the film lives in CUDA-IPC-exportable memory and is reduced by the library's own peer-memory kernel
    # (csrc/msk_peer.cu); --reduce nccl selects torch.distributed's reduce instead
    peer = None
    if world > 1 and args.reduce == "peer":
        # every rank must end up on the same reduction: agree on whether the IPC set-up succeeded everywhere
        err = None
        try:
            peer = msk_dist.PeerFilm(ctx, (sd.height, sd.width, 5), rank, world)
        except Exception as e:  # noqa: BLE001 -- e.g. no peer access between two of the devices
            err, peer = e, None
        ok = torch.tensor([0 if peer is None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            if err is not None:
                print(f"bench.py rank {rank}: peer-memory film reduction unavailable ({err}); all ranks use ncclReduce", file=sys.stderr)
            if peer is not None:
                peer.close()
            peer = None
    if peer is not None:
        film = peer.tensor(dev)
    else:
        film = torch.zeros((sd.height, sd.width, 5), dtype=torch.float32, device=dev)
    film_host = torch.zeros((sd.height, sd.width, 5), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(collect):
        with torch.cuda.stream(ext):
            flush.zero_()  # evict the scene/BVH and queue tails from L2 between steps
            st = scene.render_dev(rd_rank, film.data_ptr())
            if peer is not None:
                peer.reduce()
            else:
                msk_dist.reduce_film(film, 0)
        if collect is not None:
            collect.append(st)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step(None)
    sampler = ClockSampler(physical_gpu_index(local_rank)) if rank == 0 else None
    barrier()
    stats = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tm0 = time.time()
    ev0.record(ext)
    for _ in range(args.steps):
        step(stats)
    ev1.record(ext)
    barrier()
    tm1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop(tm0, tm1) if sampler else None
    ms_step = ms_total / args.steps
    value = paths_per_step / (ms_step * 1e-3)
    rays_rank = sum(s.rays_closest + s.rays_shadow for s in stats) / args.steps
    rays_t = torch.tensor([rays_rank], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rays_t, op=dist.ReduceOp.SUM)
    mrays = float(rays_t.item()) / (ms_step * 1e-3) / 1e6
    launches = int(sum(s.kernel_launches for s in stats))

    # ---- end to end: the public host-buffer entry point, D2H of the film inside the timed region
    e2e_steps = args.steps
    fh = film_host.numpy()

    def e2e_step():
        if world == 1:
            scene.render(rd_rank, film=fh)  # msk_gpu_render: host film in/out
        else:
            msk_dist.render_sharded(scene, rd_job, film, rank, world, host_out=film_host, peer=peer)

    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - w0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    film_bytes = npix * 5 * 4
    e2e = {"value": paths_per_step * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(capi.C.sizeof(capi.MskRenderDesc)),
           "d2h_bytes_per_step": film_bytes, "ms_per_step": e2e_s / e2e_steps * 1e3, "timer": "host wall clock around the call",
           "note": "scene + BVH stay resident across steps (as Embree's scene does across Integrator::render calls); "
                   f"one-off scene upload + BVH build = {t_scene * 1e3:.1f} ms (BVH build {info.ms_build:.2f} ms)"}

    # ---- per-kernel roofline: one extra step with stage timers, one with the instrumented traversal
    roofline = None
    if rank == 0:
        rd_t = msk_dist.shard_desc(rd_job, rank, world); rd_t.flags = capi.RENDER_STAGE_TIMERS
        with torch.cuda.stream(ext):
            flush.zero_()
            st_t = scene.render_dev(rd_t, film.data_ptr())
        rd_s = msk_dist.shard_desc(rd_job, rank, world); rd_s.flags = capi.RENDER_TRAVERSAL_STATS
        st_s = scene.render_dev(rd_s, film.data_ptr())
        nodes_c, tris_c = st_s.nodes_closest / max(st_s.rays_closest, 1), st_s.tris_closest / max(st_s.rays_closest, 1)
        nodes_s, tris_s = st_s.nodes_shadow / max(st_s.rays_shadow, 1), st_s.tris_shadow / max(st_s.rays_shadow, 1)
        # rays finished by the per-path tail kernel (k_tail, timed as its own stage) do not belong to the wavefront launches
        wf_c = 1.0 - st_t.tail_rays_closest / max(st_t.rays_closest, 1)
        wf_s = 1.0 - st_t.tail_rays_shadow / max(st_t.rays_shadow, 1)
        stages = {
            "k_intersect": (st_t.ms_intersect, st_t.n_intersect_launches,
                            wf_c * (st_s.rays_closest * (BYTES_RAY_IN + BYTES_HIT_OUT) + st_s.nodes_closest * BYTES_NODE + st_s.tris_closest * BYTES_TRI)),
            "k_shade": (st_t.ms_shade, st_t.n_shade_launches, wf_c * st_s.shaded_vertices * BYTES_SHADE_VERTEX),
            "k_shadow": (st_t.ms_shadow, st_t.n_shadow_launches,
                         wf_s * (st_s.rays_shadow * (BYTES_RAY_IN + BYTES_OCC_OUT) + st_s.nodes_shadow * BYTES_NODE + st_s.tris_shadow * BYTES_TRI)),
        }
        peaks_file = ROOT / "MEASURED_PEAKS.json"
        if peaks_file.exists():
            peak, peak_src = float(json.loads(peaks_file.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        top = max(stages, key=lambda k: stages[k][0])
        ms, nl, nbytes = stages[top]
        achieved = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        traffic = None
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists():
            try:
                traffic = json.loads(tf.read_text()).get(args.workload, {}).get(top)
            except (ValueError, OSError):
                traffic = None
        stage_ms = {"raygen": st_t.ms_raygen, "intersect": st_t.ms_intersect, "sort": st_t.ms_sort, "shade": st_t.ms_shade, "shadow": st_t.ms_shadow,
                    "film": st_t.ms_film, "tail": st_t.ms_tail, "step_with_timers": st_t.ms_render}
        roofline = {
            "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src,
            "launches_per_step": int(nl), "avg_launch_ms": ms / max(nl, 1), "algorithmic_bytes_per_launch": nbytes / max(nl, 1),
            "bytes_model": "SURVEY 8(d): per closest-hit ray 32 B in + 20 B out + visited wide nodes x 80 B + tested triangles x 48 B "
                           "(counted by the instrumented kernel on the same rays); any-hit 32 + 4; shaded vertex 400 B. "
                           "Node/triangle bytes are served mostly from L2 for this L2-resident scene, so this is an "
                           "algorithmic-traffic figure, not DRAM traffic (see `traffic`).",
            "per_ray": {"nodes_closest": nodes_c, "tris_closest": tris_c, "nodes_shadow": nodes_s, "tris_shadow": tris_s},
            "stage_ms": stage_ms,
            "tail": {"rays_closest": int(st_t.tail_rays_closest), "rays_shadow": int(st_t.tail_rays_shadow), "launches": int(st_t.n_tail_launches),
                     "note": "k_tail: one launch runs every path still alive once the queue is short (MSK_TAIL_THRESHOLD rays) to completion"},
            "stage_gbs": {k: (v[2] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0) for k, v in stages.items()},
        }

    # ---- CPU baseline (rank 0, N = 1 only): oracle on a bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, mr, inf = oracle_sample(sd, rdk, args.cpu_budget)
        cpu = {"value": v, "unit": UNIT, "cores": inf["threads"], "kind": "port", "mrays_per_s": mr, "cpu": cpu_model(),
               "sample": f"first {inf['samples']} of {rdk['spp']} samples per pixel, all {npix} pixels, same seeds "
                         f"({inf['paths']} paths, {inf['seconds']:.1f} s)",
               "note": "oracle/ CPU restatement (own SAH BVH + Moeller-Trumbore, std::thread tiles); not TBB+Embree"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": wname, "spp_per_gpu": spp_rank, "job_spp": spp_rank * world, "partition": "sample ranges + 1 film reduce/step",
                       "film_reduce": "none (1 GPU)" if world == 1 else ("msk_gpu_reduce_film: one kernel pulling peer films over NVLink (CUDA IPC)" if peer is not None else "ncclReduce"),
                       "l2": "256 MiB memset between steps (inside the timed region); path queues (~1.6 GB/batch) exceed the 126 MB L2",
                       "tris": int(info.ntris), "wide_nodes": int(info.nnodes), "bvh_build_ms": info.ms_build},
            "mrays_per_s": mrays, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": tm1 - tm0,
        }
        print(json.dumps(line), flush=True)
    # release every torch tensor that was used on the library's stream (NCCL's record_stream included) BEFORE that
    # stream is destroyed: the caching allocator records an event on each such stream when the block is freed
    if peer is not None:
        peer.check()
    film = film_host = flush = fh = None
    barrier()
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    scene.close()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4", "c5", "vol"])
    ap.add_argument("--c5-res", type=int, default=4096, help="C5: primary rays are a res x res pinhole grid")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU (development only)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds for the whole --impl reference run")
    ap.add_argument("--ref-kind", default="port", choices=["port", "reference"],
                    help="--impl reference: the oracle port (every workload) or, for C1, the reference's own compiled code from oracle/_ref")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--reduce", default="peer", choices=["peer", "nccl"], help="N > 1: film reduction by the library's NVLink peer kernel or by NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        (run_c5_reference if args.workload == "c5" else run_reference)(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # bare `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    (run_c5 if args.workload == "c5" else run_ours)(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
