#!/usr/bin/env python3
"""bench.py -- paths/s of the B200 path-tracing backend on BASELINE.json's config, next to the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|c4]

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
  --impl reference   times that CPU restatement alone (the reference itself cannot be built here: DESIGN.md)
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
    film = torch.zeros((sd.height, sd.width, 5), dtype=torch.float32, device=dev)
    film_host = torch.zeros((sd.height, sd.width, 5), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step(collect):
        with torch.cuda.stream(ext):
            flush.zero_()  # evict the scene/BVH and queue tails from L2 between steps
            st = scene.render_dev(rd_rank, film.data_ptr())
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
            msk_dist.render_sharded(scene, rd_job, film, rank, world, host_out=film_host)

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
        stages = {
            "k_intersect": (st_t.ms_intersect, st_t.n_intersect_launches,
                            st_s.rays_closest * (BYTES_RAY_IN + BYTES_HIT_OUT) + st_s.nodes_closest * BYTES_NODE + st_s.tris_closest * BYTES_TRI),
            "k_shade": (st_t.ms_shade, st_t.n_shade_launches, st_s.shaded_vertices * BYTES_SHADE_VERTEX),
            "k_shadow": (st_t.ms_shadow, st_t.n_shadow_launches,
                         st_s.rays_shadow * (BYTES_RAY_IN + BYTES_OCC_OUT) + st_s.nodes_shadow * BYTES_NODE + st_s.tris_shadow * BYTES_TRI),
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
                    "film": st_t.ms_film, "step_with_timers": st_t.ms_render}
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
                       "l2": "256 MiB memset between steps (inside the timed region); path queues (~1.6 GB/batch) exceed the 126 MB L2",
                       "tris": int(info.ntris), "wide_nodes": int(info.nnodes), "bvh_build_ms": info.ms_build},
            "mrays_per_s": mrays, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "wall_s_timed_region": tm1 - tm0,
        }
        print(json.dumps(line), flush=True)
    scene.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU (development only)")
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds for the whole --impl reference run")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    args.steps = max(args.steps, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # bare `python bench.py --gpus N`: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29531"), str(Path(__file__).resolve())] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
