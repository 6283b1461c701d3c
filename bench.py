#!/usr/bin/env python3
"""bench.py -- headline benchmark of the raytrace path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config 2]

One "step" = one pass of the hot path over one frame of the BASELINE.json config-2 workload (6x6 tessellated-sphere grid
+ floor = 101 090 triangles, 1920x1080, 1 light, primary + shadow + 1 diffuse bounce, 1 sample/pixel): every pixel-sample
("ray" in the reference's own accounting, raytrace.c:545) traced and shaded into three 16-bit planes.

  value      Mrays/s with scene, acceleration lists and framebuffer already resident in HBM, device-timed with CUDA
             events on the launch stream, L2 flushed before every timed step, max over ranks.
  e2e        the same metric through the reference-facing call RaytraceAll (C-ABI, HOST buffers in and out): upload +
             repack + trace + read-back all inside the timed region.
  roofline   algorithmic bytes (SURVEY.md section 8d formula, event counts from the counting build of the kernel) per
             launch / kernel time, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own C build of the kernel (oracle/_ref, "reference") or the C port ("port") on the host
             cores, on a bounded band of rows of the same frame.

N > 1 (launched by torchrun, one process per GPU): the frame is cut into row bands dealt round-robin to the ranks, the
scene is replicated, and the planes are assembled with one NCCL all-gather over NVLink (the path's only exchange step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

HBM_FALLBACK_GBS = 6650.0     # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--variant", type=int, default=-1)
    ap.add_argument("--band-rows", type=int, default=16)
    return ap.parse_args()


def build_workload(cfg_id):
    from opencl_render_b200 import api, scenes
    cfg = scenes.CONFIGS[cfg_id]
    sc = cfg["make"]()
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
    lists = api.camera_triangle_list(cam, sc)
    api.scene_triangle_list(sc, 256)
    return cfg, sc, cam, lists


def algorithmic_bytes(cnt: dict, rays: int) -> float:
    """B = 8*[primary ray] + 68*C_prim + 8*K_cells + 68*C_grid + 212*H + 44*O + 12 per pixel-sample, counted
    in the reference layout; totals per launch."""
    grid_candidates = cnt["gridCandidates"] + cnt.get("mailboxSkips", 0)     # the reference tests these again in every cell
    return (8.0 * rays + 68.0 * cnt["primCandidates"] + 8.0 * cnt["cells"] + 68.0 * grid_candidates + 212.0 * cnt["shadedHits"] +
            44.0 * cnt["occluderLookups"] + 12.0 * rays)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.index = index
        self.source = "nvidia-smi"
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        if self._run_nvml():
            return
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def _run_nvml(self) -> bool:
        """The same counters nvidia-smi prints, read through NVML directly: a query takes well under a millisecond, so a timed region
        of ~100 ms yields dozens of samples instead of one.  False when NVML is not usable (then nvidia-smi is polled)."""
        try:
            import pynvml
            pynvml.nvmlInit()
            index = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:      # CUDA's device index -> NVML's
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    index = int(ids[index])
            dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(dev, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM)
            pynvml.nvmlDeviceGetCurrentClocksEventReasons(dev)
        except Exception:
            return False
        bits = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}
        self.source = "nvml"
        while not self._stop.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM)))
                r = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(dev))
                for n, b in bits.items():
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.002)
        return True

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path (its kernel compiled as C, oracle/_ref; the C port
    when that build did not travel) on all host cores, on a bounded band of rows of the same frame per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cfg, sc, cam, lists = build_workload(args.config)
    cores = os.cpu_count() or 1
    kind = "port"
    try:
        import ref
        if ref.LIB.is_file() or ref.available():
            ref.load()
            kind = "reference"
            render = lambda rows: ref.render(cam, lists, sc, cfg["samples"], threads=cores, rows=rows, row_step=4)
    except Exception:
        kind = "port"
    if kind == "port":
        import port
        render = lambda rows: port.render(cam, lists, sc, cfg["samples"], threads=cores, rows=rows, row_step=4)
    h, w = cam.height, cam.width
    step_rows = 4                                # every 4th row of the WHOLE frame: a quarter of the work, sky and geometry in proportion
    rows = (0, h)
    n_rows = len(range(0, h, step_rows))
    for _ in range(args.warmup):
        render(rows)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        render(rows)
    dt = time.perf_counter() - t0
    rays = n_rows * w * cfg["samples"]
    value = rays * args.steps / dt / 1e6
    sample = f"every {step_rows}th row of all {h} ({rays} pixel-samples) per step"
    print(json.dumps({
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": cfg["name"], "triangles": sc.triangle_count, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_b200(args):
    # stdout carries ONE JSON line: everything libraries print there (NCCL's version banner, ...) is sent to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from opencl_render_b200 import api, dist as odist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the library has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg, sc, cam, lists = build_workload(args.config)
    h, w, S = cam.height, cam.width, cfg["samples"]
    rays_frame = h * w * S

    ds = api.DeviceScene(sc, local)
    fr = api.DeviceFrame(ds, cam, lists)
    part = odist.BandPartition(h, w, rank, world, args.band_rows)
    gather, gather_kind = odist.plane_exchange(fr, part, torch.device("cuda", local)) if world > 1 else (None, None)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step(timed_events=None):
        flush.zero_()                                        # L2 flush (256 MiB > 126 MB L2), outside the timed events
        if timed_events is not None:
            timed_events[0].record()
        launches = part.render(fr, S, args.variant, stream)
        if timed_events is not None:
            timed_events[2].record()                         # trace done, gather not yet: "gather excluded" figure
        if gather is not None:
            gather.run()
            launches += gather.launches
        if timed_events is not None:
            timed_events[1].record()
        return launches

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    events = [tuple(torch.cuda.Event(enable_timing=True) for _ in range(3)) for _ in range(args.steps)]
    launches = 0
    with ClockSampler(local) as clocks:
        for k in range(args.steps):
            launches += step(events[k])
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms_total = sum(a.elapsed_time(b) for a, b, _ in events)
    ms_render = sum(a.elapsed_time(c) for a, _, c in events)
    t = torch.tensor([ms_total, ms_render], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_render = float(t[0].item()), float(t[1].item())
    ms_step = ms_total / args.steps
    value = rays_frame / ms_step / 1e3

    # ---- end to end through the drop-in call (host buffers; rank-local rows; pinned host arrays) ---------------------------
    e2e = odist.EndToEnd(sc, cam, lists, part, local)
    for _ in range(2):
        e2e.step(S)
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e.step(S)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = rays_frame * e2e_steps / float(t.item()) / 1e6

    # ---- roofline of the dominant kernel, wf_pipe_kernel (rank 0's share of the frame) ------------------------------------------
    roof = cpu = None
    if rank == 0:
        _, _, cnt = part.render_counted(fr, S, 0)      # event counts in the reference's accounting: the per-pixel kernel
        frame_ms, trace_ms = [], []
        for _ in range(5):
            flush.zero_()
            frame_ms.append(part.render_timed(fr, S, args.variant))
            trace_ms.append(fr.last_trace_ms)
        ms_frame, ms_trace = float(np.mean(frame_ms)), float(np.mean(trace_ms))
        n_trace = max(fr.last_trace_launches, 1)
        rays_rank = part.owned_rows * w * S
        algo_frame = algorithmic_bytes(cnt, rays_rank)
        # the grid-walk share of the formula is what the trace kernel is responsible for: 8 B per cell looked at, 68 B per
        # candidate tested (reference layout), plus its own ray fetch (36 B) and hit store (16 B) per grid ray
        algo_trace = 8.0 * cnt["cells"] + 68.0 * cnt["gridCandidates"] + 52.0 * cnt["gridRays"]
        peak, which = hbm_peak()
        achieved = (algo_trace / n_trace) / (ms_trace / n_trace * 1e-3) / 1e9
        roof = {"kernel": "wf_pipe_kernel (+ wf_setup_kernel)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": which, "launches_per_step": n_trace, "avg_launch_ms": ms_trace / n_trace,
                "algorithmic_bytes_per_launch": algo_trace / n_trace,
                "note": "working set is L2-resident (DRAM traffic << algorithmic bytes, see traffic); the kernel is issue-bound, "
                        "profiles/ has issue-slot and lane-utilisation counters",
                "whole_step": {"ms": ms_frame, "algorithmic_bytes": algo_frame, "achieved_gbs": algo_frame / (ms_frame * 1e-3) / 1e9,
                               "frac": algo_frame / (ms_frame * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_ray": algo_frame / rays_rank},
                "per_ray_events": {k: v / rays_rank for k, v in cnt.items() if v},
                # comparable with other tracers: ring segments and grid traversals (segments' closest-hit walks +
                # shadow / occluder walks) per second, rank 0's share at its own frame time
                "rates": {"segments_per_s": cnt["segments"] / (ms_frame * 1e-3), "traversals_per_s": cnt["gridRays"] / (ms_frame * 1e-3),
                          "triangle_tests_per_s_reference_accounting": (cnt["primCandidates"] + cnt["gridCandidates"]) / (ms_frame * 1e-3)}}
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(prof):
            try:
                t = json.load(open(prof)).get(cfg["name"])
                if t:
                    roof["traffic"] = t["dram_bytes_per_launch"]
                    roof["traffic_source"] = t.get("source")
            except Exception:
                pass
        if world == 1:
            cpu = cpu_baseline(cfg, sc, cam, lists)
    # ---- the same call the way the plugin makes it: plain (pageable) arrays, as render.cpp:1086-1134 allocates them.  Informational,
    # and measured LAST: on a multi-GPU box the driver's pageable staging path was seen to slow the pinned calls
    # that followed it (285 -> 175-185 Mrays/s); the other order leaves the headline figure alone.
    e2e_pageable = None
    if world == 1 and os.environ.get("OCLR_BENCH_NO_PAGEABLE") != "1":
        import copy
        sc_pg = copy.copy(sc)
        for name in ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_pos",
                     "light_dir", "light_colour", "light_radius", "light_half", "box_min", "grid_start", "grid_list"):
            setattr(sc_pg, name, np.array(getattr(sc, name), copy=True))
        lists_pg = api.CameraLists(lists.start.copy(), lists.end.copy(), lists.list.copy())
        out_pg = tuple(np.zeros((h, w), np.uint16) for _ in range(3))
        for _ in range(2):
            api.raytrace_all(1 + local, cam, lists_pg, S, sc_pg, out=out_pg)
        t0 = time.perf_counter()
        for _ in range(5):
            api.raytrace_all(1 + local, cam, lists_pg, S, sc_pg, out=out_pg)
        e2e_pageable = rays_frame * 5 / (time.perf_counter() - t0) / 1e6
        del sc_pg, lists_pg


    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["name"], "triangles": sc.triangle_count, "width": w, "height": h, "samples": S,
                       "ray_unit": "pixel-sample (raytrace.c:545)", "l2": "flushed before every timed step (256 MiB memset)",
                       "partition": f"row bands of {args.band_rows} dealt round-robin over {world} GPU(s), scene replicated, "
                                    f"frame assembled by {gather_kind}" if world > 1 else "single GPU, whole frame",
                       "kernel_variant": args.variant},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": e2e.h2d_bytes, "d2h_bytes_per_step": e2e.d2h_bytes,
                    "steps": e2e_steps, "call": "RaytraceAll (C-ABI, host buffers)" if world == 1 else "oclr scene/frame API, rank-local rows"},
            "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roof,
            "gather": {"included_in_value": world > 1, "ms_per_step": (ms_total - ms_render) / args.steps,
                       "value_without_gather": rays_frame / (ms_render / args.steps) / 1e3},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if e2e_pageable is not None:     # informational: RaytraceAll fed from pageable host arrays (staged uploads, runtime.cu StagePool)
            line["e2e"]["value_pageable_host_arrays"] = e2e_pageable
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(cfg, sc, cam, lists):
    """Bounded sample: a quarter-frame band of the same workload on all host cores (reference build when it travelled)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    cores = os.cpu_count() or 1
    h, w = cam.height, cam.width
    step_rows = 4                                # every 4th row of the whole frame (representative of sky and geometry alike)
    n_rows = len(range(0, h, step_rows))
    kind = "port"
    fn = None
    try:
        import ref
        if ref.LIB.is_file():
            ref.load()
            kind = "reference"
            fn = lambda: ref.render(cam, lists, sc, cfg["samples"], threads=cores, row_step=step_rows)
    except Exception:
        fn = None
    if fn is None:
        import port
        kind = "port"
        fn = lambda: port.render(cam, lists, sc, cfg["samples"], threads=cores, row_step=step_rows)
    fn()
    reps = 0
    t0 = time.perf_counter()
    while True:
        fn()
        reps += 1
        if time.perf_counter() - t0 > 10.0 or reps >= 20:
            break
    dt = time.perf_counter() - t0
    rays = n_rows * w * cfg["samples"] * reps
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind,
            "sample": f"every {step_rows}th row of all {h}, {reps} passes, {dt:.1f} s"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
