#!/usr/bin/env python3
"""Trace-stage time for a few rows of config 2 (how long does a launch last when there is hardly any work?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg = scenes.CONFIGS[2]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam)
for rows in ((540, 541), (540, 604), (0, 1080)):
    _, _, c = fr.render(1, rows=rows, count=True)
    for _ in range(3):
        ms, launches, _ = fr.render(1, rows=rows)
    print(f"rows {rows}: frame {ms:.3f} ms, trace stage {fr.last_trace_ms:.3f} ms in {fr.last_trace_launches} rounds; rays {c['gridRays']} walk warp-iters {c['walkWarpIters']} "
          f"lane-iters {c['walkLaneIters']} low iters {c['walkLowIters']} exhausted iters {c['walkExhaustedIters']} bursts {c['coopBursts']} cells queued by bursts {c['coopCells']} of {c['cellsNonEmpty']}")
