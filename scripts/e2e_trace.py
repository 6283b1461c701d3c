#!/usr/bin/env python3
"""Phase timing of the drop-in call (OCLR_TRACE=1) with pinned and pageable host buffers."""
import os, sys, time
os.environ["OCLR_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from opencl_render_b200 import api, scenes, dist as odist
cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = scenes.CONFIGS[cfg_id]
sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
for label in ("pageable", "pinned"):
    if label == "pinned":
        part = odist.BandPartition(cam.height, cam.width, 0, 1, 16)
        e = odist.EndToEnd(sc, cam, lists, part, 0)
        call = lambda: e.step(cfg["samples"])
    else:
        out = tuple(np.zeros((cam.height, cam.width), np.uint16) for _ in range(3))
        call = lambda: api.raytrace_all(1, cam, lists, cfg["samples"], sc, out=out)
    for i in range(4):
        t = time.perf_counter(); call(); print(label, i, "%.2f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
