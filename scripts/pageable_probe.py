#!/usr/bin/env python3
"""RaytraceAll fed from pageable host arrays (what the plugin passes): phase timing (OCLR_TRACE) per call."""
import os, sys, time
os.environ["OCLR_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from opencl_render_b200 import api, scenes
cfg = scenes.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
out = tuple(np.zeros((cam.height, cam.width), np.uint16) for _ in range(3))
for i in range(5):
    t = time.perf_counter(); api.raytrace_all(1, cam, lists, cfg["samples"], sc, out=out)
    print("pageable call %d: %.2f ms" % (i, (time.perf_counter() - t) * 1e3), flush=True)
