#!/usr/bin/env python3
"""Lean profiling target: upload one config and launch the trace kernel a few times (run under ncu by gpurun)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes  # noqa: E402

cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variant = int(sys.argv[2]) if len(sys.argv) > 2 else -1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = scenes.CONFIGS[cfg_id]
sc = cfg["make"]()
m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc)
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0)
fr = api.DeviceFrame(ds, cam, lists)
for _ in range(reps):
    ms, launches, _ = fr.render(cfg["samples"], variant=variant)
    print(f"cfg{cfg_id} variant {variant}: {ms:.3f} ms, {cfg['width'] * cfg['height'] * cfg['samples'] / ms / 1e3:.1f} Mrays/s")
