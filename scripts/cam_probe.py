#!/usr/bin/env python3
"""Host vs device camera-list builder: time and equality on one config."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg = scenes.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
t = time.time(); lists = api.camera_triangle_list(cam, sc); th = time.time() - t
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0)
for k in range(3):
    t = time.time(); fd = api.DeviceFrame(ds, cam); td = time.time() - t
    if k < 2: fd.close()
got = fd.camera_lists()
print(f"{cfg['name']}: host builder {th*1e3:.1f} ms, device builder (frame create incl. allocs) {td*1e3:.1f} ms; refs {lists.list.size}; "
      f"equal start {np.array_equal(got.start, lists.start)} end {np.array_equal(got.end, lists.end)} list {np.array_equal(got.list, lists.list)}")
