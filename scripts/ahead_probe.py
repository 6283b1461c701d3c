#!/usr/bin/env python3
"""Run-ahead modes (oclr_set_option "ahead"): frame time on a whole frame and on one rank's 1/WORLD band share."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]); worlds = [int(x) for x in sys.argv[2].split(",")]
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam)
base = None
for world in worlds:
    for mode in (0, 1, 2):
        api.set_option("ahead", mode)
        t, tt = [], []
        for _ in range(6):
            ms, launches, _ = fr.render_bands(cfg["samples"], 16, 0, world)
            t.append(ms); tt.append(fr.last_trace_ms)
        img = fr.read()
        if world == worlds[0] and base is None:
            base = img
        same = all(np.array_equal(img[c], base[c]) for c in range(3)) if world == worlds[0] else "-"
        print(f"cfg{cfg_id} 1/{world} share, ahead={mode}: frame {min(t):.3f} ms, trace {min(tt):.3f} ms in {fr.last_trace_launches} launches, {launches} launches; equal {same}", flush=True)
