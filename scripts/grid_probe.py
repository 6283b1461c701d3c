#!/usr/bin/env python3
"""Host vs device scene-grid builder: time and equality on one config."""
import copy, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg = scenes.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]; sc = cfg["make"]()
t = time.time(); api.scene_triangle_list(sc, 256); th = time.time() - t
dev = copy.copy(sc)
for _ in range(2):
    t = time.time(); api.scene_triangle_list(dev, 256, device=0); td = time.time() - t
print(f"{cfg['name']}: host builder {th*1e3:.0f} ms, device builder {td*1e3:.0f} ms (incl. 67 MB read-back); refs {sc.grid_list.size}; "
      f"equal planes {np.array_equal(dev.box_min.view(np.uint32), sc.box_min.view(np.uint32))} start {np.array_equal(dev.grid_start, sc.grid_start)} "
      f"list {np.array_equal(dev.grid_list, sc.grid_list)}")
