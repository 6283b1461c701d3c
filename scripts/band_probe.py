#!/usr/bin/env python3
"""Band height of the multi-GPU partition: device time of EVERY rank's share of a frame, one after the other on one GPU (the step of an
N-GPU run is the slowest rank's).   python scripts/band_probe.py CFG WORLD ROWS,ROWS,..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]); world = int(sys.argv[2]); rows_list = [int(v) for v in sys.argv[3].split(",")]
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam)
for rows in rows_list:
    per = []
    for rank in range(world):
        best = min(fr.render_bands(cfg["samples"], rows, rank, world)[0] for _ in range(5))
        per.append(best)
    owned = [sum(b - a for a, b in api.band_partition(cfg["height"], rank, world, band_rows=rows)) for rank in range(world)]
    print(f"cfg{cfg_id} world {world} band rows {rows}: max {max(per):.3f} ms, mean {sum(per) / world:.3f} ms, per rank " +
          " ".join(f"{t:.3f}" for t in per) + f"; rows owned {owned}")
