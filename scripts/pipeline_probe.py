#!/usr/bin/env python3
"""Whole pipeline from the raw scene arrays (no acceleration lists): what parseAndRender does after scene extraction
(render.cpp:1311-1352: CameraTriangleList::New, SceneTriangleList::New, RaytraceAll) -- host builders vs everything on the device."""
import copy, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg = scenes.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]; raw = cfg["make"](); m = raw.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
S = cfg["samples"]
def host_path():
    sc = copy.copy(raw)
    t0 = time.time(); lists = api.camera_triangle_list(cam, sc); t1 = time.time(); api.scene_triangle_list(sc, 256); t2 = time.time()
    img = api.raytrace_all(1, cam, lists, S, sc); t3 = time.time()
    return img, (t1 - t0, t2 - t1, t3 - t2)
def device_path():
    sc = copy.copy(raw); sc.box_min = sc.grid_start = sc.grid_list = None
    t0 = time.time(); ds = api.DeviceScene(sc, 0); t1 = time.time(); fr = api.DeviceFrame(ds, cam); t2 = time.time()
    fr.render(S); img = fr.read(); t3 = time.time()
    fr.close(); ds.close()
    return img, (t1 - t0, t2 - t1, t3 - t2)
host_path(); device_path()
a, th = host_path(); b, td = device_path()
print(f"{cfg['name']}: host builders + RaytraceAll: camera lists {th[0]*1e3:.0f} ms + grid {th[1]*1e3:.0f} ms + RaytraceAll {th[2]*1e3:.1f} ms = {sum(th)*1e3:.0f} ms")
print(f"   all on the device: upload + grid + repack {td[0]*1e3:.1f} ms + frame with camera lists {td[1]*1e3:.1f} ms + trace + read back {td[2]*1e3:.1f} ms = {sum(td)*1e3:.1f} ms"
      f"; identical planes: {all(np.array_equal(x, y) for x, y in zip(a, b))}")
