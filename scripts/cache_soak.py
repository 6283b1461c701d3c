#!/usr/bin/env python3
"""Soak of the per-call resources of RaytraceAll (block caches, landing arenas, staging arena): a camera sweep -- every call has other
camera-list sizes -- on WORLD GPUs, pageable arrays; prints the device memory in use on every GPU at a few points.  It must level off.
    python scripts/cache_soak.py [WORLD] [CALLS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from opencl_render_b200 import api, scenes
world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 48
sc = scenes.terrain(160, mirror_spheres=3)
api.scene_triangle_list(sc, 256)
cams = scenes.sweep_cameras(sc, calls)
n_dev = torch.cuda.device_count()
if world > 1:
    api.set_option("devices", world)
ctype = n_dev + 1 if world > 1 else 1
def used():
    return [round((torch.cuda.mem_get_info(d)[1] - torch.cuda.mem_get_info(d)[0]) / 2**20) for d in range(max(world, 1))]
print("before", used(), flush=True)
ref = None
for k, m in enumerate(cams):
    w, h = (640 + 16 * (k % 5), 360 + 8 * (k % 3))
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], w, h)
    lists = api.camera_triangle_list(cam, sc)
    out = api.raytrace_all(ctype, cam, lists, 1, sc)
    if k % 12 == 0:
        one = api.raytrace_all(1, cam, lists, 1, sc)
        assert all(np.array_equal(out[c], one[c]) for c in range(3)), k
    if k % 8 == 7:
        print("after call", k + 1, "MiB in use per GPU", used(), flush=True)
print("ok")
