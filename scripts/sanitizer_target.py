#!/usr/bin/env python3
"""Small end-to-end target for compute-sanitizer (memcheck / racecheck / initcheck): RaytraceAll on a 96x80 mirror+glass soup with 2
samples (upload overlap, run-ahead, two ray slots), a device-built grid + camera lists, a sliced and a banded render, the float
accumulator -- every kernel of the library once.  Prints whether the planes equal the oracle's."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
from opencl_render_b200 import api, scenes
import port
sc = scenes.soup(120, seed=5, light_radius=0.4, reflective=True, transparent=True)
m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], 96, 80)
lists = api.camera_triangle_list(cam, sc)
api.scene_triangle_list(sc, 64)
want = port.render(cam, lists, sc, 2)
got = api.raytrace_all(1, cam, lists, 2, sc)
ok = all(np.array_equal(a, b) for a, b in zip(got, want))
bare = scenes.soup(120, seed=5, light_radius=0.4, reflective=True, transparent=True)
ds = api.DeviceScene(bare, 0, axes_div=64)            # grid built on the device
fr = api.DeviceFrame(ds, cam)                          # camera lists built on the device
api.set_option("slices", 2)
fr.render(2)
ok = ok and all(np.array_equal(a, b) for a, b in zip(fr.read(), want))
api.set_option("slices", 0)
out = tuple(np.zeros((80, 96), np.uint16) for _ in range(3))
for rank in range(2):
    fr.render_bands(2, 16, rank, 2)
    for rows in api.band_partition(80, rank, 2, band_rows=16):
        fr.read(rows=rows, out=out)
ok = ok and all(np.array_equal(a, b) for a, b in zip(out, want))
fr.set_accumulation(api.ACCUMULATE_FLOAT)
fr.render(2)
fr.render(1, variant=api.KERNEL_PIPE)
fr.set_accumulation(api.ACCUMULATE_REFERENCE_16BIT)
fr.render(2, variant=api.KERNEL_SIMPLE)
ok = ok and all(np.array_equal(a, b) for a, b in zip(fr.read(), want))
print("planes equal the oracle's:", ok)
sys.exit(0 if ok else 1)
