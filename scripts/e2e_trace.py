#!/usr/bin/env python3
"""Phase timing of the drop-in call (OCLR_TRACE=1) with pinned and pageable host buffers:
python scripts/e2e_trace.py [CFG] [WORLD] [CALLS] -- WORLD > 1 = RaytraceAll(all devices) on the first WORLD GPUs from this one process;
checks the planes against a 1-GPU render and prints per-call wall time (spread = run-to-run stability of the call)."""
import os, sys, time
os.environ.setdefault("OCLR_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from opencl_render_b200 import api, scenes, dist as odist
cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
world = int(sys.argv[2]) if len(sys.argv) > 2 else 1
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 6
cfg = scenes.CONFIGS[cfg_id]
sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
want = api.raytrace_all(1, cam, lists, cfg["samples"], sc)
rays = cam.width * cam.height * cfg["samples"]
for label in ("pinned", "pageable"):
    e = odist.EndToEnd(sc, cam, lists, world, 0, n_devices=torch.cuda.device_count(), pinned=(label == "pinned"))
    ts = []
    for i in range(calls):
        t = time.perf_counter(); e.step(cfg["samples"]); ts.append((time.perf_counter() - t) * 1e3)
    same = all(np.array_equal(e.out[c], want[c]) for c in range(3))
    warm = ts[2:] or ts
    print(f"{label} x{world}: calls {['%.2f' % x for x in ts]} ms; warm median {np.median(warm):.2f} ms = {rays / np.median(warm) / 1e3:.1f} Mrays/s, "
          f"spread {100 * (max(warm) - min(warm)) / np.median(warm):.1f} %; planes equal to the 1-GPU call: {same}", flush=True)
