#!/usr/bin/env python3
"""A/B of the kernel variants on one config: bit-equality of planes/ids and device time (trace kernels separately)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2"])]
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam, lists)
rays = cfg["width"] * cfg["height"] * cfg["samples"]
base = None
for v in variants:
    ms, launches, c = fr.render(cfg["samples"], variant=v, count=True)
    img = fr.read(); ids = fr.primary_ids()
    if base is None:
        base = (img, ids)
    same = all(np.array_equal(img[k], base[0][k]) for k in range(3)) and np.array_equal(ids, base[1])
    t = []; tt = []
    for _ in range(5):
        t.append(fr.render(cfg["samples"], variant=v)[0]); tt.append(fr.last_trace_ms)
    print(f"cfg{cfg_id} variant {v}: equal-to-first {same}; frame {min(t):.3f} ms ({rays / min(t) / 1e3:.1f} Mrays/s), trace {min(tt):.3f} ms in {fr.last_trace_launches} launches, {launches} launches total")
    if c["walkWarpIters"]:
        print("   walk util %.3f (%.2f warp iters/ray)  test util %.3f (%.2f warp iters/ray)  cells/ray %.1f  cand/ray %.1f  coarse steps/ray %.1f" % (
            c["walkLaneIters"] / c["walkWarpIters"] / 32, c["walkWarpIters"] / c["gridRays"], c["testLaneIters"] / c["testWarpIters"] / 32,
            c["testWarpIters"] / c["gridRays"], c["cells"] / c["gridRays"], c["gridCandidates"] / c["gridRays"], c["coarseSteps"] / c["gridRays"]))
        if c["switchWarpIters"]:
            print("   switch util %.3f (%.2f warp iters/ray)" % (c["switchLaneIters"] / c["switchWarpIters"] / 32, c["switchWarpIters"] / c["gridRays"]))
        if c.get("walkIdleLanes") is not None and c["walkWarpIters"]:
            wl = c["walkWarpIters"] * 32
            print("   walk-iteration lanes: walking %.3f idle %.3f parked %.3f finished %.3f" % (c["walkLaneIters"] / wl, c["walkIdleLanes"] / wl, c["walkParkedLanes"] / wl, c["walkFinishedLanes"] / wl))
            print("   walk iterations with <= 8 walkers: %.3f; after the queue ran dry: %.3f" % (c["walkLowIters"] / c["walkWarpIters"], c["walkExhaustedIters"] / c["walkWarpIters"]))
