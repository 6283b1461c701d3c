#!/usr/bin/env python3
"""Profiling target: one rank's share of a band-partitioned frame (rank 0 of WORLD), rendered a few times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]); world = int(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam)
best = None
for _ in range(reps):
    ms, launches, _ = fr.render_bands(cfg["samples"], 16, 0, world)
    print(f"cfg{cfg_id} rank 0 of {world}: {ms:.3f} ms, trace {fr.last_trace_ms:.3f} ms, {launches} launches")
    best = (ms, fr.last_trace_ms) if best is None or ms < best[0] else best
print(f"cfg{cfg_id} rank 0 of {world}: BEST {best[0]:.3f} ms, trace {best[1]:.3f} ms")
if os.environ.get("OCLR_SPLIT_MIN", "0") != "0":
    _, _, c = fr.render_bands(cfg["samples"], 16, 0, world, count=True)
    print("   split: attempts %d, splits %d, parts %d, cancelled %d; walk iterations after the queue ran dry %.3f" % (
        c["splitAttempts"], c["splitsDone"], c["splitParts"], c["splitCancelled"], c["walkExhaustedIters"] / max(c["walkWarpIters"], 1)))
