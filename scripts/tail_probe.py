#!/usr/bin/env python3
"""Where does a trace launch spend its time?  Counting build of wf_pipe_kernel on rank 0's share of a band-partitioned frame: when the
warps of the persistent grid learn that the ray queue is dry and when they leave (32.768-us buckets since the warp's start, all
ray-carrying launches of the frame summed), outer iterations per warp.   python scripts/tail_probe.py CFG WORLD"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]); world = int(sys.argv[2])
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam)
for _ in range(2):
    ms, launches, _ = fr.render_bands(cfg["samples"], 16, 0, world)
print(f"cfg{cfg_id} rank 0 of {world}: {ms:.3f} ms, trace {fr.last_trace_ms:.3f} ms in {fr.last_trace_launches} launches (production build)")
ms, launches, c = fr.render_bands(cfg["samples"], 16, 0, world, count=True)
print(f"counting build: {ms:.3f} ms, trace {fr.last_trace_ms:.3f} ms; rays {c['gridRays']}, warps that ran {c['warpsRun']}, outer iterations per warp mean "
      f"{c['warpOuterItersSum'] / max(c['warpsRun'], 1):.1f} max {c['warpOuterItersMax']}; walk iterations per warp mean {c['walkWarpIters'] / max(c['warpsRun'], 1):.1f}, "
      f"after the queue ran dry {c['walkExhaustedIters'] / max(c['walkWarpIters'], 1):.3f}; walking lanes per walk iteration {c['walkLaneIters'] / max(c['walkWarpIters'], 1):.1f}")
print("  queue seen dry at (x 32.8 us):", [c["exhaustHist%02d" % i] for i in range(16)])
print("  warps leaving at   (x 32.8 us):", [c["exitHist%02d" % i] for i in range(16)])
