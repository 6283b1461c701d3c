#!/usr/bin/env python3
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from opencl_render_b200 import api, scenes
cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam, lists)
ms, launches, c = fr.render(1, variant=1, count=True)
rays = cfg["width"] * cfg["height"]
print({k: round(v / rays, 3) for k, v in c.items()})
print("walk util %.3f  test util %.3f" % (c["walkLaneIters"] / c["walkWarpIters"] / 32, c["testLaneIters"] / c["testWarpIters"] / 32))
print("walk lane iters per cell %.3f ; test lane iters per candidate %.3f" % (c["walkLaneIters"] / c["cells"], c["testLaneIters"] / c["gridCandidates"]))
print("empty-brick cell share %.3f" % (c["emptyBrickCells"] / c["cells"]))
for t in range(3):
    print("%.3f ms" % fr.render(1, variant=1)[0])
