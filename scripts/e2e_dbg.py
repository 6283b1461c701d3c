import os, sys, time, copy
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from opencl_render_b200 import api, dist as odist
cfg, sc, cam, lists = bench.build_workload(2)
h, w, S = cam.height, cam.width, cfg["samples"]
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
if mode in ("both", "pageable"):
    sc_pg = copy.copy(sc)
    for name in ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_pos",
                 "light_dir", "light_colour", "light_radius", "light_half", "box_min", "grid_start", "grid_list"):
        setattr(sc_pg, name, np.array(getattr(sc, name), copy=True))
    lists_pg = api.CameraLists(lists.start.copy(), lists.end.copy(), lists.list.copy())
    out_pg = tuple(np.zeros((h, w), np.uint16) for _ in range(3))
    for i in range(3):
        t = time.perf_counter(); api.raytrace_all(1, cam, lists_pg, S, sc_pg, out=out_pg); print("pageable %.2f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    del sc_pg, lists_pg
part = odist.BandPartition(h, w, 0, 1, 16)
e2e = odist.EndToEnd(sc, cam, lists, part, 0)
for i in range(int(os.environ.get("CALLS", "5"))):
    t = time.perf_counter(); e2e.step(S); a = time.perf_counter(); torch.cuda.synchronize(); b = time.perf_counter()
    print("pinned call %.2f ms, sync after %.2f ms" % ((a - t) * 1e3, (b - a) * 1e3), flush=True)
