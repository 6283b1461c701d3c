#!/usr/bin/env python3
"""Under torchrun: device time of the band render alone, the plane gather alone, and both (config 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["NCCL_DEBUG"] = "WARN"
import torch, torch.distributed as dist
from opencl_render_b200 import api, scenes, dist as odist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = scenes.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
lists = api.camera_triangle_list(cam, sc); api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, local); fr = api.DeviceFrame(ds, cam, lists)
band = int(sys.argv[2]) if len(sys.argv) > 2 else 16
part = odist.BandPartition(cam.height, cam.width, rank, world, band)
gather = odist.PlaneGather(fr, part, torch.device("cuda", local)) if world > 1 else None
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, n=10):
    tot = 0.0
    for _ in range(3): fn()
    for _ in range(n):
        flush.zero_()
        if world > 1: dist.barrier()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    t = torch.tensor([tot / n], device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
r = timed(lambda: part.render(fr, cfg["samples"], -1, stream))
g = timed(lambda: gather.run()) if gather else 0.0
both = timed(lambda: (part.render(fr, cfg["samples"], -1, stream), gather.run() if gather else None))
if rank == 0:
    print(f"world {world} band {band}: render {r:.3f} ms, gather {g:.3f} ms, both {both:.3f} ms (max over ranks)")
if world > 1:
    dist.destroy_process_group()
