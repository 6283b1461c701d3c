#!/usr/bin/env python3
"""Under torchrun: frame assembly by peer stores (PlanePush) vs the NCCL all-gather (PlaneGather) -- equality with the single-GPU
frame and time per assembly.  torchrun --nproc-per-node N scripts/push_probe.py [CFG] [BAND_ROWS]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from opencl_render_b200 import api, scenes, dist as odist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg_id = int(sys.argv[1]) if len(sys.argv) > 1 else 2
band = int(sys.argv[2]) if len(sys.argv) > 2 else 16
cfg = scenes.CONFIGS[cfg_id]; sc = cfg["make"](); m = sc.meta["camera"]
cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
api.scene_triangle_list(sc, 256)
ds = api.DeviceScene(sc, local); fr = api.DeviceFrame(ds, cam)
fr.render(cfg["samples"]); whole = fr.read()
part = odist.BandPartition(cam.height, cam.width, rank, world, band)
stream = torch.cuda.current_stream().cuda_stream
dev = torch.device("cuda", local)
for name, make in (("push", lambda: odist.PlanePush(fr, part, dev)), ("nccl", lambda: odist.PlaneGather(fr, part, dev))):
    try:
        ex = make()
    except Exception as e:
        print(f"rank {rank}: {name} unavailable: {type(e).__name__}: {e}", flush=True)
        continue
    ok = True
    for it in range(3):
        fr.write(tuple(np.zeros_like(p) for p in whole))         # only this rank's rows are valid after the render below
        part.render(fr, cfg["samples"], -1, stream)
        full = ex.run()
        torch.cuda.synchronize()
        got = full.cpu().numpy().view(np.uint16)
        ok = ok and all(np.array_equal(got[c], whole[c]) for c in range(3))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dist.barrier(); torch.cuda.synchronize()
    ev[0].record()
    for _ in range(20):
        ex.run()
    ev[1].record(); torch.cuda.synchronize()
    t = torch.tensor([ev[0].elapsed_time(ev[1]) / 20], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"cfg{cfg_id} world {world} band {band}: {name}: assembled frame equals single-GPU frame: {ok}; {t.item() * 1e3:.1f} us per assembly", flush=True)
    ok_t = torch.tensor([int(ok)], device=dev); dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
    assert ok_t.item() == 1, name
dist.barrier(); dist.destroy_process_group()
