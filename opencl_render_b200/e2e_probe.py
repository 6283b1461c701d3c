"""One process, N GPUs: times the drop-in call RaytraceAll(all devices) (C-ABI, host buffers in and out) on the first WORLD GPUs of the
box and prints ONE JSON line.  bench.py's rank 0 runs it as `python -m opencl_render_b200.e2e_probe CFG WORLD CALLS` while the other
ranks wait at a CPU-side barrier: the caller is a plain host process like the Cinema4D plugin (render.cpp:1314), with no torch, no NCCL
and no resident state -- every call uploads, repacks, traces and reads back.

  {"ms_per_call": .., "calls": .., "spread": .., "sha256": <planes>, "pageable": {...same...}, "h2d_bytes": .., "d2h_bytes": ..}
"""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import sys
import time

import numpy as np

from . import _lib, api, scenes


def _pinned(lib_rt, a: np.ndarray):
    """A copy of `a` in page-locked host memory (cudaHostAlloc through the CUDA runtime the library links)."""
    a = np.ascontiguousarray(a)
    if a.nbytes == 0:
        return a, None
    p = C.c_void_p()
    if lib_rt.cudaHostAlloc(C.byref(p), C.c_size_t(a.nbytes), C.c_uint(1)) != 0:      # cudaHostAllocPortable
        raise RuntimeError("cudaHostAlloc failed")
    v = np.frombuffer((C.c_char * a.nbytes).from_address(p.value), dtype=a.dtype).reshape(a.shape)
    v[...] = a
    return v, p


SCENE_ARRAYS = ("vertex", "tri_idx", "tri_mat", "tri_uv", "tri_normal", "mat_size", "mat_start", "textures", "light_type", "light_pos",
                "light_dir", "light_colour", "light_radius", "light_half", "box_min", "grid_start", "grid_list")


def measure(cfg_id: int, world: int, calls: int, pinned: bool, rt):
    import copy
    cfg = scenes.CONFIGS[cfg_id]
    sc = cfg["make"]()
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
    lists = api.camera_triangle_list(cam, sc)
    api.scene_triangle_list(sc, 256)
    keep = []

    def conv(a):
        if not pinned:
            return np.array(a, copy=True)
        v, p = _pinned(rt, a)
        keep.append(p)
        return v

    s2 = copy.copy(sc)
    for name in SCENE_ARRAYS:
        setattr(s2, name, conv(getattr(sc, name)))
    l2 = api.CameraLists(conv(lists.start), conv(lists.end), conv(lists.list))
    out = tuple(conv(np.zeros((cam.height, cam.width), np.uint16)) for _ in range(3))
    n_dev = int(_lib.load().oclr_device_count())
    if world > 1:
        api.set_option("devices", world)
        ctype = n_dev + 1
    else:
        ctype = 1
    S = cfg["samples"]
    for _ in range(6):      # (the stream-ordered allocator needs a few calls to reach its steady state: the first ones are 2-5x slower)
        api.raytrace_all(ctype, cam, l2, S, s2, out=out)
    ts = []
    for _ in range(calls):
        t = time.perf_counter()
        api.raytrace_all(ctype, cam, l2, S, s2, out=out)
        ts.append((time.perf_counter() - t) * 1e3)
    h = hashlib.sha256()
    for p in out:
        h.update(np.ascontiguousarray(p).tobytes())
    h2d = int(sum(getattr(s2, n).nbytes for n in SCENE_ARRAYS) + l2.start.nbytes + l2.end.nbytes + l2.list.nbytes)
    med = float(np.median(ts))
    return {"ms_per_call": float(np.mean(ts)), "ms_median": med, "ms_min": float(min(ts)), "ms_max": float(max(ts)),
            "spread": (max(ts) - min(ts)) / med, "ms_all": [round(t, 3) for t in ts], "calls": calls, "sha256": h.hexdigest(), "h2d_bytes": h2d,
            "d2h_bytes": int(6 * cam.width * cam.height), "rays": int(cam.width * cam.height * S), "devices_visible": n_dev}


def main():
    cfg_id, world, calls = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    want_pageable = len(sys.argv) < 5 or sys.argv[4] != "0"
    lib = _lib.load()
    if lib.oclr_device_count() < world:
        raise SystemExit(f"e2e_probe: {world} GPUs asked for, {lib.oclr_device_count()} visible")
    rt = None
    for name in ("libcudart.so.12", "libcudart.so", "/usr/local/cuda/lib64/libcudart.so"):
        try:
            rt = C.CDLL(name)      # (only for cudaHostAlloc: page-locked memory is a property of the driver, whichever runtime copy asks for it)
            break
        except OSError:
            continue
    res = {}
    if rt is not None:
        res = measure(cfg_id, world, calls, True, rt)
        res["host_arrays"] = "pinned (cudaHostAlloc)"
    else:
        res = measure(cfg_id, world, calls, False, None)
        res["host_arrays"] = "pageable (no libcudart found for pinned allocations)"
    if want_pageable and rt is not None:
        pg = measure(cfg_id, world, max(5, calls // 2), False, None)
        pg["host_arrays"] = "pageable (plain numpy arrays, like the plugin's new[])"
        res["pageable"] = pg
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
