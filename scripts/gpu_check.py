#!/usr/bin/env python3
"""Developer check on a GPU box: CUDA path vs the compiled reference (oracle/_ref) on configs 1-2 (+ optional 3),
device timings and event counters.  Not part of the product; tests/ holds the real parity tests."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from opencl_render_b200 import api, scenes  # noqa: E402
import ref  # noqa: E402


def run(cfg_id, variants=(0, 2), ref_rows=None, reps=5):
    cfg = scenes.CONFIGS[cfg_id]
    t = time.time(); sc = cfg["make"](); t_scene = time.time() - t
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
    t = time.time(); lists = api.camera_triangle_list(cam, sc); t_cam = time.time() - t
    t = time.time(); api.scene_triangle_list(sc, 256); t_grid = time.time() - t
    print(f"[cfg{cfg_id}] {cfg['name']}: tris={sc.triangle_count} scene {t_scene:.2f}s camlist {t_cam:.2f}s ({lists.list.size} refs) "
          f"grid {t_grid:.2f}s ({sc.grid_list.size} refs)", flush=True)
    t = time.time(); ds = api.DeviceScene(sc, 0); fr = api.DeviceFrame(ds, cam, lists); t_up = time.time() - t
    print(f"  upload {t_up:.2f}s device bytes {ds.device_bytes/1e6:.1f} MB", flush=True)
    S = cfg["samples"]
    out = {}
    for v in variants:
        ms, launches, cnt = fr.render(S, variant=v, count=True)
        times = []
        for _ in range(reps):
            ms, launches, _ = fr.render(S, variant=v)
            times.append(ms)
        img = fr.read()
        ids = fr.primary_ids()
        flags = fr.undefined_flags()
        rays = cfg["width"] * cfg["height"] * S
        best = min(times)
        print(f"  variant {v}: device ms {['%.3f' % x for x in times]} -> {rays / best / 1e3:.1f} Mrays/s; launches {launches}", flush=True)
        print("   per-ray counters:", {k: round(c / rays, 3) for k, c in cnt.items()}, flush=True)
        out[v] = (img, ids, flags, best)
    rows = ref_rows or (0, cfg["height"])
    t = time.time(); R = ref.render(cam, lists, sc, S, rows=rows); t_ref = time.time() - t
    nref = (rows[1] - rows[0]) * cfg["width"] * S
    print(f"  reference C kernel on {os.cpu_count()} threads rows {rows}: {t_ref:.2f}s -> {nref / t_ref / 1e6:.3f} Mrays/s", flush=True)
    for v, (img, ids, flags, best) in out.items():
        sl = slice(rows[0], rows[1])
        diff = [(img[c][sl] != R[c][sl]) for c in range(3)]
        anyd = diff[0] | diff[1] | diff[2]
        mx = max(int(np.abs(img[c][sl].astype(np.int64) - R[c][sl].astype(np.int64)).max()) for c in range(3))
        unfl = anyd & (flags[sl] == 0)
        mse = np.mean([(img[c][sl].astype(np.float64) / 65535 - R[c][sl].astype(np.float64) / 65535) ** 2 for c in range(3)])
        psnr = float("inf") if mse == 0 else 10 * np.log10(1.0 / mse)
        print(f"  variant {v} vs reference: differing pixels {int(anyd.sum())} (unflagged {int(unfl.sum())}, flagged px {int(flags[sl].sum())}) "
              f"max|d|={mx} ({mx / 65535:.2e}) PSNR={psnr:.1f} dB", flush=True)
    if len(out) == 2:
        a, b = out[variants[0]], out[variants[1]]
        print(f"  variant{variants[0]} == variant{variants[1]}:", all(np.array_equal(a[0][c], b[0][c]) for c in range(3)), "ids equal:", np.array_equal(a[1], b[1]))
    fr.close(); ds.close()


def run_sweep(cfg_id=5, check_frames=(0, 21, 42)):
    """Config 5: scene uploaded once, 64 cameras on a circle, per-frame camera lists; mirror chains up to depth 12."""
    cfg = scenes.CONFIGS[cfg_id]
    sc = cfg["make"]()
    api.scene_triangle_list(sc, 256)
    t = time.time(); ds = api.DeviceScene(sc, 0); t_up = time.time() - t
    print(f"[cfg{cfg_id}] {cfg['name']}: tris={sc.triangle_count} upload once {t_up:.2f}s ({ds.device_bytes/1e6:.1f} MB)", flush=True)
    cams = scenes.sweep_cameras(sc, cfg["frames"])
    total_ms, total_rays, t_lists, worst = 0.0, 0, 0.0, 0
    for k, m in enumerate(cams):
        cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], cfg["width"], cfg["height"])
        t = time.time(); fr = api.DeviceFrame(ds, cam); t_lists += time.time() - t       # camera lists built on the device
        lists = None
        ms, launches, _ = fr.render(cfg["samples"])
        total_ms += ms; total_rays += cfg["width"] * cfg["height"] * cfg["samples"]
        if k in check_frames:
            img = fr.read(); flags = fr.undefined_flags()
            lists = api.camera_triangle_list(cam, sc)                # host builder, for the reference run only
            got = fr.camera_lists()
            assert np.array_equal(got.start, lists.start) and np.array_equal(got.end, lists.end) and np.array_equal(got.list, lists.list)
            rows = (500, 532)
            R = ref.render(cam, lists, sc, cfg["samples"], rows=rows)
            sl = slice(*rows)
            d = sum(int(((img[c][sl] != R[c][sl]) & (flags[sl] == 0)).sum()) for c in range(3))
            worst = max(worst, d)
            print(f"  frame {k}: {ms:.2f} ms, {launches} launches, lit {int((img[0] > 0).sum())}, diff vs reference rows {rows}: {d}", flush=True)
        fr.close()
    print(f"  {len(cams)} frames: device {total_ms:.1f} ms total -> {total_rays / total_ms / 1e3:.1f} Mrays/s; frame creation incl. device camera lists {t_lists:.2f}s; worst diff {worst}")
    ds.close()


if __name__ == "__main__":
    print("types:", api.computation_types())
    for cid in [x for x in (sys.argv[1:] or ["1", "2"])]:
        if cid == "5s":
            run_sweep(5)
        else:
            cid = int(cid)
            run(cid, ref_rows=None if cid <= 2 else (1000, 1064), reps=5 if cid <= 3 else 2)
