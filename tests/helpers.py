"""Shared helpers for the parity tests: seeded scene cases, image comparison, the host-emulation door."""
import ctypes as C

import numpy as np

from opencl_render_b200 import _lib, api, scenes

RGB_TOL = 1e-3      # north_star: max-abs 1e-3 per channel (of full scale 65535)
PSNR_MIN = 60.0     # north_star: PSNR >= 60 dB


def make_case(name):
    """Small seeded instances of the five config families (+ edge cases).  Returns (scene, camera, lists, samples)."""
    cases = {
        "soup": (lambda: scenes.soup(400, seed=11), 192, 160, 1, 256),
        "soup_s4": (lambda: scenes.soup(200, seed=12, light_radius=0.3), 96, 80, 4, 64),
        "soup_mirror_glass": (lambda: scenes.soup(300, seed=5, light_radius=0.4, reflective=True, transparent=True), 160, 120, 3, 256),
        "spheres": (lambda: scenes.sphere_grid(3, 12, 24), 320, 180, 1, 256),
        "spheres_mirror": (lambda: scenes.sphere_grid(2, 10, 20, reflection=128, light_radius=0.5), 200, 120, 2, 256),
        "terrain": (lambda: scenes.terrain(48), 240, 135, 1, 256),
        "terrain_textured": (lambda: scenes.terrain(40, textured=True, tile_quads=8, mirror_spheres=3), 240, 135, 1, 256),
        "coarse_grid": (lambda: scenes.soup(150, seed=3), 100, 75, 2, 16),
    }
    make, w, h, samples, axes = cases[name]
    sc = make()
    m = sc.meta["camera"]
    cam = api.set_camera(m["eye"], m["look_at"], m["up"], m["fov"], w, h)
    lists = api.camera_triangle_list(cam, sc)
    api.scene_triangle_list(sc, axes)
    return sc, cam, lists, samples


CASE_NAMES = ["soup", "soup_s4", "soup_mirror_glass", "spheres", "spheres_mirror", "terrain", "terrain_textured", "coarse_grid"]


def compare_rgb(a, b, mask=None):
    """Returns dict(diff_pixels, max_abs (fraction of full scale), psnr) over the pixels where mask is True (default all)."""
    diff = np.zeros(a[0].shape, bool)
    mx = 0
    se = 0.0
    n = 0
    for c in range(3):
        x = a[c].astype(np.int64)
        y = b[c].astype(np.int64)
        if mask is not None:
            x, y = x[mask], y[mask]
            diff[mask] |= (x != y)
        else:
            diff |= (x != y)
        if x.size:
            mx = max(mx, int(np.abs(x - y).max()))
            se += float((((x - y) / 65535.0) ** 2).sum())
            n += x.size
    mse = se / max(n, 1)
    return dict(diff_pixels=int(diff.sum()), max_abs=mx / 65535.0, psnr=float("inf") if mse == 0 else 10 * np.log10(1.0 / mse))


def hostemu_render(lib, cam, lists, sc, samples=1, threads=4, rows=None):
    h, w = cam.height, cam.width
    out = [np.zeros((h, w), np.uint16) for _ in range(3)]
    ids = np.full((h, w), 0xFFFFFFFF, np.uint32)
    flags = np.zeros((h, w), np.uint8)
    cnt = _lib.Counters()
    d = sc.desc()
    c = cam.c()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lst = lists.list if lists.list.size else np.zeros(1, np.uint32)
    r0, r1 = rows if rows else (0, h)
    lib.hostemu_render.restype = C.c_int
    ok = lib.hostemu_render(C.byref(d), C.byref(c), p(lists.start), p(lists.end), p(lst), C.c_uint32(samples), C.c_uint32(r0),
                            C.c_uint32(r1), p(out[0]), p(out[1]), p(out[2]), p(ids), p(flags), C.byref(cnt), C.c_int(threads))
    assert ok
    return tuple(out), ids, flags, cnt.as_dict()
