"""Shared helpers for the parity tests: seeded scene cases, image comparison, the host-emulation door."""
import ctypes as C

import numpy as np

from opencl_render_b200 import _lib, api, scenes

RGB_TOL = 1e-3      # north_star: max-abs 1e-3 per channel (of full scale 65535)
PSNR_MIN = 60.0     # north_star: PSNR >= 60 dB


def _with_lights(sc, entries):
    """Replaces the scene's lights (several lights per scene: the light loop raytrace_opencl.c:563-637 runs one shadow query per
    light in order; the omni type contributes nothing, :585-588)."""
    lt, pos, dr, col, rad, half = scenes._lights(entries)
    sc.light_type, sc.light_pos, sc.light_dir, sc.light_colour, sc.light_radius, sc.light_half = lt, pos, dr, col, rad, half
    return sc.normalise()


_SPOT = dict(type=api.LIGHT_SPOT, pos=(3.0, 8.0, -5.0), colour=(0.9, 0.8, 0.7), radius=0.25)
_SUN = dict(type=api.LIGHT_DISTANT, dir=(-0.4, -0.8, 0.3), colour=(0.5, 0.55, 0.7), radius=0.52)
_OMNI = dict(type=api.LIGHT_OMNI, pos=(0.0, 6.0, 0.0), colour=(1, 1, 1), radius=0.1)
_AREA = dict(type=api.LIGHT_AREA, pos=(-4.0, 5.0, -3.0), colour=(0.6, 0.6, 0.6), radius=0.6, half=9.0)


# every light type the reference's switch (raytrace_opencl.c:566-607) knows but the config scenes never use -- SPOTRECT (2), PARALLEL (4),
# PARSPOT (5), PARSPOTRECT (6), TUBE (7), PHOTOMETRIC (9) -- plus a type outside the switch (contributes nothing, like OMNI)
_ALL_TYPES = [
    dict(type=api.LIGHT_SPOTRECT, pos=(2.5, 7.0, -4.0), colour=(0.5, 0.45, 0.4), radius=0.2, half=14.0),
    dict(type=api.LIGHT_PARALLEL, dir=(0.3, -0.9, 0.2), colour=(0.25, 0.3, 0.35), radius=0.4),
    dict(type=api.LIGHT_PARSPOT, dir=(-0.5, -0.7, -0.1), colour=(0.3, 0.2, 0.2), radius=1.5, half=40.0),
    dict(type=api.LIGHT_PARSPOTRECT, dir=(0.1, -1.0, -0.4), colour=(0.15, 0.25, 0.15), radius=0.0),
    dict(type=api.LIGHT_TUBE, pos=(-3.5, 4.0, 2.0), colour=(0.4, 0.4, 0.2), radius=0.5),
    dict(type=11, pos=(0.0, 5.0, 0.0), colour=(1, 1, 1), radius=0.3),
    dict(type=api.LIGHT_PHOTOMETRIC, pos=(0.5, 9.0, 1.0), colour=(0.3, 0.3, 0.45), radius=0.05, half=6.0),
]


def make_case(name):
    """Small seeded instances of the five config families (+ edge cases).  Returns (scene, camera, lists, samples)."""
    cases = {
        # several lights: shadow queries of one hit in sequence; the last light that casts one / an omni light in last place
        "soup_lights_sun_last": (lambda: _with_lights(scenes.soup(260, seed=21, reflective=True, transparent=True), [_OMNI, _SPOT, _AREA, _SUN]),
                                 144, 112, 2, 256),
        # axis-parallel shadow rays (a distant light straight overhead with zero angular size: two direction components are exactly
        # 0, so two of the three crossing values are +-inf or NaN in every cell -- the tie rule of raytrace_opencl.c:387-398 and the
        # "no brick-level walk" path are what is tested) plus a spot light exactly above the look-at point
        "soup_axis_light": (lambda: _with_lights(scenes.soup(320, seed=31, reflective=True),
                                                 [dict(type=api.LIGHT_DISTANT, dir=(0.0, -1.0, 0.0), colour=(0.8, 0.8, 0.8), radius=0.0),
                                                  dict(type=api.LIGHT_SPOT, pos=(0.0, 9.0, 0.0), colour=(0.5, 0.4, 0.3), radius=0.0)]),
                            160, 120, 1, 256),
        "soup_lights_omni_last": (lambda: _with_lights(scenes.soup(260, seed=22, reflective=True, transparent=True), [_SPOT, _SUN, _OMNI]),
                                  144, 112, 2, 256),
        "soup_all_light_types": (lambda: _with_lights(scenes.soup(220, seed=41, reflective=True, transparent=True), _ALL_TYPES), 128, 96, 2, 256),
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


CASE_NAMES = ["soup", "soup_s4", "soup_mirror_glass", "spheres", "spheres_mirror", "terrain", "terrain_textured", "coarse_grid",
              "soup_lights_sun_last", "soup_lights_omni_last", "soup_axis_light", "soup_all_light_types"]


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
