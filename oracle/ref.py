"""oracle/ref.py -- TEST INFRASTRUCTURE ONLY.  ctypes doors onto oracle/_ref/libref_raytrace.so, i.e. the UNMODIFIED
reference (source/opencl/raytrace.c + raytrace_opencl.c + util/trianglelist.cpp) compiled by oracle/build_ref.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product package (opencl_render_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_ref" / "libref_raytrace.so"
LIB_COUNTED = HERE / "_ref" / "libref_raytrace_counted.so"
COUNTER_NAMES = ["tests", "gridRays", "cells", "emptyCells", "segments", "primCandidates", "shadedHits", "gridCandidates", "occluderLookups"]
_lib = None
_counted = None


def available() -> bool:
    return LIB.is_file() or (Path("/root/reference/source/opencl/raytrace.c").is_file())


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    sys.path.insert(0, str(HERE))
    import build_ref
    if not LIB.is_file() or build_ref.reference_available():     # (re)build when the recipe changed; the GPU box uses what travelled
        if build_ref.build(verbose=False) is None:
            raise RuntimeError("oracle/_ref/libref_raytrace.so missing and /root/reference not present to build it")
    lib = C.CDLL(str(LIB))
    lib.ref_camera_list_new.restype = C.c_void_p
    lib.ref_camera_list_new.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                        C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.ref_camera_list_size.restype = C.c_ssize_t
    lib.ref_camera_list_size.argtypes = [C.c_void_p]
    lib.ref_camera_list_copy.restype = None
    lib.ref_camera_list_copy.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_camera_list_free.argtypes = [C.c_void_p]
    lib.ref_camera_list_free.restype = None
    lib.ref_scene_axes_division.restype = C.c_int
    lib.ref_scene_list_new.restype = C.c_void_p
    lib.ref_scene_list_new.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.ref_scene_list_size.restype = C.c_uint32
    lib.ref_scene_list_size.argtypes = [C.c_void_p]
    lib.ref_scene_list_copy.restype = None
    lib.ref_scene_list_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_scene_list_free.argtypes = [C.c_void_p]
    lib.ref_scene_list_free.restype = None
    lib.ref_raytrace_threads.restype = None
    lib.ref_raytrace_threads.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ref_raytrace_all.restype = C.c_uint32
    lib.ref_raytrace_all.argtypes = [C.c_uint32] + lib.ref_raytrace_threads.argtypes[3:]
    lib.ref_raytrace_threads_step.restype = None
    lib.ref_raytrace_threads_step.argtypes = lib.ref_raytrace_threads.argtypes[:3] + [C.c_uint32] + lib.ref_raytrace_threads.argtypes[3:]
    if hasattr(lib, "ref_set_camera"):
        lib.ref_set_camera.restype = None
        lib.ref_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def set_camera(eye, look_at, up, fov: float, width: int, height: int):
    """The reference's SetCamera (render.cpp:461-491).  Returns (eye_to_top_left[3], left_to_right[3], top_to_bottom[3], pixel_size_inv)."""
    lib = load()
    e, o, u = _f4(eye), _f4(look_at), _f4(up)
    tl, lr, tb, psi = np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(4, np.float32), np.zeros(1, np.float32)
    lib.ref_set_camera(_p(e), _p(o), _p(u), C.c_float(fov), width, height, _p(tl), _p(lr), _p(tb), _p(psi))
    return tl[:3].copy(), lr[:3].copy(), tb[:3].copy(), float(psi[0])


def counted_available() -> bool:
    return LIB_COUNTED.is_file()


def render_counted(camera, lists, scene, samples: int = 1, threads: int | None = None, rows=None, row_step: int = 1):
    """The INSTRUMENTED copy of the reference kernel (build_ref.py): same planes as render() plus the event
    counts of the rendered rows as a dict keyed by COUNTER_NAMES."""
    global _counted
    if _counted is None:
        if not LIB_COUNTED.is_file():
            load()          # builds both libraries when /root/reference is present
        lib = C.CDLL(str(LIB_COUNTED))
        lib.ref_raytrace_threads_step.restype = None
        lib.ref_raytrace_threads_step.argtypes = load().ref_raytrace_threads_step.argtypes
        lib.ref_counters_reset.restype = None
        lib.ref_counters_read.restype = None
        lib.ref_counters_read.argtypes = [C.c_void_p]
        _counted = lib
    lib = _counted
    h, w = camera.height, camera.width
    r0, r1 = rows if rows is not None else (0, h)
    out = [np.zeros((h, w), np.uint16) for _ in range(3)]
    args, keep = _scene_args(camera, lists, scene, samples)
    lib.ref_counters_reset()
    lib.ref_raytrace_threads_step(threads or (os.cpu_count() or 1), r0, r1, row_step, *args, _p(out[0]), _p(out[1]), _p(out[2]))
    raw = np.zeros(16, np.uint64)
    lib.ref_counters_read(_p(raw))
    return tuple(out), {k: int(raw[i]) for i, k in enumerate(COUNTER_NAMES)}


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _f4(v):
    a = np.zeros(4, np.float32)
    a[:3] = np.asarray(v, np.float32)[:3]
    return a


def camera_lists(camera, scene):
    """The reference's CameraTriangleList::New.  Returns (start, end, list)."""
    lib = load()
    eye, tl, lr, tb = _f4(camera.eye), _f4(camera.eye_to_top_left), _f4(camera.left_to_right), _f4(camera.top_to_bottom)
    h = lib.ref_camera_list_new(camera.width, camera.height, _p(eye), _p(tl), _p(lr), _p(tb), C.c_float(camera.pixel_size_inv),
                                scene.vertex_count, scene.triangle_count, _p(scene.vertex), _p(scene.tri_idx))
    if not h:
        raise RuntimeError("reference CameraTriangleList::New failed")
    try:
        p = camera.width * camera.height
        n = int(lib.ref_camera_list_size(h))
        start, end, lst = np.empty(p, np.uint32), np.empty(p, np.uint32), np.empty(max(n, 1), np.uint32)
        lib.ref_camera_list_copy(h, p, _p(start), _p(end), _p(lst))
    finally:
        lib.ref_camera_list_free(h)
    return start, end, lst[:n]


def scene_grid(scene):
    """The reference's SceneTriangleList::New (AXES_DIVISION fixed at 256).  Returns (box_min, start, list)."""
    lib = load()
    n = lib.ref_scene_axes_division()
    h = lib.ref_scene_list_new(scene.vertex_count, scene.triangle_count, _p(scene.vertex), _p(scene.tri_idx))
    if not h:
        raise RuntimeError("reference SceneTriangleList::New failed")
    try:
        total = int(lib.ref_scene_list_size(h))
        box, start, lst = np.empty((n + 1, 4), np.float32), np.empty(n ** 3 + 1, np.uint32), np.empty(max(total, 1), np.uint32)
        lib.ref_scene_list_copy(h, _p(box), _p(start), _p(lst))
    finally:
        lib.ref_scene_list_free(h)
    return box, start, lst[:total]


def _scene_args(camera, lists, scene, samples):
    start = np.ascontiguousarray(lists.start, np.uint32)
    end = np.ascontiguousarray(lists.end, np.uint32)
    lst = np.ascontiguousarray(lists.list, np.uint32)
    if lst.size == 0:
        lst = np.zeros(1, np.uint32)
    keep = [start, end, lst, _f4(camera.eye), _f4(camera.eye_to_top_left), _f4(camera.left_to_right), _f4(camera.top_to_bottom)]
    args = [camera.width, camera.height, _p(keep[3]), _p(keep[4]), _p(keep[5]), _p(keep[6]), C.c_float(camera.pixel_size_inv),
            _p(start), _p(end), _p(lst), samples, _p(scene.vertex), scene.triangle_count, _p(scene.tri_idx), _p(scene.tri_mat),
            _p(scene.tri_uv), _p(scene.tri_normal), scene.axes_div, _p(scene.box_min), _p(scene.grid_start), _p(scene.grid_list),
            _p(scene.mat_size), _p(scene.mat_start), _p(scene.textures), scene.light_count, _p(scene.light_type), _p(scene.light_pos),
            _p(scene.light_dir), _p(scene.light_colour), _p(scene.light_radius), _p(scene.light_half)]
    return args, keep


def render(camera, lists, scene, samples: int = 1, threads: int | None = None, rows=None, row_step: int = 1):
    """The reference kernel (compiled as C) over rows [rows[0], rows[1]) -- every row_step-th of them -- on `threads` host threads
    (default: all cores).
    Bit-identical to RaytraceAll(0, ...) for any thread count (the C-path seed depends only on pixel and sample)."""
    lib = load()
    h, w = camera.height, camera.width
    r0, r1 = rows if rows is not None else (0, h)
    out = [np.zeros((h, w), np.uint16) for _ in range(3)]
    args, keep = _scene_args(camera, lists, scene, samples)
    lib.ref_raytrace_threads_step(threads or (os.cpu_count() or 1), r0, r1, row_step, *args, _p(out[0]), _p(out[1]), _p(out[2]))
    return tuple(out)


def raytrace_all(camera, lists, scene, samples: int = 1):
    """The reference's own RaytraceAll(computationType = 0, ...) -- "Local CPU single thread" (raytrace.c:604-655)."""
    lib = load()
    h, w = camera.height, camera.width
    out = [np.zeros((h, w), np.uint16) for _ in range(3)]
    args, keep = _scene_args(camera, lists, scene, samples)
    ok = lib.ref_raytrace_all(0, *args, _p(out[0]), _p(out[1]), _p(out[2]))
    if not ok:
        raise RuntimeError("reference RaytraceAll returned CL_FALSE")
    return tuple(out)
