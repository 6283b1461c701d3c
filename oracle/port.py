"""oracle/port.py -- TEST INFRASTRUCTURE ONLY.  Builds and drives oracle/raytrace_port.c (the plain-C CPU restatement of
the reference's raytrace path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "raytrace_port.c"
OUT = HERE / "_build"
LIB = OUT / "libraytrace_port.so"
_lib = None


class V3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Job(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("eye", V3), ("top_left", V3), ("step_right", V3), ("step_down", V3),
        ("pixel_size_inv", C.c_float), ("cam_start", C.c_void_p), ("cam_end", C.c_void_p), ("cam_list", C.c_void_p),
        ("samples", C.c_uint32), ("vertex", C.c_void_p), ("tri_index", C.c_void_p), ("tri_material", C.c_void_p),
        ("tri_uv", C.c_void_p), ("tri_normal", C.c_void_p), ("divisions", C.c_int32), ("planes", C.c_void_p),
        ("cell_start", C.c_void_p), ("cell_list", C.c_void_p), ("mat_size", C.c_void_p), ("mat_start", C.c_void_p),
        ("texels", C.c_void_p), ("light_count", C.c_uint32), ("light_type", C.c_void_p), ("light_pos", C.c_void_p),
        ("light_dir", C.c_void_p), ("light_colour", C.c_void_p), ("light_radius", C.c_void_p), ("light_half", C.c_void_p),
        ("out_r", C.c_void_p), ("out_g", C.c_void_p), ("out_b", C.c_void_p), ("primary_id", C.c_void_p),
    ]


def build(force: bool = False) -> Path:
    if LIB.is_file() and not force and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    OUT.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-std=gnu11", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", str(SRC), "-o", str(LIB), "-lm",
                    "-lpthread"], check=True)
    return LIB


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.port_job_size.restype = C.c_size_t
        assert _lib.port_job_size() == C.sizeof(Job), "port_job layout mismatch"
        _lib.port_render_rows.restype = None
        _lib.port_render_rows.argtypes = [C.POINTER(Job), C.c_uint32, C.c_uint32, C.c_int]
        _lib.port_render_rows_step.restype = None
        _lib.port_render_rows_step.argtypes = [C.POINTER(Job), C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        _lib.port_kat_rand.restype = C.c_float
        _lib.port_kat_rand.argtypes = [C.POINTER(C.c_uint64), C.c_float, C.c_float]
        _lib.port_kat_hit.restype = C.c_int
        _lib.port_kat_hit.argtypes = [C.c_void_p] * 2 + [C.c_float] * 2 + [C.c_void_p] * 4
        _lib.port_kat_ball.restype = None
        _lib.port_kat_ball.argtypes = [C.POINTER(C.c_uint64), C.c_float, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def render(camera, lists, scene, samples: int = 1, threads: int | None = None, rows=None, want_ids: bool = False, row_step: int = 1):
    """CPU render of rows [rows[0], rows[1]) -- every row_step-th of them -- with the C port.  Returns (r, g, b) uint16 [H,W]
    (+ ids uint32 [H,W])."""
    lib = load()
    h, w = camera.height, camera.width
    r0, r1 = rows if rows is not None else (0, h)
    out = [np.zeros((h, w), np.uint16) for _ in range(3)]
    ids = np.full((h, w), 0xFFFFFFFF, np.uint32) if want_ids else None
    start = np.ascontiguousarray(lists.start, np.uint32)
    end = np.ascontiguousarray(lists.end, np.uint32)
    lst = np.ascontiguousarray(lists.list, np.uint32)
    j = Job()
    j.width, j.height = w, h
    for name, v in (("eye", camera.eye), ("top_left", camera.eye_to_top_left), ("step_right", camera.left_to_right),
                    ("step_down", camera.top_to_bottom)):
        setattr(j, name, V3(float(v[0]), float(v[1]), float(v[2])))
    j.pixel_size_inv = float(camera.pixel_size_inv)
    j.cam_start, j.cam_end, j.cam_list = _p(start), _p(end), _p(lst)
    j.samples = samples
    j.vertex, j.tri_index, j.tri_material = _p(scene.vertex), _p(scene.tri_idx), _p(scene.tri_mat)
    j.tri_uv, j.tri_normal = _p(scene.tri_uv), _p(scene.tri_normal)
    j.divisions = scene.axes_div
    j.planes, j.cell_start, j.cell_list = _p(scene.box_min), _p(scene.grid_start), _p(scene.grid_list)
    j.mat_size, j.mat_start, j.texels = _p(scene.mat_size), _p(scene.mat_start), _p(scene.textures)
    j.light_count = scene.light_count
    j.light_type, j.light_pos, j.light_dir = _p(scene.light_type), _p(scene.light_pos), _p(scene.light_dir)
    j.light_colour, j.light_radius, j.light_half = _p(scene.light_colour), _p(scene.light_radius), _p(scene.light_half)
    j.out_r, j.out_g, j.out_b = _p(out[0]), _p(out[1]), _p(out[2])
    j.primary_id = _p(ids) if want_ids else None
    lib.port_render_rows_step(C.byref(j), r0, r1, row_step, threads or (os.cpu_count() or 1))
    return (out[0], out[1], out[2], ids) if want_ids else tuple(out)


def rand_sequence(seed: int, count: int, lo: float = 0.0, hi: float = 1.0):
    lib = load()
    s = C.c_uint64(seed)
    return np.array([lib.port_kat_rand(C.byref(s), lo, hi) for _ in range(count)], np.float32), int(s.value)
