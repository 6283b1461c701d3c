"""Calling-convention check of include/oclr_abi.h: a C caller compiled against OUR header drives the REFERENCE's compiled
`RaytraceAll` (oracle/_ref) with by-value OpenCL vector unions; the image must equal the pointer-door result.  (CPU only.)"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import helpers

ROOT = Path(__file__).resolve().parent.parent


def test_header_calling_convention_matches_reference(ref):
    import ref as refmod
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    lib = out / "libabi_caller.so"
    subprocess.run(["gcc", "-O1", "-std=gnu11", "-fPIC", "-shared", str(ROOT / "tests" / "abi_caller.c"), str(refmod.LIB),
                    f"-Wl,-rpath,{refmod.LIB.parent}", "-o", str(lib)], check=True)
    caller = C.CDLL(str(lib))
    sc, cam, lists, samples = helpers.make_case("soup_s4")
    want = ref.raytrace_all(cam, lists, sc, samples)
    h, w = cam.height, cam.width
    got = [np.zeros((h, w), np.uint16) for _ in range(3)]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    f4 = lambda v: np.ascontiguousarray(v, np.float32)
    dim = np.array([w, h], np.uint32)
    keep = [f4(cam.eye), f4(cam.eye_to_top_left), f4(cam.left_to_right), f4(cam.top_to_bottom)]
    caller.abi_call_by_value.restype = C.c_uint32
    ok = caller.abi_call_by_value(p(dim), p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), C.c_float(cam.pixel_size_inv),
                                  p(lists.start), p(lists.end), p(lists.list), C.c_uint32(samples), p(sc.vertex),
                                  C.c_uint32(sc.triangle_count), p(sc.tri_idx), p(sc.tri_mat), p(sc.tri_uv), p(sc.tri_normal),
                                  C.c_int32(sc.axes_div), p(sc.box_min), p(sc.grid_start), p(sc.grid_list), p(sc.mat_size),
                                  p(sc.mat_start), p(sc.textures), C.c_uint32(sc.light_count), p(sc.light_type), p(sc.light_pos),
                                  p(sc.light_dir), p(sc.light_colour), p(sc.light_radius), p(sc.light_half), p(got[0]), p(got[1]),
                                  p(got[2]))
    assert ok == 1
    for c in range(3):
        assert np.array_equal(got[c], want[c])
    assert int((got[0] > 0).sum()) > 100


def test_host_helpers_equal_reference_build(ref):
    """raytrace.h:37-44 (dot, cross, normalize, vector, bindf, GetPointToLineSqLen, RayIntersectsTriangle, GetBoxAddress): the
    library's exports, called BY VALUE through include/oclr_abi.h's types, return bit for bit what the reference build returns on
    20 000 pseudo-random inputs and the boundary cases (t == min / max, abL + acL == 1, parallel ray, degenerate triangle, NaN,
    positions on / outside the split planes)."""
    import ref as refmod
    from opencl_render_b200 import _lib
    _lib.load()
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "helpers_kat"
    subprocess.run(["gcc", "-O1", "-std=gnu11", str(ROOT / "tests" / "helpers_kat.c"), "-o", str(exe), "-ldl", "-lm"], check=True)
    mine = subprocess.run([str(exe), str(_lib.LIB_PATH)], capture_output=True, check=True).stdout
    theirs = subprocess.run([str(exe), str(refmod.LIB)], capture_output=True, check=True).stdout
    assert len(mine) == len(theirs) and len(mine) > 20000 * 60
    a, b = np.frombuffer(mine, np.uint32), np.frombuffer(theirs, np.uint32)
    differ = a != b
    if differ.any():      # the only tolerated difference: two NaNs with different payload / sign bits
        fa, fb = a.view(np.float32)[differ], b.view(np.float32)[differ]
        assert np.isnan(fa).all() and np.isnan(fb).all(), int(differ.sum())
    assert (a.view(np.float32)[np.isfinite(a.view(np.float32))] != 0).sum() > 100000


def test_plain_c_example_builds_against_the_header_and_links():
    """examples/raytrace_all_min.c: the whole boundary from C -- helpers, builders, device list, RaytraceAll by value, image writer --
    compiles against include/oclr_abi.h and links against the library.  Without a GPU it must stop at the device list with exit code
    3 (no CPU path); with one it renders."""
    from opencl_render_b200 import _lib
    lib = _lib.load()
    out = ROOT / "tests" / "_build"
    out.mkdir(exist_ok=True)
    exe = out / "raytrace_all_min"
    subprocess.run(["gcc", "-O1", "-std=gnu11", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / "raytrace_all_min.c"),
                    f"-L{_lib.LIB_PATH.parent}", "-lopencl_render_b200", f"-Wl,-rpath,{_lib.LIB_PATH.parent}", "-lm", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), str(out / "example.bmp")], capture_output=True, text=True, timeout=300)
    if lib.oclr_device_count() == 0:
        assert r.returncode == 3 and "no CUDA device" in r.stderr
    else:
        assert r.returncode == 0, r.stderr
        assert (out / "example.bmp").stat().st_size == 54 + 320 * 240 * 3
