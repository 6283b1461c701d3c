"""The C-ABI library loads and exports every symbol include/oclr_abi.h declares; without a GPU every compute entry point
fails loudly (no CPU fallback).  No compute calls are made here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from opencl_render_b200 import _lib, api, scenes

ROOT = Path(__file__).resolve().parent.parent


def declared_functions():
    text = (ROOT / "include" / "oclr_abi.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set()
    for m in re.finditer(r"^[A-Za-z_][A-Za-z0-9_ \*]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", text, flags=re.M):
        name = m.group(1)
        if name not in ("defined", "__attribute__", "aligned", "vector_size", "OCLR_ALIGNED"):
            names.add(name)
    return names


def test_header_symbols_exported():
    lib = _lib.load()
    names = declared_functions()
    assert {"RaytraceAll", "InitOpenCL", "GetProgress", "oclr_scene_create", "oclr_frame_render", "GetBoxAddress"} <= names
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, f"declared in oclr_abi.h but not exported: {missing}"
    assert set(_lib.EXPORTS) <= names | {"oclr_version"}


def test_reference_symbol_names_present():
    # every function of source/opencl/raytrace.h:37-106
    lib = _lib.load()
    for n in ["dot", "cross", "normalize", "vector", "bindf", "GetPointToLineSqLen", "RayIntersectsTriangle", "GetBoxAddress",
              "InitOpenCL", "ResetComputationType", "GetIsComputationTypeUpdated", "GetComputationTypeCount",
              "GetComputationTypeName", "GetProgress", "SetProgress", "GetStartTime", "GetEndTime", "ResetTime", "RaytraceAll"]:
        assert hasattr(lib, n), n


def test_progress_and_time_cells():
    lib = _lib.load()
    lib.SetProgress(C.c_float(0.25))
    assert abs(lib.GetProgress() - 0.25) < 1e-7
    lib.ResetTime()
    assert lib.GetStartTime() == 0 and lib.GetEndTime() == 0


def test_computation_type_list_and_cpu_label():
    names = api.computation_types()
    assert names[0] == "Local CPU single thread"        # raytrace.c:136-143: index 0 keeps its label
    lib = _lib.load()
    assert lib.GetIsComputationTypeUpdated() == 1
    assert lib.GetComputationTypeCount() == len(names)
    buf = C.create_string_buffer(4)
    assert lib.GetComputationTypeName(0, 3, buf) == 0     # too short -> CL_FALSE
    n = len(names[0])
    buf = C.create_string_buffer(b"\xAA" * (n + 2), n + 2)
    assert lib.GetComputationTypeName(0, n, buf) == 0      # fits only WITHOUT its terminator: refused, nothing written past the buffer
    assert buf.raw == b"\xAA" * (n + 2)
    assert lib.GetComputationTypeName(0, n + 1, buf) == 1 and buf.raw[:n + 1] == names[0].encode() + b"\0" and buf.raw[n + 1] == 0xAA
    lib.ResetComputationType()
    assert lib.GetIsComputationTypeUpdated() == 0


def test_band_partition_covers_rows_once():
    for h, world in [(1080, 1), (1080, 2), (2160, 8), (100, 4), (128, 3), (4320, 8)]:
        seen = np.zeros(h, np.int32)
        for r in range(world):
            for b, e in api.band_partition(h, r, world):
                assert 0 <= b < e <= h and (b % 128 == 0)
                seen[b:e] += 1
        assert (seen == 1).all()


def test_cpu_type_is_refused_and_no_cpu_fallback():
    sc = scenes.soup(20, seed=1)
    cam = api.set_camera((0, 4.4, -8), (0, 0, 0), (0, 1, 0), 0.9, 16, 16)
    lists = api.camera_triangle_list(cam, sc)
    api.scene_triangle_list(sc, 16)
    with pytest.raises(api.OclrError, match="no CPU fallback|not implemented"):
        api.raytrace_all(0, cam, lists, 1, sc)
    if _lib.load().oclr_device_count() == 0:
        with pytest.raises(api.OclrError, match="no CUDA device|no such CUDA device"):
            api.raytrace_all(1, cam, lists, 1, sc)
        with pytest.raises(api.OclrError):
            api.DeviceScene(sc, 0)
