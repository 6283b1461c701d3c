#!/usr/bin/env python3
"""Recipe that compiles the UNMODIFIED reference hot path into oracle/_ref/ (TEST INFRASTRUCTURE ONLY).

Nothing under oracle/ is product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may load what this script builds.

What is built (binaries only -- no reference source is copied into the repository):

  oracle/_ref/libref_raytrace.so
      /root/reference/source/opencl/raytrace.c compiled as C where it lies.  That file textually
      includes the OpenCL kernel source (raytrace.c:70 -> raytrace_opencl.c), so the .so exports the
      reference's own `RaytraceAll` (computationType 0 = "Local CPU single thread", raytrace.c:604-655),
      `Raytrace`, `RayIntersectsTriangle`, `RayIntersectsTriangles`, `randF`, `GetSpherePoint`,
      `GetTriangleNormal`, `Get2dTableValue3`, `BindInCube`, `GetBoxAddress`.
      It also holds /root/reference/source/util/trianglelist.cpp (the acceleration-list builders),
      compiled from a transient copy with the two g++-incompatible declarations at :566 and :581
      split into declaration + assignment (MSVC accepts the `goto` over an initialisation, g++ does
      not), plus oracle/ref_shim.cpp which gives the two C++ `New` factories a C calling convention.
      `SetCamera` (/root/reference/source/render.cpp:461-491, thirty lines of plain C inside a file that otherwise needs the
      Cinema4D SDK) is cut out of render.cpp into a transient translation unit at build time and linked in as well, behind
      `ref_set_camera` (pointer-only door).

  oracle/_ref/libref_raytrace_counted.so
      The same kernel with EVENT COUNTERS, for the algorithmic-bytes cross-check of SURVEY.md section 8d / Appendix C: a
      transient copy of raytrace_opencl.c (next to transient copies of raytrace.c / raytrace.h / raytrace_opencl.h, because
      raytrace.c:70 includes the kernel by a quoted relative name) gets `ref_cnt[k]++` statements at the places Appendix C
      names -- entry of RayIntersectsTriangle (:126), RayIntersectsTriangles (:347), every cell visited (:365, + empty cells),
      every grid candidate tested (:369), every ring segment (:510), every camera-list candidate tested (:517), every shaded
      hit (:532), every transparent-occluder lookup (:619).  Nothing else changes; the copies are deleted after compiling.

The recipe follows SURVEY.md section 8c / Appendix B:
  * `xxd -i` is a Windows pre-build step of the reference (opencl_render.vcxproj:83-86); the
    equivalent byte arrays are generated here because raytrace.c:7-8 includes them.
  * `<cl/cl.h>` is spelled lower-case in the reference (case-insensitive FS), so an include dir
    with a symlink `cl -> 3rdparty/opencl-1.2/include/CL` is created.
  * raytrace.h:16-31 redefines `int`, `float`, ... as macros, so the libc headers must be
    pre-included.
  * `-ffp-contract=off` and no -march flags: the bit reference is plain x86-64 SSE2 fp32
    (no FMA contraction).  The CUDA product is compiled with -fmad=false to match.

Run:  python oracle/build_ref.py [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("OCLR_REFERENCE_ROOT", "/root/reference"))
OPENCL_ICD = "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1"
LIB = OUT / "libref_raytrace.so"
LIB_COUNTED = OUT / "libref_raytrace_counted.so"

# (line after which the statement goes, text the line must contain, statement) -- raytrace_opencl.c, SURVEY.md Appendix C
COUNTER_NAMES = ["tests", "gridRays", "cells", "emptyCells", "segments", "primCandidates", "shadedHits", "gridCandidates",
                 "occluderLookups"]
COUNTER_SITES = [
    (126, "bool intersects = false;", "ref_cnt[0]++;"),
    (347, "uint closestTriangleIndex = (uint)-1;", "ref_cnt[1]++;"),
    (365, "longId = cubeId.x + axesDivCount * cubeId.y",
     "ref_cnt[2]++; if (scenePixelTriangleListStart[longId] == scenePixelTriangleListStart[longId + 1]) ref_cnt[3]++;"),
    (369, "if (excludedTriangleIndex != triangleIndex) {", "ref_cnt[7]++;"),
    (510, "float closestTriangleMult = maxDistance[cursorBegin];", "ref_cnt[4]++;"),
    (517, "if (excludedTriangleIndex[cursorBegin] != index) {", "ref_cnt[5]++;"),
    (532, "if ((uint)-1 != closestTriangleIndex) {", "ref_cnt[6]++;"),
]
OCCLUDER_LINE = (619, "transp = Get2dTableValue3(", "transp = (ref_cnt[8]++, Get2dTableValue3(")   # the guarded statement of the `if` at :618


def _set_camera_unit(src: Path) -> str:
    """SetCamera (render.cpp:461-491) as a translation unit of its own: the function's lines are taken from render.cpp at build
    time (never stored in the repository), in front of a pointer-only door for ctypes."""
    lines = (src / "render.cpp").read_text().splitlines(keepends=True)
    begin = next(i for i, l in enumerate(lines) if l.startswith("void SetCamera("))
    end = next(i for i in range(begin, len(lines)) if lines[i].startswith("}"))
    body = "".join(lines[begin:end + 1])
    assert "midToLeftLength" in body and end - begin < 40, "SetCamera moved: check render.cpp"
    return ('extern "C" {\n#include "opencl/raytrace.h"\n}\n' + body +
            'extern "C" void ref_set_camera(const cl_float* position, const cl_float* object, const cl_float* up, cl_float fov, cl_uint w, cl_uint h,\n'
            '                               cl_float* outTopLeft, cl_float* outLeftToRight, cl_float* outTopToBottom, cl_float* outPixelSizeInv) {\n'
            '    cl_float3 p, o, u, tl, lr, tb; cl_uint2 dim; dim.s[0] = w; dim.s[1] = h;\n'
            '    for (cl_int i = 0; i < 3; ++i) { p.s[i] = position[i]; o.s[i] = object[i]; u.s[i] = up[i]; }\n'
            '    p.s[3] = o.s[3] = u.s[3] = 0; tl.s[3] = lr.s[3] = tb.s[3] = 0;\n'
            '    SetCamera(&tl, &lr, &tb, outPixelSizeInv, p, o, u, fov, dim);\n'
            '    for (cl_int i = 0; i < 3; ++i) { outTopLeft[i] = tl.s[i]; outLeftToRight[i] = lr.s[i]; outTopToBottom[i] = tb.s[i]; }\n'
            '}\n')


def _counted_kernel(src: Path) -> str:
    lines = (src / "opencl" / "raytrace_opencl.c").read_text().splitlines(keepends=True)
    no, must, repl = OCCLUDER_LINE
    assert must in lines[no - 1], (no, lines[no - 1])
    line = lines[no - 1].replace(must, repl)
    k = line.rstrip().rfind(";")
    lines[no - 1] = line[:k] + ")" + line[k:]
    for no, must, stmt in sorted(COUNTER_SITES, reverse=True):      # bottom-up: earlier line numbers stay valid
        assert must in lines[no - 1], (no, lines[no - 1])
        lines.insert(no, stmt + "\n")
    return "extern __thread unsigned long long ref_cnt[16];\n" + "".join(lines)

PRE = ["unistd.h", "string.h", "stdio.h", "stdlib.h", "math.h", "time.h", "stddef.h"]


def _xxd_i(src: Path, name: str) -> str:
    data = src.read_bytes()
    body = ",".join(str(b) for b in data)
    return f"unsigned char {name}[] = {{{body}}};\nunsigned int {name}_len = {len(data)};\n"


def reference_available() -> bool:
    return (REF / "source" / "opencl" / "raytrace.c").is_file()


def build(force: bool = False, verbose: bool = True) -> Path | None:
    """Build oracle/_ref/libref_raytrace.so.  Returns its path, or None when /root/reference is absent
    (the GPU box: it only uses the prebuilt file that travelled with the snapshot)."""
    shim = Path(__file__).resolve().parent / "ref_shim.cpp"
    mine = max(shim.stat().st_mtime, Path(__file__).stat().st_mtime)                                 # our shim / this recipe changed
    have = LIB.is_file() and LIB_COUNTED.is_file()
    stale = have and reference_available() and mine > min(LIB.stat().st_mtime, LIB_COUNTED.stat().st_mtime)
    if have and not force and not stale:
        return LIB
    if not reference_available():
        return LIB if LIB.is_file() else None
    src = REF / "source"
    gen = OUT / "gen"
    inc = OUT / "inc"
    tmp = OUT / "tmp"
    for d in (gen / "opencl", inc, tmp):
        d.mkdir(parents=True, exist_ok=True)
    cl_link = inc / "cl"
    if not cl_link.exists():
        cl_link.symlink_to(src / "3rdparty" / "opencl-1.2" / "include" / "CL")
    # xxd -i equivalents (names fixed by raytrace.c:322-325)
    (gen / "opencl" / "raytrace_opencl.bin.h").write_text(
        _xxd_i(src / "opencl" / "raytrace_opencl.h", "source_opencl_raytrace_opencl_h"))
    (gen / "opencl" / "raytrace_opencl.bin.c").write_text(
        _xxd_i(src / "opencl" / "raytrace_opencl.c", "source_opencl_raytrace_opencl_c"))
    # stub for the Cinema4D SDK header: trianglelist.cpp only uses DebugAssert from it
    (gen / "c4d.h").write_text("#include <string.h>\n#include <math.h>\n#define DebugAssert(...) ((void)0)\n")

    common = ["-O2", "-ffp-contract=off", "-fPIC", "-w", "-DuSEC_PER_MSEC=1000"]
    for h in PRE:
        common += ["-include", h]
    incs = [f"-I{gen}", f"-I{inc}", f"-I{src / '3rdparty' / 'opencl-1.2' / 'include'}", f"-I{src}",
            f"-I{src / 'opencl'}", f"-I{src / 'util'}"]

    def run(cmd):
        if verbose:
            print("[oracle/_ref]", " ".join(str(c) for c in cmd), file=sys.stderr)
        subprocess.run([str(c) for c in cmd], check=True)

    # 1. the kernel + its host function, unmodified, compiled where it lies
    run(["gcc", "-std=gnu11", *common, *incs, "-c", src / "opencl" / "raytrace.c", "-o", tmp / "raytrace.o"])

    # 2. the builders: transient copy with the two MSVC-only lines split (deleted after compiling)
    tl = (src / "util" / "trianglelist.cpp").read_text().splitlines(keepends=True)
    patched = []
    for i, line in enumerate(tl, start=1):
        if i == 566:
            assert "cl_uint outputImageSize =" in line, line
            line = line.replace("cl_uint outputImageSize =", "cl_uint outputImageSize; outputImageSize =")
        if i == 581:
            assert "cl_uint compressionTriangleCount =" in line, line
            line = line.replace("cl_uint compressionTriangleCount =",
                                "cl_uint compressionTriangleCount; compressionTriangleCount =")
        patched.append(line)
    tl_copy = tmp / "trianglelist_transient.cpp"
    tl_copy.write_text("".join(patched))
    cxx = ["g++", "-std=gnu++14", "-fpermissive", *common, "-include", "map", "-include", "set",
           "-include", "utility", *incs]
    try:
        run([*cxx, "-c", tl_copy, "-o", tmp / "trianglelist.o"])
    finally:
        tl_copy.unlink(missing_ok=True)
    run([*cxx, "-c", HERE / "ref_shim.cpp", "-o", tmp / "ref_shim.o"])
    # 2b. SetCamera, cut out of render.cpp (transient)
    sc_copy = tmp / "setcamera_transient.cpp"
    sc_copy.write_text(_set_camera_unit(src))
    try:
        run([*cxx, "-c", sc_copy, "-o", tmp / "setcamera.o"])
    finally:
        sc_copy.unlink(missing_ok=True)

    # 3. link; the ICD loader only resolves the cl* symbols of the unused OpenCL branch
    run(["g++", "-shared", "-o", LIB, tmp / "raytrace.o", tmp / "trianglelist.o", tmp / "ref_shim.o", tmp / "setcamera.o",
         OPENCL_ICD, "-lm", "-lpthread"])

    # 4. the counting copy (Appendix C): transient instrumented kernel next to transient copies of its includers
    cdir = tmp / "counted"
    cdir.mkdir(parents=True, exist_ok=True)
    try:
        for name in ("raytrace.c", "raytrace.h", "raytrace_opencl.h"):
            shutil.copyfile(src / "opencl" / name, cdir / name)
        (cdir / "raytrace_opencl.c").write_text(_counted_kernel(src))
        # raytrace.c finds "opencl/raytrace_opencl.bin.h" through -I gen, everything else by relative name inside cdir
        cincs = [f"-I{gen}", f"-I{inc}", f"-I{src / '3rdparty' / 'opencl-1.2' / 'include'}", f"-I{cdir}"]
        run(["gcc", "-std=gnu11", *common, *cincs, "-c", cdir / "raytrace.c", "-o", tmp / "raytrace_counted.o"])
        run([*cxx, "-DREF_COUNTED", "-c", HERE / "ref_shim.cpp", "-o", tmp / "ref_shim_counted.o"])
        run(["g++", "-shared", "-o", LIB_COUNTED, tmp / "raytrace_counted.o", tmp / "trianglelist.o", tmp / "ref_shim_counted.o",
             tmp / "setcamera.o", OPENCL_ICD, "-lm", "-lpthread"])
    finally:
        shutil.rmtree(cdir, ignore_errors=True)
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.rmtree(gen, ignore_errors=True)
    shutil.rmtree(inc, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference not available and no prebuilt oracle/_ref found")
