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
    stale = LIB.is_file() and reference_available() and shim.stat().st_mtime > LIB.stat().st_mtime   # our shim changed
    if LIB.is_file() and not force and not stale:
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

    # 3. link; the ICD loader only resolves the cl* symbols of the unused OpenCL branch
    run(["g++", "-shared", "-o", LIB, tmp / "raytrace.o", tmp / "trianglelist.o", tmp / "ref_shim.o",
         OPENCL_ICD, "-lm", "-lpthread"])
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.rmtree(gen, ignore_errors=True)
    shutil.rmtree(inc, ignore_errors=True)
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference not available and no prebuilt oracle/_ref found")
