#!/usr/bin/env python3
"""Builds libopencl_render_b200.so in-tree (opencl_render_b200/): nvcc for the CUDA runtime + kernels (sm_100a only,
-fmad=false so fp32 results equal the reference's non-FMA x86 build), g++/gcc -ffp-contract=off for the C-ABI, the
scene packer and the host builders.  No CPU compute path is compiled into the library.

    python -m opencl_render_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "csrc" / "build"
LIB = PKG / "libopencl_render_b200.so"
CUDA = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false", "-prec-div=true",
              "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-Xptxas", "-v"]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-Wno-unused-function", f"-I{CUDA / 'include'}", "-pthread"]
C_FLAGS = ["-O2", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-Wall"]

SOURCES = [("runtime.cu", "nvcc"), ("abi.cpp", "g++"), ("scene_pack.cpp", "g++"), ("builders.cpp", "g++"), ("image_io.cpp", "g++"),
           ("abi_helpers.c", "gcc")]


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.c*")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "oclr_abi.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build_variant(name: str, defs: str, verbose: bool = False) -> Path:
    """A/B build of the same library with extra nvcc defines (e.g. -DOCLR_CELLQ_CAP=128): libopencl_render_b200_<name>.so next to the
    production library, loaded through OCLR_LIB (developer knob of _lib.load)."""
    out = PKG / f"libopencl_render_b200_{name}.so"
    obj = OBJ / f"runtime_{name}.o"
    OBJ.mkdir(parents=True, exist_ok=True)
    build()
    stamp = OBJ / f"stamp_{name}.txt"
    dig = _digest() + defs
    if out.is_file() and stamp.is_file() and stamp.read_text() == dig:
        return out
    if not (CUDA / "bin" / "nvcc").is_file():
        if out.is_file():
            return out  # prebuilt variant travelled with the snapshot
        raise RuntimeError(f"nvcc not found and no prebuilt {out.name}")
    cmd = [str(CUDA / "bin" / "nvcc"), *NVCC_FLAGS, *defs.split(), "-c", str(CSRC / "runtime.cu"), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (OBJ / f"runtime_{name}.ptxas.txt").write_text(r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stderr}")
    objs = [str(obj)] + [str(OBJ / (n + ".o")) for n, _ in SOURCES if n != "runtime.cu"]
    r = subprocess.run([str(CUDA / "bin" / "nvcc"), "-shared", "-o", str(out), *objs, "-cudart", "static", "-Xlinker", "-Bsymbolic", "-lpthread", "-lz"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed\n{r.stderr}")
    stamp.write_text(dig)
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = OBJ / "stamp.txt"
    dig = _digest() + os.environ.get("OCLR_NVCC_DEFS", "")
    if LIB.is_file() and not force and stamp.is_file() and stamp.read_text() == dig:
        return LIB
    if not (CUDA / "bin" / "nvcc").is_file():
        if LIB.is_file():
            return LIB  # prebuilt library travelled with the snapshot
        raise RuntimeError("nvcc not found and no prebuilt libopencl_render_b200.so")
    OBJ.mkdir(parents=True, exist_ok=True)
    objs = []
    for name, tool in SOURCES:
        src = CSRC / name
        obj = OBJ / (name + ".o")
        if tool == "nvcc":
            extra = os.environ.get("OCLR_NVCC_DEFS", "").split()      # experiment knob, e.g. -DOCLR_TRACE_MIN_CTAS=10
            cmd = [str(CUDA / "bin" / "nvcc"), *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        elif tool == "g++":
            cmd = ["g++", *CXX_FLAGS, "-c", str(src), "-o", str(obj)]
        else:
            cmd = ["gcc", *C_FLAGS, "-c", str(src), "-o", str(obj)]
        if verbose:
            print("[build]", " ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if tool == "nvcc":
            (OBJ / (name + ".ptxas.txt")).write_text(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        objs.append(str(obj))
    cmd = [str(CUDA / "bin" / "nvcc"), "-shared", "-o", str(LIB), *objs, "-cudart", "static", "-Xlinker", "-Bsymbolic", "-lpthread", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return LIB


if __name__ == "__main__":
    if "--variant" in sys.argv:      # python -m opencl_render_b200.build --variant q128 "-DOCLR_CELLQ_CAP=128"
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2]))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv or "-v" in sys.argv))
