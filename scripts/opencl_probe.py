#!/usr/bin/env python3
"""is there an OpenCL platform on this box on which the reference's own OpenCL branch
(raytrace.c:283-603) could run?  dlopen the ICD loader that ships with the CUDA toolkit, ask for platforms as the box is, then once
more with an ICD file that points at the NVIDIA driver's OpenCL library (if the driver ships one).  Prints what it finds; exit 0."""
import ctypes as C
import glob
import os
import subprocess
import sys
import tempfile

LOADER = "/usr/local/cuda/targets/x86_64-linux/lib/libOpenCL.so.1"


def platforms():
    code = f"""
import ctypes as C
cl = C.CDLL({LOADER!r})
n = C.c_uint(0)
rc = cl.clGetPlatformIDs(0, None, C.byref(n))
print("clGetPlatformIDs rc", rc, "platforms", n.value)
ids = (C.c_void_p * max(n.value, 1))()
if n.value:
    cl.clGetPlatformIDs(n.value, ids, None)
    for p in ids[:n.value]:
        buf = C.create_string_buffer(256)
        cl.clGetPlatformInfo(C.c_void_p(p), 0x0902, 256, buf, None)
        nd = C.c_uint(0)
        rc = cl.clGetDeviceIDs(C.c_void_p(p), 0xFFFFFFFF, 0, None, C.byref(nd))
        print("  platform", buf.value.decode(), "devices", nd.value, "rc", rc)
"""
    return code


def run(env_extra, label):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", platforms()], env=env, capture_output=True, text=True, timeout=120)
    print(f"[{label}] rc={r.returncode}\n{r.stdout}{r.stderr[-500:]}", flush=True)


print("loader present:", os.path.isfile(LOADER), "| /etc/OpenCL/vendors:", glob.glob("/etc/OpenCL/vendors/*"), flush=True)
libs = sorted(set(glob.glob("/usr/lib/x86_64-linux-gnu/libnvidia-opencl.so*") + glob.glob("/usr/lib64/libnvidia-opencl.so*") +
                  glob.glob("/usr/local/nvidia/lib64/libnvidia-opencl.so*") + glob.glob("/usr/lib/x86_64-linux-gnu/libpocl*")))
print("driver / PoCL OpenCL libraries:", libs, flush=True)
run({}, "as the box is")
if libs:
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "probe.icd"), "w") as f:
        f.write(libs[0] + "\n")
    run({"OCL_ICD_VENDORS": d, "OPENCL_VENDOR_PATH": d}, f"with an ICD file -> {libs[0]}")
