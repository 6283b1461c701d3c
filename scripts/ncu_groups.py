"""Instruction shares of the trace kernel by phase (source-line ranges are looked up from marker comments in rt_trace.cuh)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]; kernel = sys.argv[2] if len(sys.argv) > 2 else "wf_pipe_kernel"
src = open("opencl_render_b200/csrc/rt_trace.cuh").read().split("\n")
# find line ranges inside the kernel by markers
start = next(i for i, l in enumerate(src) if kernel + "(" in l and "__global__" in "".join(src[max(0,i-1):i+1])) + 1
marks = []
for i in range(start, len(src)):
    m = re.match(r"\s*// ---- ([A-Za-z]+)", src[i])
    if m: marks.append((i + 1, m.group(1).lower()))
    if src[i].startswith("}"): end = i + 1; break
# function a source line belongs to: the last `OCLR_HD... name(` / `__device__ ... name(` definition at or above it
_defs = {}
def fn_of(fname, line):
    if fname not in _defs:
        d = []
        for i, l in enumerate(open("opencl_render_b200/csrc/" + fname).read().split("\n")):
            m = re.match(r"(?:template <[^>]*>\s*)?(?:OCLR_HD\w*(?:\(\d\))?|__device__ __forceinline__|__global__)\s+[\w:<> ]*?\b(\w+)\(", l)
            if m: d.append((i + 1, m.group(1)))
        _defs[fname] = d
    name = None
    for l0, n in _defs[fname]:
        if l0 <= line: name = n
    return name
def phase(line):
    if line < start or line > end: return None
    p = "prologue"
    for l, name in marks:
        if line >= l: p = name
    return p
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1].split('/')[-1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or not r or r[0] == "": continue
    d = dict(zip(hdr, r))
    try: inst = int(d["Instructions Executed"]); thr = int(d["Thread Instructions Executed"]); samp = int(d["# Samples"])
    except Exception: continue
    l = int(d["Line No"])
    if cur == "rt_trace.cuh":
        g = phase(l) or ("next_entry" if fn_of("rt_trace.cuh", l) == "next_entry" else "rt_trace other")
    elif cur == "rt_walk.h":
        f = fn_of("rt_walk.h", l)
        g = "switch (rt_walk.h)" if f in ("pwalk_enter_coarse", "prefine_axis", "pwalk_refine", "pwalk_super_allowed") else "walk (rt_walk.h step/bit/rank)"
    elif cur == "rt_core.h":
        g = "tri_test (rt_core.h)" if fn_of("rt_core.h", l) in ("tri_test", "dot3") else "rt_core other"
    else:
        g = cur
    a = agg.setdefault(g, [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += samp
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("total warp instructions %.3e" % tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %5.1f%% inst  %5.1f%% samples  act %.1f" % (k, 100 * a[0] / tot, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
