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
        g = phase(l) or ("next_entry" if 295 <= l <= 312 else "rt_trace other")
    elif cur == "rt_walk.h":
        g = "walk (rt_walk.h step/bit/rank)" if l <= 116 else "switch (rt_walk.h)"
    elif cur == "rt_core.h":
        g = "tri_test (rt_core.h)" if (115 <= l <= 137 or l == 41) else ("switch (refine_axis)" if 322 <= l <= 345 else "rt_core other")
    else:
        g = cur
    a = agg.setdefault(g, [0, 0, 0]); a[0] += inst; a[1] += thr; a[2] += samp
tot = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print("total warp instructions %.3e" % tot)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-34s %5.1f%% inst  %5.1f%% samples  act %.1f" % (k, 100 * a[0] / tot, 100 * a[2] / max(ts, 1), a[1] / max(a[0], 1)))
