"""Per-source-line instruction / stall-sample shares from an .ncu-rep captured with --import-source on (all profiled launches summed).
    python scripts/ncu_hot_lines.py x.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = {}
for r in csv.reader(io.StringIO(txt)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur = r[1]; continue
    if len(r) == 2: continue
    if r and r[0] == "Line No": hdr = r; continue
    if not hdr or not r or r[0] == "": continue
    d = dict(zip(hdr, r))
    try:
        inst = int(d["Instructions Executed"]); samp = int(d["# Samples"]); thr = int(d["Thread Instructions Executed"])
    except Exception:
        continue
    k = (cur.split('/')[-1], int(d["Line No"]))
    a = agg.setdefault(k, [0, 0, 0, r[1].strip()[:100]])
    a[0] += inst; a[1] += samp; a[2] += thr
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instructions %d, stall samples %d, avg active threads %.2f" % (tot, ts, sum(a[2] for a in agg.values()) / max(tot, 1)))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% samp  act %4.1f  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1), k[0], k[1], a[3]))
