#!/usr/bin/env python3
"""Sweep the scheduling weights of wf_trace2_kernel (fresh process per setting: they are read once per process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
combos = [tuple(int(x) for x in c.split(",")) for c in sys.argv[2:]] or [(2, 3, 6), (1, 1, 1), (1, 2, 4), (1, 2, 8), (1, 3, 8), (1, 4, 8), (2, 3, 16), (1, 1, 4)]
for ww, wt, wsw in combos:
    env = dict(os.environ, OCLR_W_WALK=str(ww), OCLR_W_TEST=str(wt), OCLR_W_SWITCH=str(wsw))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ab_variants.py"), cfg, "2"], env=env, capture_output=True, text=True)
    print(f"--- wWalk {ww} wTest {wt} wSwitch {wsw}\n" + out.stdout.strip() + (out.stderr[-300:] if out.returncode else ""), flush=True)
