#!/usr/bin/env python3
"""Sweep the scheduling knobs of wf_trace3_kernel (fresh process per setting: they are read once per process).
   python scripts/sweep3.py CFG drain,walk,switch,refill ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
combos = [tuple(int(x) for x in c.split(",")) for c in sys.argv[2:]]
for dm, wm, sm, rm in combos:
    env = dict(os.environ, OCLR_DRAIN_MIN=str(dm), OCLR_WALK_MIN3=str(wm), OCLR_SWITCH_MIN=str(sm), OCLR_REFILL_MIN=str(rm))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ab_variants.py"), cfg, "2"], env=env, capture_output=True, text=True)
    print(f"--- drain {dm} walk {wm} switch {sm} refill {rm}\n" + out.stdout.strip() + (out.stderr[-300:] if out.returncode else ""), flush=True)
