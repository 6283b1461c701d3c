#!/usr/bin/env python3
"""Sweep the trace kernel's vote thresholds (each setting in a fresh process: they are read once per process)."""
import itertools, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
for wm, tm, rm in [(10, 10, 8), (16, 16, 4), (20, 16, 4), (16, 20, 4), (24, 24, 4), (12, 12, 2), (20, 20, 1), (16, 12, 4), (24, 12, 4), (28, 8, 4)]:
    env = dict(os.environ, OCLR_WALK_MIN=str(wm), OCLR_TEST_MIN=str(tm), OCLR_REFILL_MIN=str(rm))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_target.py"), cfg, "1", "4"], env=env, capture_output=True, text=True).stdout
    print(wm, tm, rm, "->", out.strip().splitlines()[-1], flush=True)
