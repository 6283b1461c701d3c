#!/usr/bin/env python3
"""Tuning knobs of the trace kernel on one rank's 1/WORLD band share: python scripts/share_sweep.py CFG WORLD "A=1 B=2" ... (fresh process each)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, world = sys.argv[1], sys.argv[2]
for setting in sys.argv[3:]:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("="); env[k] = v
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_target_band.py"), cfg, world, "6"], env=env, capture_output=True, text=True)
    best = min((l for l in out.stdout.strip().splitlines() if " ms" in l), key=lambda l: float(l.split(":")[1].replace("BEST", "").split("ms")[0]),
               default=out.stderr[-300:])
    print(f"--- {setting}: {best}", flush=True)
