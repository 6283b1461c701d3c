#!/usr/bin/env python3
"""Generic env sweep: python scripts/sweep_env.py CFG VARIANT "A=1 B=2" "A=3" ...  (fresh process per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg, variant = sys.argv[1], sys.argv[2]
for setting in sys.argv[3:]:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("="); env[k] = v
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ab_variants.py"), cfg, variant], env=env, capture_output=True, text=True)
    print(f"--- {setting}\n" + out.stdout.strip() + (out.stderr[-300:] if out.returncode else ""), flush=True)
