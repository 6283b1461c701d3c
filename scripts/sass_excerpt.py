#!/usr/bin/env python3
"""SASS evidence for the dominant kernel (runs on the CPU box: cuobjdump reads the built library).

    python scripts/sass_excerpt.py [kernel-substring] > profiles/rNN_wf_pipe_sass.txt

Prints, for every matching kernel of libopencl_render_b200.so: instruction count, the mnemonic histogram, and every global / shared
memory, atomic, vote / shuffle and MUFU instruction line -- the lines that show the 128-bit fetches (LDG.E.128), the shared-memory
queues (LDS / STS / ATOMS), the warp votes and shuffles used for compaction, and the IEEE division sequences (MUFU.RCP + FFMA)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opencl_render_b200", "libopencl_render_b200.so")
want = sys.argv[1] if len(sys.argv) > 1 else "wf_pipe_kernelILb0ELb0"
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
for b in blocks[1:]:
    name = b.split("\n", 1)[0].strip()
    if want not in name:
        continue
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    lines = [l for l in b.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
    ops = []
    for l in lines:
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append(m.group(1))
    print(f"== {demangled}\n   {len(ops)} SASS instructions")
    hist = collections.Counter(o.split(".")[0] for o in ops)
    print("   by opcode: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(28)))
    full = collections.Counter(o for o in ops if o.split(".")[0] in ("LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "ATOM", "MUFU", "VOTE", "SHFL",
                                                                      "MATCH", "LDC", "LDL", "STL", "BAR", "WARPSYNC", "POPC", "FLO", "BREV"))
    print("   memory / vote / special: " + ", ".join(f"{k} {v}" for k, v in sorted(full.items())))
    print("   -- lines:")
    for l in lines:
        if re.search(r"\b(LDG|STG|ATOMS|ATOMG|RED|ATOM)\b|\bLDG\.|\bSTG\.|ATOMS\.|MATCH|VOTE|SHFL|MUFU|LDL|STL", l):
            print("   " + re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l).strip())
