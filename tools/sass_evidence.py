"""SASS instruction counts per kernel of the built library (no GPU needed):
    python tools/sass_evidence.py > profiles/r2_sass_evidence.txt
DMMA = FP64 warp-level MMA (tensor pipe), UTMALDG = tensor-map TMA load, UBLKCP = bulk TMA, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arnoldimethod.jl_b200", "libb200arnoldi.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
COLS = [("DMMA", r"\bDMMA"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"),
        ("proxy fence", r"FENCE\.VIEW\.ASYNC|\bFENCE\b.*PROXY"), ("DFMA", r"\bDFMA"), ("LDS", r"\bLDS"), ("LDG", r"\bLDG"),
        ("STG", r"\bSTG"), ("ATOM/RED", r"\b(ATOM|RED|ATOMG)\b"), ("UTC*MMA", r"\bUTC\w*MMA")]
blocks = out.split("Function : ")[1:]
rows = []
for blk, nm in zip(blocks, names):
    body = blk.split("\n", 1)[1] if "\n" in blk else ""
    cnt = [len(re.findall(rx, body)) for _, rx in COLS]
    rows.append((nm.strip(), cnt))
print("SASS evidence (cuobjdump -sass libb200arnoldi.so, sm_100a): instruction counts per kernel")
print("kernel | " + " | ".join(c for c, _ in COLS))
tot = collections.Counter()
for nm, cnt in rows:
    short = re.sub(r"\(.*", "", nm)
    print(short + " | " + " | ".join(str(c) for c in cnt))
    for (c, _), v in zip(COLS, cnt):
        tot[c] += v
print("TOTAL | " + " | ".join(str(tot[c]) for c, _ in COLS))
