#!/usr/bin/env python3
"""stdin: `ncu -i rep --page source --csv` (SASS view).  Prints executed warp-instructions and stall samples per opcode."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(sys.stdin))
while rows and not any(h.strip() == "Source" for h in rows[0]):     # skip the kernel-name preamble of the source page
    rows.pop(0)
if not rows:
    print("-- opcode mix: no source page --")
    sys.exit(0)
hdr = rows[0]


def col(*names):
    for n in names:
        for i, h in enumerate(hdr):
            if h.strip().lower() == n.lower():
                return i
    return None


c_src = col("Source")
c_exec = col("# Warp Instructions Executed", "Instructions Executed", "Warp Instructions Executed")
c_samp = col("# Samples", "Sampling Data (All)", "Samples")
if c_src is None or c_exec is None:
    print("-- opcode mix: columns not found:", hdr[:12])
    sys.exit(0)
ex, st = defaultdict(float), defaultdict(float)
for r in rows[1:]:
    if len(r) <= max(c_src, c_exec):
        continue
    parts = r[c_src].replace("@!", "@").split()
    if not parts:
        continue
    op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "IMAD", "SHFL", "LDS", "STS", "LDG", "STG", "BAR", "UBLKCP")) else op.split(".")[0]
    try:
        ex[op] += float(r[c_exec].replace(",", "") or 0)
        if c_samp is not None:
            st[op] += float(r[c_samp].replace(",", "") or 0)
    except ValueError:
        pass
tot, tots = sum(ex.values()) or 1.0, sum(st.values()) or 1.0
print(f"-- executed warp-instructions per opcode (total {tot:.0f}; stall samples {tots:.0f}) --")
for op, v in sorted(ex.items(), key=lambda kv: -kv[1])[:32]:
    print(f"   {op:14s} {v:14.0f} {100 * v / tot:6.2f}%   stall samples {100 * st[op] / tots:6.2f}%")
