#!/usr/bin/env python
"""Opcode histogram of executed warp instructions from an ncu source page:
  ncu -i X.ncu-rep --page source --csv --kernel-name K --print-source sass > k.csv ; python scripts/sass_hist.py k.csv"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, stall = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= ei: continue
    toks = r[si].split()
    if not toks: continue
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0] if len(sys.argv) < 3 else op
    n = int(r[ei] or 0)
    ops[op] += n; tot += n
    stall[op] += int(r[st] or 0)
ts = sum(stall.values())
print("total warp instructions executed:", tot, " samples:", ts)
for op, n in ops.most_common(40):
    print("%-12s %12d %6.2f%%   samples %6.2f%%" % (op, n, 100.0 * n / tot, 100.0 * stall[op] / max(ts, 1)))
