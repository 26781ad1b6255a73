#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 = UTC*, TMEM = LDTM/STTM, TMA tensor
copies = UTMALDG/UTMASTG, bulk copies = UBLKCP, packed FP32 = FFMA2/FADD2/FMUL2) in the shipped library:
  python scripts/sass_grep.py > profiles/r02_sass_grep.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cudasift_b200", "lib", "libcudasift_b200.so")
KEYS = ["UTCHMMA", "UTCCP", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2",
        "FFMA", "LDGSTS", "LDS", "STS", "SHFL", "BAR", "TEX", "TLD4"]
out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
cur, tab = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
        tab[cur] = collections.Counter()
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        tab[cur][m.group(1)] += 1
        tab[cur]["_all"] += 1
print("# cuobjdump -sass cudasift_b200/lib/libcudasift_b200.so (sm_100a): static instruction counts per kernel")
print("%-44s %6s " % ("kernel", "instr") + " ".join("%7s" % k for k in KEYS))
tot = collections.Counter()
for k, c in tab.items():
    print("%-44s %6d " % (k[-44:], c["_all"]) + " ".join("%7d" % c[x] for x in KEYS))
    tot.update(c)
print("%-44s %6d " % ("TOTAL", tot["_all"]) + " ".join("%7d" % tot[x] for x in KEYS))
