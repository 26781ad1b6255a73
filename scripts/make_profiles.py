#!/usr/bin/env python
"""Turn the ncu reports a GPU session left in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/make_profiles.py r01

writes profiles/<round>_<report>.txt (key metrics per kernel, from `ncu -i ... --page raw --csv`),
profiles/<round>_launches_*.csv (per-launch device times of the bench command) and a per-kernel
share table."""
import collections
import csv
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")


def shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    lines = ["%-44s %6s %10s %10s %7s" % ("kernel", "n", "mean_us", "total_us", "share")]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        lines.append("%-44s %6d %10.2f %10.1f %6.1f%%" % (k[:44], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    return "\n".join(lines)


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(SRC, "*.ncu-rep"))):
        name = os.path.splitext(os.path.basename(rep))[0]
        txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep],
                             stdout=subprocess.PIPE, text=True).stdout
        open(os.path.join(OUT, "%s_%s.txt" % (rnd, name)), "w").write(
            "# ncu --set full --clock-control none --import-source on  (%s.ncu-rep, B200)\n" % name + txt)
    for c in sorted(glob.glob(os.path.join(SRC, "launches_*.csv"))):
        name = os.path.splitext(os.path.basename(c))[0]
        shutil.copyfile(c, os.path.join(OUT, "%s_%s.csv" % (rnd, name)))
        open(os.path.join(OUT, "%s_%s_shares.txt" % (rnd, name)), "w").write(
            "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n"
            + shares(c) + "\n")
    for j in ("bench_b200.json", "bench_ref.json", "scale_2gpu.json", "allpairs_2gpu.json"):
        p = os.path.join(SRC, j)
        if os.path.exists(p):
            shutil.copyfile(p, os.path.join(OUT, "%s_%s" % (rnd, j)))
    print("profiles written to", OUT)


if __name__ == "__main__":
    main()
