#!/bin/bash
# Matcher experiments (CS_TC_DBG bits, see TcPlan::dbg): per-kernel durations under ncu, then the real
# (warm-cache) per-kernel times from CUDA events and the bit-exact comparison against the exact path.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for dbg in ${1:-0}; do
  echo "== CS_TC_DBG=$dbg"
  CS_TC_DBG=$dbg timeout 120 ncu --clock-control none --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     -k regex:"t3_|match_exact" -s 8 -c 4 python scripts/prof_one.py match 2>/dev/null | grep -E '^"' | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
k=h.index('Kernel Name'); m=h.index('Metric Name'); v=h.index('Metric Value')
out={}
for r in rows[1:]:
    out.setdefault((r[0],r[k].split('(')[0]),{})[r[m]]=r[v]
for (i,n),d in out.items(): print('  %-22s %8s ns  tensor %6s %%'%(n,d.get('gpu__time_duration.sum'),d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')))
"
done
timeout 90 python scripts/tc_bench.py 2000 10000 2>&1 | tail -6
CUDASIFT_MATCH_TIMING=1 timeout 90 python scripts/tc_bench.py 10000 2>&1 | grep "match timing" | tail -2
