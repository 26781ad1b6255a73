#!/bin/bash
# A/B of the detector variants: parity tests, stage times from bench.py, instruction counts from ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -8
for nc in 4 8; do
  echo "== NC=$nc"
  CUDASIFT_DETECT_NC=$nc timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/bench_nc$nc.json
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_nc$nc.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "stage_ms", d["roofline"]["stage_ms"])
PY
  CUDASIFT_DETECT_NC=$nc timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_elapsed.avg -k regex:detect_kernel -s 3 -c 1 python scripts/prof_one.py extract 2>&1 | grep -E "detect_kernel|gpu__time|inst_executed|issue_active|cycles_" 
done
