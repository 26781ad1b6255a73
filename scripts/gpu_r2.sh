#!/bin/bash
# Round-2 GPU-box session: parity tests, smoke, both bench arms, launch list and full ncu captures of the
# batched pipeline's kernels.  Usage (under gpurun, from the repo root):  bash scripts/gpu_r2.sh [tests|bench|ncu ...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WHAT="${*:-tests bench ncu}"
NCU="ncu --clock-control none"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for w in $WHAT; do
case $w in
tests)
  echo "== pytest -m gpu"
  timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
  echo "== smoke"
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
  ;;
bench)
  echo "== bench reference"
  timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/bench_ref.err | tail -1 | tee gpurun_out/bench_ref.json
  echo "== bench b200"
  timeout 900 python bench.py 2>gpurun_out/bench_b200.err | tail -1 | tee gpurun_out/bench_b200.json
  tail -5 gpurun_out/bench_b200.err
  ;;
ncu)
  echo "== ncu"
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_bench.csv \
      python bench.py --steps 3 --warmup 3 --rounds 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
  timeout 600 $NCU --set full --import-source on -k regex:"pyr_|detect|cap32|describe" -s 5 -c 5 -f -o gpurun_out/prof_extract \
      python scripts/r2_prof.py 16 3 > gpurun_out/prof_extract.log 2>&1
  timeout 600 $NCU --set full --import-source on -k regex:"t3_" -s 3 -c 3 -f -o gpurun_out/prof_match \
      python scripts/prof_one.py match > gpurun_out/prof_match.log 2>&1
  ls -la gpurun_out/*.ncu-rep
  ;;
esac
done
