#!/bin/bash
# One GPU-box session: parity tests, smoke, golden fixtures, bench (both arms), launch list.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [quick]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" 
timeout 900 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.txt
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.txt
if [ "$1" != "quick" ]; then
echo "== golden"
timeout 300 python tests/golden/make_golden.py 2>&1 | tail -8
echo "== bench reference"
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
echo "== bench b200"
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_b200.json
fi
