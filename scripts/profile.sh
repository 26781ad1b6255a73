#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the top kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_one.csv \
    python scripts/prof_one.py both > gpurun_out/prof_one.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:detect_kernel -s 3 -c 2 -f -o gpurun_out/prof_detect \
    python scripts/prof_one.py extract > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"lowpass_kernel|scaledown_kernel|describe_kernel" -s 18 -c 6 -f -o gpurun_out/prof_pyramid \
    python scripts/prof_one.py extract > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"t3_" -s 3 -c 3 -f -o gpurun_out/prof_match \
    python scripts/prof_one.py match > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
