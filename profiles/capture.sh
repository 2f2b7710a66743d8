#!/bin/bash
# Captures the two ncu artefacts the profiles/ summaries are made from (run on the GPU box via gpurun):
#   gpurun_out/launches.csv  per-launch durations of the default bench command
#   gpurun_out/prof.ncu-rep  one --set full capture (with source) of a warmed-up agar_step_kernel launch
# usage: bash profiles/capture.sh [world]
W=${1:-cfg2}
set -e
python bench.py --world $W --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$W.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --world $W --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agar_step_kernel -s 12 -c 1 -f -o gpurun_out/prof \
    python bench.py --world $W --steps 5 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_$W.json
