#!/bin/bash
# Round-2 session-4 artefacts of the shipped step kernel (mass-decay check compiled in, rule off), run on the GPU box via gpurun;
# every ncu pass follows a plain run of the same command:
#   r03_launches.csv        per-launch durations of `bench.py --no-extras --age 200 --steps 20 --warmup 5`
#   r03_cfg2_aged.ncu-rep   --set full of launch 1013 of the aged cfg2 world
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 2"
$B --no-extras --age 200 --steps 20 --warmup 5 > gpurun_out/r03_plain.json 2> gpurun_out/r03_plain.err || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r03_launches.csv \
    $B --no-extras --age 200 --steps 20 --warmup 5 > gpurun_out/r03_ncu_launches.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:agar_step_kernel -s 1012 -c 1 -f -o gpurun_out/r03_cfg2_aged \
    $B --steps 5 --warmup 10 --no-extras --age 1000 > gpurun_out/r03_ncu_aged.log 2>&1
ls -la gpurun_out | tail -8
