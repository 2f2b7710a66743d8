#!/bin/bash
# Round-2 ncu artefacts (run on the GPU box via gpurun, after the same commands exited 0 without ncu):
#   gpurun_out/r02_launches.csv    per-launch durations of the default bench command (ageing launches included)
#   gpurun_out/r02_cfg2_aged.ncu-rep / r02_cfg2_young.ncu-rep   --set full of one warmed-up agar_step_kernel launch (cfg2, 4096 arenas)
#   gpurun_out/r02_gae.ncu-rep     --set full of one agar_gae_tma_kernel launch at [128, 65536]
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 2"
$B --steps 20 --warmup 5 > gpurun_out/r02_bench_plain.json 2> gpurun_out/r02_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches.csv \
    $B --steps 20 --warmup 5 > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agar_step_kernel -s 1012 -c 1 -f -o gpurun_out/r02_cfg2_aged \
    $B --steps 5 --warmup 10 --no-extras --age 1000 > gpurun_out/r02_ncu_aged.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:agar_step_kernel -s 12 -c 1 -f -o gpurun_out/r02_cfg2_young \
    $B --steps 5 --warmup 10 --no-extras --age 0 > gpurun_out/r02_ncu_young.log 2>&1
python tools/bench_kernels.py --only gae > gpurun_out/r02_gae_plain.json 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:agar_gae_tma -s 5 -c 1 -f -o gpurun_out/r02_gae \
    python tools/bench_kernels.py --only gae --reps 5 > gpurun_out/r02_ncu_gae.log 2>&1
ls -la gpurun_out
