#!/bin/bash
# the two captures capture_r02.sh did not reach within its time limit (young cfg2 launch, GAE at [128, 65536])
mkdir -p gpurun_out
B="python bench.py --no-cpu-baseline --e2e-steps 2"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agar_step_kernel -s 12 -c 1 -f -o gpurun_out/r02_cfg2_young \
    $B --steps 5 --warmup 10 --no-extras --age 0 > gpurun_out/r02_ncu_young.log 2>&1
python tools/bench_kernels.py --only gae > gpurun_out/r02_gae_plain.json 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:agar_gae_tma -s 5 -c 1 -f -o gpurun_out/r02_gae \
    python tools/bench_kernels.py --only gae --reps 5 > gpurun_out/r02_ncu_gae.log 2>&1
cat gpurun_out/r02_gae_plain.json
