#!/bin/bash
# round-2 probe: young vs aged cfg2, 64 threads per arena, raster share, phase clocks (one gpurun call)
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --steps 200 --warmup 10"
S='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],1))'
for age in 0 1000 2000; do
  $B --age $age 2>/dev/null | python -c "$S" "t128 age$age"
  AGAR_THREADS=64 $B --age $age 2>/dev/null | python -c "$S" "t64 age$age"
done
for age in 0 1000; do $B --age $age --no-obs 2>/dev/null | python -c "$S" "t128 noobs age$age"; done
python tools/aged_stats.py 3000 4096
for age in 0 1500; do AGAR_B200_LIB=$PWD/agarai_b200/_ab/clocks.so python tools/phase_clocks.py 1036 cfg2 AGE=$age; done
AGAR_B200_LIB=$PWD/agarai_b200/_ab/clocks.so python tools/phase_clocks.py 148 cfg2 AGE=1500
