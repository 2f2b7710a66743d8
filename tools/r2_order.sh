#!/bin/bash
# A/B of the costly-arenas-first launch order (AGAR_NO_ORDER=1 switches it off) at several episode ages
S='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],1))'
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --steps 200 --warmup 10"
for age in ${1:-0 1000 2000}; do
  AGAR_NO_ORDER=1 $B --age $age 2>/dev/null | python -c "$S" "identity order age$age"
  $B --age $age 2>/dev/null | python -c "$S" "costly first   age$age"
done
