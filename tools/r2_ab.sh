#!/bin/bash
# A/B of library builds on one box, young and aged cfg2: tools/r2_ab.sh "<tag> ..." [rounds]
TAGS=$1; R=${2:-2}
S='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]/1e6,2), "M  us", round(d["roofline"]["launch_us"],1))'
for r in $(seq $R); do for age in 0 1000; do for t in $TAGS; do
  if [ $t = cur ]; then L=agarai_b200/libagar_b200.so; else L=agarai_b200/_ab/$t.so; fi
  AGAR_B200_LIB=$PWD/$L python bench.py --no-cpu-baseline --e2e-steps 2 --steps 200 --warmup 10 --no-extras --age $age 2>/dev/null | python -c "$S" "$t age$age"
done; done; done
