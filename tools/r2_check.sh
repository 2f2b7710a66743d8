#!/bin/bash
# round-2 check on one box: GPU parity suite, long lockstep + fuzz parity, then bench lines (young / aged cfg2, cfg1, cfg4)
# usage: tools/r2_check.sh [quick|full] [extra bench worlds...]
MODE=${1:-quick}
S='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]/1e6,2), "M  frac", round(d["roofline"]["frac"],3), "us", round(d["roofline"]["launch_us"],1))'
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --steps 200 --warmup 10"
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
if [ $MODE = full ]; then
  python tests/long_parity.py 2>&1 | tail -6
  python tests/fuzz_parity.py --configs 150 --steps 40 --seed $RANDOM 2>&1 | tail -4
fi
for age in 0 1000; do $B --age $age 2>/dev/null | python -c "$S" "cfg2 age$age"; done
for w in cfg1 cfg4; do $B --world $w --age 0 2>/dev/null | python -c "$S" "$w age0"; done
