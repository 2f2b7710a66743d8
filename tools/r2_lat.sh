#!/bin/bash
# launch time against the number of arenas (one CTA per SM ... several rounds) for the old and the compact step kernel
S='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d["value"]/1e6,2), "M  us", round(d["roofline"]["launch_us"],1))'
B="python bench.py --no-cpu-baseline --no-extras --e2e-steps 2 --steps 100 --warmup 10 --age ${AGE:-0}"
for n in $1; do
  $B --arenas $n 2>/dev/null | python -c "$S" "old n=$n"
  for nt in ${NTS:-32 64}; do AGAR_STEP2=$nt $B --arenas $n 2>/dev/null | python -c "$S" "step2/$nt n=$n"; done
done
