#!/bin/bash
# stall breakdown of one warmed-up cfg2 step launch for library builds: tools/r2_ncu_stalls.sh "<tag> ..." [age]
AGE=${2:-0}
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
for s in barrier long_scoreboard short_scoreboard wait branch_resolving no_instruction not_selected mio_throttle lg_throttle math_pipe_throttle dispatch_stall drain imc_miss membar sleeping tex_throttle; do
  M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio
done
for t in $1; do
  if [ $t = cur ]; then L=agarai_b200/libagar_b200.so; else L=agarai_b200/_ab/$t.so; fi
  echo "== $t age $AGE"
  AGAR_B200_LIB=$PWD/$L ncu --metrics $M --clock-control none -k regex:agar_step_kernel -s $((AGE + 12)) -c 1 \
    python bench.py --no-cpu-baseline --e2e-steps 2 --steps 5 --warmup 10 --no-extras --age $AGE 2>&1 | grep -E "^\s+(gpu__|smsp__)" | sed -e 's/smsp__average_warps_issue_stalled_//' -e 's/_per_issue_active.ratio//' | awk '{printf "%-60s %s\n", $1, $NF}'
done
