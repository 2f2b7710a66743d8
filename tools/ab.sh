#!/bin/bash
# A/B timing on one box: tools/ab.sh "<tag> <tag> ..." "<world> ..." [rounds] [steps]
# tag = name of agarai_b200/_ab/<name>.so ("cur" = the in-tree build), optionally "@<AGAR_MIN_BLOCKS>"
TAGS=$1; WORLDS=${2:-cfg2}; R=${3:-3}; STEPS=${4:-200}
for w in $WORLDS; do for r in $(seq $R); do for t in $TAGS; do
  n=${t%@*}; mb=; [[ $t == *@* ]] && mb=${t#*@}
  if [ $n = cur ]; then L=agarai_b200/libagar_b200.so; else L=agarai_b200/_ab/$n.so; fi
  AGAR_MIN_BLOCKS=$mb AGAR_B200_LIB=$PWD/$L python bench.py --world $w --no-cpu-baseline --e2e-steps 2 --steps $STEPS --no-extras --age 0 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', '$t', round(d['value']/1e6,2), 'M  frac', round(d['roofline']['frac'],3))"
done; done; done
