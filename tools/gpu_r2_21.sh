#!/bin/bash
# staged vs direct DCN forward at the inner-loop sizes (whole GPU and the pool's 37-CTA budget)
for shape in "5 44 80" "5 22 40" "5 88 160"; do for b in 0 37; do for st in 1 2; do
  echo -n "dcn $shape budget $b staged=$st: "; timeout 60 python tools/one_dcn.py $shape --staged $st --cta-budget $b 2>&1 | tail -1
done; done; done
