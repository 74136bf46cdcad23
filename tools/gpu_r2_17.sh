#!/bin/bash
for abl in 0 1 2 3; do for shape in "5 44 80" "1 176 320"; do
  echo -n "wgrad ablate $abl $shape (pool policy): "
  DVSR_WG_ABLATE=$abl timeout 120 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:conv_wgrad_tc -s 2 -c 2 python tools/one_wgrad.py $shape --pool-policy 2>&1 | grep -E "gpu__time|grid_size" | awk '{printf "%s ", $NF} END{print ""}'
done; done
