#!/bin/bash
# A/B of the tensor-memory gradient kernel's refill point (VAEMDL_TM_REFILL: pair of the first pass; 5 = after the pass)
for rep in 1 2; do
  for wl in "$@"; do
    for env in "VAEMDL_TM=0" "VAEMDL_TM=1 VAEMDL_TM_REFILL=4" "VAEMDL_TM=1 VAEMDL_TM_REFILL=5"; do
      echo -n "$env | "; env $env timeout 90 python tools/step_breakdown.py $wl
    done
  done
done
