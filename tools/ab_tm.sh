#!/bin/bash
# interleaved A/B: two-pass gradient kernel with its tile in tensor memory (VAEMDL_TM=1, csrc/modl_tm.cuh) vs in shared memory;
# VAEMDL_TM_WARPS=16|14 (128 / 144 registers), VAEMDL_TM_REFILL=<pair of the first pass before which the slot is refilled>
for rep in 1 2; do
  for wl in "$@"; do
    for env in "VAEMDL_TM=0" "VAEMDL_TM=1 VAEMDL_TM_WARPS=16" "VAEMDL_TM=1 VAEMDL_TM_REFILL=0" "VAEMDL_TM=1 VAEMDL_TM_REFILL=1" "VAEMDL_TM=1 VAEMDL_TM_REFILL=2" "VAEMDL_TM=1 VAEMDL_TM_REFILL=3"; do
      echo -n "$env | "; env $env timeout 90 python tools/step_breakdown.py $wl
    done
  done
done
