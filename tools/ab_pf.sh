#!/bin/bash
# interleaved A/B: bulk L2 prefetch of a warp's next tile (VAEMDL_L2 "pf=": bit 0 forward, bit 1 backward kernel,
# bit 2 evict_last hint on the prefetch, bit 3 backward: issued at the start of the second pass)
for rep in 1 2; do
  for wl in "$@"; do
    for opt in pf=0 pf=6 pf=10 pf=14 pf=6,hint=1 pf=14,hint=3; do
      echo -n "$opt "; VAEMDL_L2=$opt timeout 60 python tools/step_breakdown.py $wl
    done
  done
done
