#!/bin/bash
# forward kernel warps per CTA at the small training shapes (VAEMDL_FWD_WARPS=<n>; the gradient kernels keep their default)
for rep in 1 2; do
  for wl in "$@"; do
    echo -n "default | "; timeout 60 python tools/step_breakdown.py $wl
    for w in 10 11 12 13; do echo -n "fwd warps $w | "; VAEMDL_FWD_WARPS=$w timeout 60 python tools/step_breakdown.py $wl; done
  done
done
