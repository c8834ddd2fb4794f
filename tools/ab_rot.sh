#!/bin/bash
# interleaved A/B of VAEMDL_ROT for the tiled kernels
for rep in 1 2; do
for wl in cfg5_64_m10 cfg5_64_m20 cfg5_64_m30 cfg1; do
for r in 0 1; do
  echo -n "ROT=$r "; VAEMDL_ROT=$r python tools/step_breakdown.py $wl
done; done; done
