#!/bin/bash
# n_mix 30: one-pass gradient from the forward pass's sums (default) vs two-pass gradient on tensor memory (VAEMDL_STATS=none), by size
for wl in 5,64,32,32,30 5,128,32,32,30 16,16,64,64,30 16,24,64,64,30 16,32,64,64,30; do
  for env in "VAEMDL_STATS=30" "VAEMDL_STATS=none"; do
    echo -n "$env | "; env $env timeout 90 python tools/step_breakdown.py $wl
  done
done
