#!/bin/bash
# two-slot gradient kernels (n_mix 5 float32, bfloat16 tiles): where in a tile the other slot is refilled (VAEMDL_REFILL2=<pair>)
for rep in 1 2; do
  for wl in cfg5_64_m5 cfg1_m5; do
    for r in 0 1 2; do echo -n "REFILL2=$r | "; VAEMDL_REFILL2=$r timeout 90 python tools/step_breakdown.py $wl; done
  done
  for r in 0 2 3 4; do echo "REFILL2=$r | bf16:"; VAEMDL_REFILL2=$r timeout 120 python tools/bf16_step_probe.py 2>&1 | tail -4; done
done
