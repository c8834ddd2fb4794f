#!/bin/bash
# n_mix 5 at the config-5 size: pixel-pair kernel vs the 32-row tile kernel with one / two slots per warp
for rep in 1 2; do
echo -n "pp            : "; timeout 60 python tools/step_breakdown.py cfg5_64_m5
echo -n "tile 1 slot ST: "; VAEMDL_PP=0 VAEMDL_M5_SLOTS=1 timeout 60 python tools/step_breakdown.py cfg5_64_m5
echo -n "tile 2 slot ST: "; VAEMDL_PP=0 timeout 60 python tools/step_breakdown.py cfg5_64_m5
echo -n "tile 2 slot 2p: "; VAEMDL_PP=0 VAEMDL_STATS=none timeout 60 python tools/step_breakdown.py cfg5_64_m5
done
echo -n "cfg1_m5 fused  : "; timeout 60 python tools/cfg1_probe.py cfg1_m5
echo -n "cfg1_m5 3 launch 2 slots: "; VAEMDL_FUSED=0 timeout 60 python tools/cfg1_probe.py cfg1_m5
echo -n "cfg1_m5 3 launch 1 slot : "; VAEMDL_FUSED=0 VAEMDL_M5_SLOTS=1 timeout 60 python tools/cfg1_probe.py cfg1_m5
