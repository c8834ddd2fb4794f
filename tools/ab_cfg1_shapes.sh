#!/bin/bash
# BASELINE configs[0] as three launches: warps per CTA of the tiled kernels (VAEMDL_TUNE), one slot per warp
for w in 16 14 12 10 8 6; do
  echo -n "warps $w: "; VAEMDL_FUSED=0 VAEMDL_TUNE="fwd=1:$w,bwd=1:$w" timeout 60 python tools/step_breakdown.py cfg1
done
echo -n "default 3-launch: "; VAEMDL_FUSED=0 timeout 60 python tools/step_breakdown.py cfg1
echo -n "default fused   : "; timeout 60 python tools/cfg1_probe.py cfg1
