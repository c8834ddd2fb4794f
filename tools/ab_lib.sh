#!/bin/bash
# interleaved A/B of two builds of the library in one gpurun call: vae_mdl_b200/libvaemdl_b200_alt.so (A) vs the in-tree build (B)
ALT=$PWD/vae_mdl_b200/libvaemdl_b200_alt.so
for rep in 1 2 3; do
  for wl in "$@"; do
    echo -n "A "; VAEMDL_LIB_PATH=$ALT timeout 60 python tools/step_breakdown.py $wl
    echo -n "B "; timeout 60 python tools/step_breakdown.py $wl
  done
done
