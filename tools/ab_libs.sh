#!/bin/bash
# interleaved A/B of several builds of the library in one gpurun call: vae_mdl_b200/libvaemdl_b200_<tag>.so vs the in-tree build
for rep in 1 2; do
  for wl in "$@"; do
    echo -n "base | "; timeout 60 python tools/step_breakdown.py $wl
    for so in vae_mdl_b200/libvaemdl_b200_*.so; do
      echo -n "$(basename $so .so | sed s/libvaemdl_b200_//) | "; VAEMDL_LIB_PATH=$PWD/$so timeout 60 python tools/step_breakdown.py $wl
    done
  done
done
