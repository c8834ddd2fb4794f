for cfg in cfg1 cfg1_m5; do
for e in "" "keep=32" "keep=64" "keep=128" "hint=1" "hint=2" "hint=3" "keep=64,hint=3" "keep=128,hint=3" "keep=128,hint=2"; do
  VAEMDL_L2S="$e" python tools/cfg1_probe.py $cfg
done; done
