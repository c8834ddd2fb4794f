#!/bin/bash
# round-2 ncu evidence (one GPU), part $1: 1 = launch list of the bench command + full capture of the headline kernels,
# 2 = full captures of the bf16 kernels and the configs[0] kernels.  Outputs under gpurun_out/ (<= 64 MiB per call).
B="python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --no-eval --no-split --no-small --sustain-s 0 --e2e-steps 1"
if [ "$1" = "1" ]; then
  timeout 300 $B > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv $B > /dev/null 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"modl_tile_kernel|finish_warp" -s 9 -c 3 -o gpurun_out/r02_modl_headline -f $B > /dev/null 2>&1
else
  timeout 300 ncu --set full --clock-control none -k regex:modl_tile_kernel -s 12 -c 2 -o gpurun_out/r02_modl_bf16 -f python tools/bf16_step_probe.py > /dev/null 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:"modl_tile_kernel|finish_warp" -s 30 -c 3 -o gpurun_out/r02_modl_cfg1 -f python tools/cfg1_probe.py cfg1 5 > /dev/null 2>&1
fi
ls -la gpurun_out/
