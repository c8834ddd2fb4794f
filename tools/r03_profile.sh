#!/bin/bash
# round-2 (second session) ncu evidence for the tensor-memory gradient kernel: launch list of the bench command, full captures
# of the headline kernels and of the BASELINE configs[0] kernels.  Outputs under gpurun_out/ (<= 64 MiB per call).
B="python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --no-eval --no-split --no-small --sustain-s 0 --e2e-steps 1"
timeout 300 $B > gpurun_out/r03_prof_plain.json 2> gpurun_out/r03_prof_plain.err || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r03_launches.csv $B > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"modl_tile|finish_warp" -s 9 -c 3 -o gpurun_out/r03_modl_headline -f $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"modl_tile|finish_warp" -s 30 -c 3 -o gpurun_out/r03_modl_cfg1 -f python tools/cfg1_probe.py cfg1 5 > /dev/null 2>&1
ls -la gpurun_out/ | grep r03
