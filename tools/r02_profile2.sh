#!/bin/bash
# instruction counts / durations of the headline kernels (cold, serialised) + full capture of the direct-bf16 kernels
B="python bench.py --steps 4 --warmup 3 --no-also --no-cpu-baseline --no-eval --no-split --no-small --sustain-s 0 --e2e-steps 1"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"modl_tile_kernel|finish_warp" -s 9 -c 3 --csv --log-file gpurun_out/r02t_headline_counts.csv $B > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"512, 0, 2," -s 4 -c 2 -o gpurun_out/r02_modl_bf16 -f python tools/bf16_step_probe.py > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02t_headline_counts.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]: d.setdefault((int(r[ii]),r[ik][:60]),{})[r[im]]=r[iv]
for k in sorted(d): print(k, d[k])
PY
for i in 1 2; do timeout 60 python tools/step_breakdown.py cfg5_64_m10; done
