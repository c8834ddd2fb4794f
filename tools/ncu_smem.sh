#!/bin/bash
# shared-memory wavefronts / bank conflicts of the tiled kernels with and without the pair rotation (ncu, one GPU)
M=${1:-10}
METRICS=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,smsp__inst_executed_op_shared_ld.sum,smsp__inst_executed_op_shared_st.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__cycles_elapsed.max
for r in 0 1; do
  echo "== VAEMDL_ROT=$r M=$M"
  VAEMDL_ROT=$r MS=$M ncu --metrics $METRICS --clock-control none -k regex:modl_tile -c 6 --csv --log-file gpurun_out/ncu_smem_rot${r}_m${M}.csv python tools/prof_m.py > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_smem_rot${r}_m${M}.csv")) if len(r)>10]
hdr=rows[0]; i_k=hdr.index("Kernel Name"); i_m=hdr.index("Metric Name"); i_v=hdr.index("Metric Value"); i_id=hdr.index("ID")
seen={}
for r in rows[1:]:
    seen.setdefault((r[i_id], r[i_k][:60]),{})[r[i_m]]=r[i_v]
for (i,k),m in seen.items():
    if i in ("2","5"): print(k, {a.replace("l1tex__data_","").replace("_mem_shared","").replace("pipe_lsu_",""):b for a,b in m.items()})
PY
done
