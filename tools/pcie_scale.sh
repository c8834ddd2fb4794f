#!/bin/bash
# host <-> device copy ceilings with 1 / 2 / 4 / 8 ranks copying at once (gpurun --gpus 8) -> gpurun_out/e2e_pcie_N.json
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 120 python tools/pcie_probe.py 2>/dev/null | tail -1 > gpurun_out/e2e_pcie_$n.json
  else
    timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) tools/pcie_probe.py 2>/dev/null | tail -1 > gpurun_out/e2e_pcie_$n.json
  fi
  cat gpurun_out/e2e_pcie_$n.json
done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -25 > gpurun_out/lscpu.txt 2>&1
