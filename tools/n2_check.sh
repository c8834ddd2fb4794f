B="bench.py --gpus 2 --steps 50 --warmup 5 --no-also --no-eval --no-split --no-small --sustain-s 0 --e2e-steps 1 --no-cpu-baseline"
for env in "VAEMDL_TM=1" "VAEMDL_TM=1" "VAEMDL_TM=0" "VAEMDL_TM=1 VAEMDL_BENCH_NCCL=1"; do
  env $env timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 $B 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$env', d['value'], d['ms_per_step'], d['kernel_ms'])"
done
