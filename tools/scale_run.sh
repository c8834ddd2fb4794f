#!/bin/bash
# bench.py at 1 / 2 / 4 / 8 GPUs in one `gpurun --gpus 8` call (+ the largest config-5 point at N = 8) -> gpurun_out/scale_r02_*.json
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-also > gpurun_out/scale_r02_$n.json 2> gpurun_out/scale_r02_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_r02_$n.json 2> gpurun_out/scale_r02_$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_r02_$n.json").read())
    print("N=$n value %.4g ms %.4f sustained %.4g split %.4g cfg0 %.1f us cfg0' %.1f us e2e %.4g (%.1f ms) eval %.0f img/s  link: %s" % (d["value"], d["ms_per_step"], d.get("value_sustained",0), d["sample_split"]["value"], d["configs0"]["us_per_step"], d["configs0_ref_default"]["us_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["iwae_eval_5000is"]["images_per_s"], str(d.get("collective_in_step"))[:40]))
except Exception as e:
    print("N=$n failed:", e)
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 5 --workload cfg5_128_m30 --no-eval --no-small --no-split --sustain-s 0 --e2e-steps 1 > gpurun_out/scale_r02_8_cfg5_128_m30.json 2> gpurun_out/scale_r02_8_cfg5_128_m30.err
python -c "
import json; d=json.loads(open('gpurun_out/scale_r02_8_cfg5_128_m30.json').read()); print('cfg5_128_m30 N=8: value %.4g ms %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['step']['frac']))"
