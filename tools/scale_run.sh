for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-also --no-cpu-baseline > gpurun_out/scale_r01d_$n.json 2> gpurun_out/scale_d_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 100 --warmup 5 --no-also --no-cpu-baseline > gpurun_out/scale_r01d_$n.json 2> gpurun_out/scale_d_$n.err
  fi
  tail -c 200 gpurun_out/scale_r01d_$n.json | head -c 200; echo
done
