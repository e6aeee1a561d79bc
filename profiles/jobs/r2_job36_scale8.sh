set -x
for n in 1 2 4 8; do
  if [ $n = 1 ]; then python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j36_scale_n1.json 2> gpurun_out/r2_j36_scale_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_j36_scale_n$n.json 2> gpurun_out/r2_j36_scale_n$n.err; fi
  tail -c 200 gpurun_out/r2_j36_scale_n$n.json; grep -c "NCCL INFO" gpurun_out/r2_j36_scale_n$n.err
done
