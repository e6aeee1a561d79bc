set -x
python -m pytest tests/test_gpu_robustness.py -m gpu -x -q > gpurun_out/r2_j7_robust.log 2>&1
tail -30 gpurun_out/r2_j7_robust.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_j7_bench2.json 2> gpurun_out/r2_j7_bench2.err
tail -c 1500 gpurun_out/r2_j7_bench2.json
grep -c "NCCL INFO" gpurun_out/r2_j7_bench2.err; grep -i "nranks" gpurun_out/r2_j7_bench2.err | head -5
tail -5 gpurun_out/r2_j7_bench2.err
