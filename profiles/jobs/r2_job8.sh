set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2_j8_bench2.json 2> gpurun_out/r2_j8_bench2.err
wc -l gpurun_out/r2_j8_bench2.json; head -c 200 gpurun_out/r2_j8_bench2.json
grep -c "NCCL INFO" gpurun_out/r2_j8_bench2.err; grep -i "nranks" gpurun_out/r2_j8_bench2.err | head -3
python -m pytest tests/test_gpu_robustness.py -m gpu -x -q > gpurun_out/r2_j8_robust.log 2>&1
tail -5 gpurun_out/r2_j8_robust.log
