set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_j17_tests.log 2>&1
tail -3 gpurun_out/r2_j17_tests.log
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j17_$i.json 2> gpurun_out/r2_j17_$i.err
done
python - <<'PY'
import json
for v in ("1","2"):
    d=json.load(open(f"gpurun_out/r2_j17_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {c:round(x["ms"],3) for c,x in k.items()})
PY
grep -E "64->64 k5|128->128 k5|64->128 k5|128->64 k5" gpurun_out/r2_j17_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_j17_bench2.json 2> gpurun_out/r2_j17_bench2.err
wc -l gpurun_out/r2_j17_bench2.json; grep -c "NCCL INFO" gpurun_out/r2_j17_bench2.err; grep -i "nranks" gpurun_out/r2_j17_bench2.err | head -4
