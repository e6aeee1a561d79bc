set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench8.json 2> gpurun_out/r2_bench8.err
wc -l gpurun_out/r2_bench8.json; grep -c "NCCL INFO" gpurun_out/r2_bench8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
python - <<'PY'
import json
for n in (8,4):
    d=json.load(open(f"gpurun_out/r2_bench{n}.json"))
    print(n, d["value"], d["ms_per_step"], d["collective_ms"], d["scoring_ms"], d["e2e"]["value"], d["clocks"])
PY
