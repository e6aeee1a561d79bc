set -x
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j20_n128_$i.json 2> gpurun_out/r2_j20_n128_$i.err
CINDM_CONV_BIAS64=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j20_n64_$i.json 2> gpurun_out/r2_j20_n64_$i.err
done
python - <<'PY'
import json
for v in ("n128_1","n64_1","n128_2","n64_2"):
    d=json.load(open(f"gpurun_out/r2_j20_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {c:round(x["ms"],3) for c,x in k.items()})
PY
grep -E "bias" gpurun_out/r2_j20_n128_1.err | sort > /tmp/a.txt; grep -E "bias" gpurun_out/r2_j20_n64_1.err | sort > /tmp/b.txt; paste -d'|' /tmp/a.txt /tmp/b.txt | cut -c1-220
CINDM_CONV_BIAS64=1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tcgen05 or driver_cli_batches" > gpurun_out/r2_j20_tests.log 2>&1
tail -3 gpurun_out/r2_j20_tests.log
