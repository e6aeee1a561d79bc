set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_ebm.py -m gpu -x -q > gpurun_out/r2_j21_tests.log 2>&1
tail -3 gpurun_out/r2_j21_tests.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j21_a.json 2> gpurun_out/r2_j21_a.err
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile --full-sample > gpurun_out/r2_j21_full.json 2> gpurun_out/r2_j21_full.err
python - <<'PY'
import json
for v in ("a","full"):
    d=json.load(open(f"gpurun_out/r2_j21_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {c:round(x["ms"],3) for c,x in k.items()}, d.get("full_sample_api"))
PY
