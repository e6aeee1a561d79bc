set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conditioned.py tests/test_gpu_ebm.py tests/test_gpu_robustness.py tests/test_integration_stub.py -m gpu -x -q > gpurun_out/r2_j24_tests.log 2>&1
tail -25 gpurun_out/r2_j24_tests.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j24_mma.json 2> gpurun_out/r2_j24_mma.err
CINDM_STEM_SIMT=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j24_simt.json 2> gpurun_out/r2_j24_simt.err
python - <<'PY'
import json
for v in ("mma","simt"):
    d=json.load(open(f"gpurun_out/r2_j24_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {c:round(x["ms"],3) for c,x in k.items()})
PY
python __graft_entry__.py --smoke 2>&1 | tail -2
