set -x
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j13_hint_$i.json 2> gpurun_out/r2_j13_hint_$i.err
CINDM_B200_LIB=$PWD/cindm_b200/lib/libcindm_b200_spin.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j13_spin_$i.json 2> gpurun_out/r2_j13_spin_$i.err
done
python - <<'PY'
import json
for v in ("hint_1","spin_1","hint_2","spin_2"):
    d=json.load(open(f"gpurun_out/r2_j13_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], d["clocks"]["power_w"], {c:round(x["ms"],3) for c,x in k.items()})
PY
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2_j13_gpu_suite.log 2>&1
tail -5 gpurun_out/r2_j13_gpu_suite.log
