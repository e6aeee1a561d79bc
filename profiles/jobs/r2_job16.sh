set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_j16_tests.log 2>&1
tail -3 gpurun_out/r2_j16_tests.log
for i in 1 2; do
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j16_a20_$i.json 2> gpurun_out/r2_j16_a20_$i.err
CINDM_B200_LIB=$PWD/cindm_b200/lib/libcindm_b200_attn16.so python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j16_a16_$i.json 2> gpurun_out/r2_j16_a16_$i.err
done
python - <<'PY'
import json
for v in ("a20_1","a16_1","a20_2","a16_2"):
    d=json.load(open(f"gpurun_out/r2_j16_{v}.json")); k=d["kernel_classes_one_evaluation"]
    print(v, round(d["value"],3), round(d["ms_per_step"],2), d["clocks"]["sm_mhz"], {c:round(x["ms"],3) for c,x in k.items()})
PY
grep "attn_tc" gpurun_out/r2_j16_a20_1.err; grep "attn_tc" gpurun_out/r2_j16_a16_1.err
