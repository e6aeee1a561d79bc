set -x
timeout 300 python profiles/other_models_bench.py > gpurun_out/r2_j64_other_models.json 2> gpurun_out/r2_j64_other_models.err
cat gpurun_out/r2_j64_other_models.json | python -c "
import json,sys
for k,v in json.load(sys.stdin).items(): print(k, round(v['ms_per_ddpm_step'],2), round(v['designs_per_s'],3), v['finite'])"
