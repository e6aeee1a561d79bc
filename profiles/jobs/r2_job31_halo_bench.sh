set -x
for hm in 0 1 3 0 1 3; do
  CINDM_CONV_CM_HALO=$hm python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j31_bench_h$hm.json 2>> gpurun_out/r2_j31_bench.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r2_j31_bench_h$hm.json')); print('halo $hm', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'], d['clocks']['power_w'], d['kernel_classes_one_evaluation']['conv_tc']['ms'])"
done
