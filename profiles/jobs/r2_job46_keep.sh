set -x
for v in 0 1 0 1; do
  CINDM_L2_KEEP=$v timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j46_prof_k$v.txt 2>&1
  grep -E "candidates" gpurun_out/r2_j46_prof_k$v.txt
done
CINDM_L2_KEEP=1 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic_nc_keep1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_nc_run.log 2>&1
for v in 0 1 0 1; do
CINDM_L2_KEEP=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j46_bench_k$v.json 2>> gpurun_out/r2_j46_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_j46_bench_k$v.json')); print('keep $v', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['power_w'], d['kernel_classes_one_evaluation']['conv_tc']['ms'], d['kernel_classes_one_evaluation']['attn_tc']['ms'])"
done
