set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 and not variants" > gpurun_out/r2_j45_parity.log 2>&1
tail -3 gpurun_out/r2_j45_parity.log
for v in 0 1 0 1; do
  CINDM_L2_STREAM=$v timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j45_prof_l$v.txt 2>&1
  grep -E "candidates" gpurun_out/r2_j45_prof_l$v.txt
done
for cfg in "0 0" "1 0" "1 1"; do
set -- $cfg
CINDM_SNAKE=$1 CINDM_L2_STREAM=$2 ncu --cache-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic_nc_s$1_l$2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_nc_run.log 2>&1
done
for v in 0 1 0 1; do
CINDM_L2_STREAM=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j45_bench_l$v.json 2>> gpurun_out/r2_j45_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_j45_bench_l$v.json')); print('l2stream $v', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['power_w'], d['kernel_classes_one_evaluation']['conv_tc']['ms'], d['kernel_classes_one_evaluation']['attn_tc']['ms'])"
done
