set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 or graph or sharding or candidate_independence" > gpurun_out/r2_j47_parity.log 2>&1
tail -3 gpurun_out/r2_j47_parity.log
for v in 0 1; do
  CINDM_SNAKE=$v timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j47_prof_s$v.txt 2>&1
  grep -E "candidates" gpurun_out/r2_j47_prof_s$v.txt
done
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j47_bench.json 2> gpurun_out/r2_j47_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_j47_bench.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['isolated']['frac'], d['clocks']['sm_mhz'], d['clocks']['power_w'], d['kernel_classes_one_evaluation']['conv_tc']['ms'], d['kernel_classes_one_evaluation']['attn_tc']['ms'])"
