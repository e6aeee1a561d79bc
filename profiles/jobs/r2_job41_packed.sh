set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05" > gpurun_out/r2_j41_parity.log 2>&1
tail -3 gpurun_out/r2_j41_parity.log
timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j41_prof.txt 2>&1
grep -E "candidates| gn" gpurun_out/r2_j41_prof.txt
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j41_bench.json 2> gpurun_out/r2_j41_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2_j41_bench.json')); print('bench', d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['isolated']['frac'], d['clocks'], d['kernel_classes_one_evaluation']['conv_tc']['ms'])"
