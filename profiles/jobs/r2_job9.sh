set -x
env | grep -i nccl
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j9_bench_pair128.json 2> gpurun_out/r2_j9_bench_pair128.err
CINDM_CONV_PAIR128=0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j9_bench_nopair128.json 2> gpurun_out/r2_j9_bench_nopair128.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_j9_tests.log 2>&1
tail -3 gpurun_out/r2_j9_tests.log
python profiles/small_batch_profile.py 50 2 0 > gpurun_out/r2_j9_c1_profile.txt 2>&1
head -50 gpurun_out/r2_j9_c1_profile.txt
