set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not stale_script" > gpurun_out/r2_j4_tests.log 2>&1
tail -3 gpurun_out/r2_j4_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j4_bench_occ2.json 2> gpurun_out/r2_j4_bench_occ2.err
CINDM_CONV_OCC2=0 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j4_bench_occ1.json 2> gpurun_out/r2_j4_bench_occ1.err
head -6 gpurun_out/r2_j4_bench_occ2.err gpurun_out/r2_j4_bench_occ1.err
