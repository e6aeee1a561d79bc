set -x
python -m pytest tests/test_gpu_parity.py tests/test_integration_stub.py -m gpu -x -q > gpurun_out/r2_j3_tests.log 2>&1
tail -3 gpurun_out/r2_j3_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r2_j3_bench.json 2> gpurun_out/r2_j3_bench.err
head -12 gpurun_out/r2_j3_bench.err
