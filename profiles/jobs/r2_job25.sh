set -x
python -m pytest tests/test_gpu_conditioned.py tests/test_gpu_ebm.py -m gpu -x -q > gpurun_out/r2_j25_tests.log 2>&1
tail -25 gpurun_out/r2_j25_tests.log
