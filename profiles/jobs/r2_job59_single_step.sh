set -x
timeout 600 python -m pytest tests/test_gpu_models.py tests/test_gpu_conditioned.py -m gpu -q > gpurun_out/r2_j59_tests.log 2>&1
tail -15 gpurun_out/r2_j59_tests.log
