set -x
timeout 100 python -m pytest tests/test_dataset_reader.py -m gpu -q > gpurun_out/r2_j66_tests.log 2>&1
tail -6 gpurun_out/r2_j66_tests.log
