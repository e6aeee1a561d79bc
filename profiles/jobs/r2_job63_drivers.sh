set -x
timeout 300 python -m pytest tests -m gpu -q -k "stale or single_step or driver_cli or integration_stub" > gpurun_out/r2_j63_tests.log 2>&1
tail -4 gpurun_out/r2_j63_tests.log
