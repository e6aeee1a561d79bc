set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_j51_gpu_suite.log 2>&1
tail -4 gpurun_out/r2_j51_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j51_smoke.log 2>&1; tail -2 gpurun_out/r2_j51_smoke.log
