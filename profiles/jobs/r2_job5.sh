set -x
python -m pytest tests/test_gpu_conditioned.py -m gpu -x -q > gpurun_out/r2_j5_cond.log 2>&1
tail -25 gpurun_out/r2_j5_cond.log
python -m pytest tests/test_gpu_chain_statistics.py -m gpu -q -s -k "4body" > gpurun_out/r2_j5_chain4.log 2>&1
tail -5 gpurun_out/r2_j5_chain4.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_nbody.py -m gpu -q -k "not stale_script" > gpurun_out/r2_j5_parity.log 2>&1
tail -3 gpurun_out/r2_j5_parity.log
