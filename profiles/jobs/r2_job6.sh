set -x
python -m pytest tests/test_gpu_ebm.py -m gpu -x -q > gpurun_out/r2_j6_ebm.log 2>&1
tail -30 gpurun_out/r2_j6_ebm.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conditioned.py tests/test_integration_stub.py -m gpu -q > gpurun_out/r2_j6_parity.log 2>&1
tail -15 gpurun_out/r2_j6_parity.log
python -m pytest tests/test_gpu_chain_statistics.py -m gpu -q -s -k "c1_2body_std" > gpurun_out/r2_j6_chain.log 2>&1
tail -5 gpurun_out/r2_j6_chain.log
