set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_j32_gpu_suite.log 2>&1
tail -6 gpurun_out/r2_j32_gpu_suite.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_unet_forward_tcgen05_many_tiles or test_composed_eps_tcgen05_fp16" > gpurun_out/r2_j32_memcheck_cm.log 2>&1
tail -8 gpurun_out/r2_j32_memcheck_cm.log
