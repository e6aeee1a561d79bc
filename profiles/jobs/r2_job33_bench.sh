set -x
( time python bench.py ) > gpurun_out/r2_j33_bench_default.json 2> gpurun_out/r2_j33_bench_default.err
tail -c 400 gpurun_out/r2_j33_bench_default.json
timeout 900 python -m pytest tests/test_gpu_chain_statistics.py -m gpu -q -s -k "c4" > gpurun_out/r2_j33_chain_c4.log 2>&1
grep -E "c4_8body|passed|failed" gpurun_out/r2_j33_chain_c4.log
