set -x
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_j58_gpu_suite.log 2>&1
tail -6 gpurun_out/r2_j58_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j58_smoke.log 2>&1; tail -2 gpurun_out/r2_j58_smoke.log
( time python bench.py ) > gpurun_out/r2_j58_bench_default.json 2> gpurun_out/r2_j58_bench_default.err
tail -c 300 gpurun_out/r2_j58_bench_default.json; tail -4 gpurun_out/r2_j58_bench_default.err
