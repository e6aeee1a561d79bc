set -x
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_j48_gpu_suite.log 2>&1
tail -4 gpurun_out/r2_j48_gpu_suite.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_j48_smoke.log 2>&1; tail -2 gpurun_out/r2_j48_smoke.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic_v4.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_v4_run.log 2>&1
( time python bench.py ) > gpurun_out/r2_j48_bench_default.json 2> gpurun_out/r2_j48_bench_default.err
tail -c 200 gpurun_out/r2_j48_bench_default.json
