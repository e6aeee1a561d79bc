set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_run.log 2>&1
( time python bench.py ) > gpurun_out/r2_j12_bench_default.json 2> gpurun_out/r2_j12_bench_default.err
tail -c 600 gpurun_out/r2_j12_bench_default.json; tail -5 gpurun_out/r2_j12_bench_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2_j12_bench_ref.json 2> gpurun_out/r2_j12_bench_ref.err
tail -c 400 gpurun_out/r2_j12_bench_ref.json; tail -4 gpurun_out/r2_j12_bench_ref.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_ebm.py tests/test_gpu_conditioned.py tests/test_gpu_robustness.py -m gpu -x -q -k "fp16 or robust or overflow or survives" > gpurun_out/r2_memcheck_new_paths.txt 2>&1
tail -8 gpurun_out/r2_memcheck_new_paths.txt
