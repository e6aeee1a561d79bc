set -x
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_run.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:stem_mma_kernel|head_kernel" -s 2 -c 2 -f -o gpurun_out/r2_stemhead_v2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_stemhead_v2.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, \(int\)64, \(int\)8, \(int\)0, \(int\)1, \(int\)2" -s 12 -c 2 -f -o gpurun_out/r2_bias64occ2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_bias64occ2.log 2>&1
( time python bench.py ) > gpurun_out/r2_final_bench_default.json 2> gpurun_out/r2_final_bench_default.err
tail -c 300 gpurun_out/r2_final_bench_default.json
