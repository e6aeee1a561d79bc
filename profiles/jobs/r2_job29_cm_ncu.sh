set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)12, \(int\)128, \(int\)16, \(bool\)1" -s 2 -c 1 -f -o gpurun_out/r2_cm128res python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm128res.log 2>&1
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)24, \(int\)64, \(int\)16, \(bool\)0" -s 2 -c 1 -f -o gpurun_out/r2_cm64 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm64.log 2>&1
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)6, \(int\)256, \(int\)16, \(bool\)1" -s 2 -c 1 -f -o gpurun_out/r2_cm256res python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm256res.log 2>&1
tail -3 gpurun_out/r2_cm128res.log gpurun_out/r2_cm64.log gpurun_out/r2_cm256res.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j29_bench.json 2> gpurun_out/r2_j29_bench.err; tail -c 1200 gpurun_out/r2_j29_bench.json
