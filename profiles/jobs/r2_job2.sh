set -x
python -m pytest tests/test_gpu_chain_statistics.py -m gpu -x -q -s > gpurun_out/r2_chain_test.log 2>&1
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:conv_tc_kernel<__half, \(int\)64, \(int\)8, \(int\)1" -s 6 -c 1 -f -o gpurun_out/r2_base_gn64 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_gn64.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, \(int\)128, \(int\)16, \(int\)1" -s 6 -c 1 -f -o gpurun_out/r2_base_gn128 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_gn128.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, \(int\)128, \(int\)8, \(int\)0" -s 20 -c 2 -f -o gpurun_out/r2_base_bias128 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_bias128.log 2>&1
tail -3 gpurun_out/r2_chain_test.log
