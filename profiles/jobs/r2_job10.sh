set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:conv_tc_kernel<__half, \(int\)128, \(int\)16, \(int\)1, \(int\)2" -s 7 -c 1 -f -o gpurun_out/r2_gn128pair python profiles/layer_probe.py --evals 2 > gpurun_out/r2_gn128pair.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, \(int\)256, \(int\)32, \(int\)1, \(int\)2" -s 7 -c 1 -f -o gpurun_out/r2_gn256pair python profiles/layer_probe.py --evals 2 > gpurun_out/r2_gn256pair.log 2>&1
$NCU -k "regex:qkv_attn_kernel<__half, \(int\)24" -s 1 -c 1 -f -o gpurun_out/r2_attn24 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_attn24.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, \(int\)64, \(int\)8, \(int\)1, \(int\)1, \(int\)2" -s 6 -c 1 -f -o gpurun_out/r2_gn64occ2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_gn64occ2.log 2>&1
ls -la gpurun_out/*.ncu-rep
