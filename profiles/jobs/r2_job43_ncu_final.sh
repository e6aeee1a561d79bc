set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)12, \(int\)128, \(int\)16, \(bool\)1" -s 2 -c 1 -f -o gpurun_out/r2_cm128res_v2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm128res_v2.log 2>&1
$NCU -k "regex:conv_tc_cm_halo_kernel<__half, \(int\)24, \(int\)64, \(int\)12, \(bool\)1" -s 2 -c 1 -f -o gpurun_out/r2_cmhalo64res python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cmhalo64res.log 2>&1
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)24, \(int\)64, \(int\)16, \(bool\)0" -s 2 -c 1 -f -o gpurun_out/r2_cm64_v2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm64_v2.log 2>&1
$NCU -k "regex:conv_tc_cm_kernel<__half, \(int\)6, \(int\)256, \(int\)16, \(bool\)1" -s 2 -c 1 -f -o gpurun_out/r2_cm256res_v2 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_cm256res_v2.log 2>&1
ls -la gpurun_out/*_v2.ncu-rep gpurun_out/r2_cmhalo64res.ncu-rep
