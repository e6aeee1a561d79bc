set -x
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv1d_simt128_kernel<float, float, \(int\)128" -s 30 -c 1 -f -o gpurun_out/r2_simt128_posmajor python profiles/layer_probe.py --evals 1 --precision fp32 --engine simt --candidates 64 > gpurun_out/r2_simt128_posmajor.log 2>&1
tail -2 gpurun_out/r2_simt128_posmajor.log
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv1d_simt_kernel<float, float, \(int\)32" -s 30 -c 1 -f -o gpurun_out/r2_simt32 python profiles/layer_probe.py --evals 1 --precision fp32 --engine simt --candidates 1 > gpurun_out/r2_simt32.log 2>&1
tail -2 gpurun_out/r2_simt32.log
ls -la gpurun_out/*.ncu-rep | tail -3
