set -x
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$NCU -k "regex:conv_tc_kernel<__half, 64, 8, 1" -s 6 -c 1 -f -o gpurun_out/r2_base_gn64 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_gn64.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, 128, 16, 1" -s 6 -c 1 -f -o gpurun_out/r2_base_gn128 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_gn128.log 2>&1
$NCU -k "regex:conv_tc_kernel<__half, 128, 8, 0" -s 20 -c 2 -f -o gpurun_out/r2_base_bias128 python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_bias128.log 2>&1
$NCU -k "regex:score_designs" -c 1 -f -o gpurun_out/r2_base_score python profiles/layer_probe.py --evals 0 --score 100000 > gpurun_out/r2_base_score.log 2>&1
$NCU -k "regex:stem_kernel|head_kernel" -s 2 -c 2 -f -o gpurun_out/r2_base_stemhead python profiles/layer_probe.py --evals 2 > gpurun_out/r2_base_stemhead.log 2>&1
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --profile > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err
ls -la gpurun_out
