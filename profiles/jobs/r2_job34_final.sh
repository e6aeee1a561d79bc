set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or many_tiles or layer_taps" > gpurun_out/r2_j34_variants.log 2>&1
tail -4 gpurun_out/r2_j34_variants.log
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "test_unet_forward_tcgen05_fp16_vs_golden" > gpurun_out/r2_j34_racecheck_cm.log 2>&1
tail -12 gpurun_out/r2_j34_racecheck_cm.log
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_traffic_v2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/r2_traffic_v2_run.log 2>&1
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2_j34_bench.json 2> gpurun_out/r2_j34_bench.err; tail -c 300 gpurun_out/r2_j34_bench.json
