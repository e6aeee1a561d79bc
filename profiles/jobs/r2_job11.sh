set -x
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j11_after.json 2> gpurun_out/r2_j11_after.err
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_j11_tests.log 2>&1
tail -3 gpurun_out/r2_j11_tests.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j11_after2.json 2> gpurun_out/r2_j11_after2.err
grep -E "^# conv_tc|128->128 k5|k1 bias" gpurun_out/r2_j11_after.err | head -20
