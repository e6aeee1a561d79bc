set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_identical or test_unet_forward_fp32" > gpurun_out/r2_j38_simt_parity.log 2>&1
tail -3 gpurun_out/r2_j38_simt_parity.log
python bench.py --precision fp32 --engine simt --steps 2 --warmup 1 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j38_fp32.json 2> gpurun_out/r2_j38_fp32.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_j38_fp32.json').read().strip().splitlines()[-1]); print('tile128', d['ms_per_step'], d['value'], d.get('model_tflops_per_gpu'), d['clocks'])
print({k:(v['launches'], round(v['ms'],2)) for k,v in d.get('kernel_classes_one_evaluation',{}).items()})"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv1d_simt128_kernel<float, float, \(int\)128" -s 30 -c 1 -f -o gpurun_out/r2_simt128 python profiles/layer_probe.py --evals 1 --precision fp32 --engine simt --candidates 64 > gpurun_out/r2_simt128.log 2>&1
tail -2 gpurun_out/r2_simt128.log
