set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bit_identical or test_unet_forward_fp32 or layer_taps_fp32 or composed_eps_fp32" > gpurun_out/r2_j37_simt_parity.log 2>&1
tail -4 gpurun_out/r2_j37_simt_parity.log
for f in 0 1; do
  CINDM_SIMT_TILE128=$f python bench.py --precision fp32 --engine simt --steps 2 --warmup 1 --no-cpu-baseline --no-extra --profile > gpurun_out/r2_j37_fp32_tile$f.json 2> gpurun_out/r2_j37_fp32_tile$f.err
  python -c "
import json
d=json.loads(open('gpurun_out/r2_j37_fp32_tile$f.json').read().strip().splitlines()[-1]); print('tile128=$f', d['ms_per_step'], d['value'], d.get('model_tflops_per_gpu'), d['clocks'])
print({k:(v['launches'], round(v['ms'],2)) for k,v in d.get('kernel_classes_one_evaluation',{}).items()})"
done
