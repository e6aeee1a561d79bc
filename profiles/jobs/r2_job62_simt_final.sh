set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py tests/test_gpu_ebm.py -m gpu -q -x -k "dispatch_switches or fp32 or simt or other_shapes or philox or graph_replay or ebm or single_step" > gpurun_out/r2_j62_tests.log 2>&1
tail -4 gpurun_out/r2_j62_tests.log
timeout 400 python profiles/simt_tile_sweep.py > gpurun_out/r2_j62_simt_tile_sweep.json 2> gpurun_out/r2_j62_simt_tile_sweep.err
cat gpurun_out/r2_j62_simt_tile_sweep.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,row in d.items():
    print(k, ' | '.join(f\"{n}: {v['ms_per_evaluation']:.3f}{'' if v['bit_identical_to_first'] else ' DIFF'}\" for n,v in row.items()))
"
