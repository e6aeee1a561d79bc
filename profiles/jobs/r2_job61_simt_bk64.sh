set -x
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_models.py -m gpu -q -x -k "dispatch_switches or fp32 or simt or other_shapes or philox or graph_replay" > gpurun_out/r2_j61_tests.log 2>&1
tail -5 gpurun_out/r2_j61_tests.log
timeout 400 python profiles/simt_tile_sweep.py > gpurun_out/r2_j61_simt_tile_sweep.json 2> gpurun_out/r2_j61_simt_tile_sweep.err
cat gpurun_out/r2_j61_simt_tile_sweep.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,row in d.items():
    print(k, ' | '.join(f\"{n}: {v['ms_per_evaluation']:.3f}{'' if v['bit_identical_to_first'] else ' DIFF'}\" for n,v in row.items()))
"
tail -3 gpurun_out/r2_j61_simt_tile_sweep.err
