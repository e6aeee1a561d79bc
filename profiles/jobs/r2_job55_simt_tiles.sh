set -x
timeout 400 python profiles/simt_tile_sweep.py > gpurun_out/r2_j55_simt_tile_sweep.json 2> gpurun_out/r2_j55_simt_tile_sweep.err
cat gpurun_out/r2_j55_simt_tile_sweep.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,row in d.items():
    print(k, ' | '.join(f\"{n}: {v['ms_per_evaluation']:.3f}{'' if v['bit_identical_to_first'] else ' DIFF'}\" for n,v in row.items()))
"
tail -3 gpurun_out/r2_j55_simt_tile_sweep.err
