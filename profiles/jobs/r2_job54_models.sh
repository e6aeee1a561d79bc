set -x
timeout 600 python -m pytest tests/test_gpu_models.py -m gpu -q > gpurun_out/r2_j54_models.log 2>&1
tail -25 gpurun_out/r2_j54_models.log
timeout 300 python profiles/other_models_bench.py > gpurun_out/r2_j54_other_models.json 2> gpurun_out/r2_j54_other_models.err
cat gpurun_out/r2_j54_other_models.json; tail -3 gpurun_out/r2_j54_other_models.err
