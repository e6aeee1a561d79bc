set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05 and not variants" > gpurun_out/r2_j53_parity.log 2>&1
tail -3 gpurun_out/r2_j53_parity.log
for i in 1 2; do
timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j53_prof_$i.txt 2>&1
grep -E "candidates|head|stem" gpurun_out/r2_j53_prof_$i.txt
done
