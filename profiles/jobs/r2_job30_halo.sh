set -x
for hm in 1 2; do
  CINDM_CONV_CM_HALO=$hm timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05" > gpurun_out/r2_j30_parity_h$hm.log 2>&1
  tail -4 gpurun_out/r2_j30_parity_h$hm.log
done
for hm in 0 1 2; do
  CINDM_CONV_CM_HALO=$hm timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j30_prof_h$hm.txt 2>&1
  grep -E "candidates| gn" gpurun_out/r2_j30_prof_h$hm.txt
done
