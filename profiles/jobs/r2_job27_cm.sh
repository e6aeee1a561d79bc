set -x
for m in 0 1 2 4; do
  CINDM_CONV_CM=$m timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j27_prof_m$m.txt 2>&1
  grep -E "candidates| gn" gpurun_out/r2_j27_prof_m$m.txt
done
CINDM_CONV_CM=7 CINDM_CONV_CM_EW=16 timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j27_prof_m7_ew16.txt 2>&1
grep -E "candidates| gn" gpurun_out/r2_j27_prof_m7_ew16.txt
for m in 1 2 4; do
  CINDM_CONV_CM=$m timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tcgen05" > gpurun_out/r2_j27_parity_m$m.log 2>&1
  tail -5 gpurun_out/r2_j27_parity_m$m.log
done
