set -x
for ew in 8 12 16; do
  CINDM_CONV_CM=7 CINDM_CONV_CM_EW=$ew timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j28_prof_ew$ew.txt 2>&1
  grep -E "candidates| gn" gpurun_out/r2_j28_prof_ew$ew.txt
done
CINDM_CONV_CM=0 timeout 300 python profiles/small_batch_profile.py 512 8 2 > gpurun_out/r2_j28_prof_old.txt 2>&1
CINDM_CONV_CM=7 CINDM_CONV_CM_EW=12 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tcgen05" > gpurun_out/r2_j28_parity.log 2>&1
tail -5 gpurun_out/r2_j28_parity.log
