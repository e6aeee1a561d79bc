set -x
CINDM_CONV_PAIR=0 CINDM_CONV_PAIR128=0 timeout 1500 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke > gpurun_out/r2_racecheck_single_cta.txt 2>&1
tail -5 gpurun_out/r2_racecheck_single_cta.txt
timeout 1500 compute-sanitizer --tool racecheck python __graft_entry__.py --smoke > gpurun_out/r2_racecheck_default.txt 2>&1
tail -5 gpurun_out/r2_racecheck_default.txt
grep -c "Race reported" gpurun_out/r2_racecheck_default.txt; grep "Race reported" gpurun_out/r2_racecheck_default.txt | sed -E 's/.*(conv_tc_kernel<[^>]*>|qkv_attn_kernel<[^>]*>|[a-z_]+_kernel).*/\1/' | sort | uniq -c
