set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_conditioned.py tests/test_gpu_ebm.py -m gpu -x -q > gpurun_out/r2_j23_tests.log 2>&1
tail -5 gpurun_out/r2_j23_tests.log
python - <<'PY' > gpurun_out/r2_j23_small.txt 2>&1
import os, sys, subprocess, json
for env in ("1", "0", "1", "0"):
    e = dict(os.environ, CINDM_FORK_RES=env)
    out = subprocess.run([sys.executable, "profiles/secondary_bench.py"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout)
        print("FORK_RES=" + env, {k.split(" ")[0] + ("R1" if "R=1" in k and "R=10" not in k else ""): round(v["ms_per_ddpm_step"], 3) for k, v in d.items() if "ms_per_ddpm_step" in v})
    except Exception as ex:
        print("FORK_RES=" + env, "failed", ex, out.stderr[-800:])
PY
cat gpurun_out/r2_j23_small.txt
