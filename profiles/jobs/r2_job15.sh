set -x
python -m pytest tests/test_gpu_nbody.py -m gpu -x -q > gpurun_out/r2_j15_nbody.log 2>&1
tail -15 gpurun_out/r2_j15_nbody.log
python - <<'PY' > gpurun_out/r2_j15_c5.txt 2>&1
import torch, time, sys
sys.path.insert(0, ".")
from cindm_b200.utils import score_designs, simulation
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
b = 100000
d = torch.rand(b, 44, 32, device=dev, generator=g) * 0.76 + 0.12
d[..., 2::4] -= 0.5; d[..., 3::4] -= 0.5
score_designs(d[:1000]); torch.cuda.synchronize()
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mae, obj = score_designs(d); e1.record(); torch.cuda.synchronize()
    print("score 1e5:", e0.elapsed_time(e1), "ms", float(mae.mean()), float(obj.mean()))
s0 = (d[:, 0].reshape(b, 8, 4) * 200.0).double()
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr = simulation(s0, 172, stride=4, device=dev); e1.record(); torch.cuda.synchronize()
    print("rollout 1e5 with trajectory:", e0.elapsed_time(e1), "ms")
PY
cat gpurun_out/r2_j15_c5.txt
ncu --set full --clock-control none --import-source on -k regex:score_designs -c 1 -f -o gpurun_out/r2_score_v3 python profiles/layer_probe.py --evals 0 --score 100000 > gpurun_out/r2_score_v3.log 2>&1
