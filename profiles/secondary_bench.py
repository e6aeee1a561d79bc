"""Secondary measurements quoted in DESIGN.md (not the bench.py headline): designs/sec of the smaller
configurations C1-C3 of SURVEY.md section 8 on one GPU, R = 10 and R = 1 for C4, and the C5 scoring rollout.

    python profiles/secondary_bench.py > gpurun_out/secondary.json
"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from cindm_b200 import _lib
from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D, get_design_fn
from cindm_b200.model.params import init_unet_params
from cindm_b200.utils import score_designs, simulation

CONFIGS = {   # name: (B, n_bodies, n_composed, guidance)
    "C1 2-body W=1 B=50 R=10": (50, 2, 0, "standard-recurrence-10"),
    "C2 2-body W=3 B=500 R=10": (500, 2, 2, "standard-recurrence-10"),
    "C3 4-body W=1 B=500 R=10": (500, 4, 0, "standard-recurrence-10"),
    "C4 8-body W=3 B=512 R=10": (512, 8, 2, "standard-recurrence-10"),
    "C4 8-body W=3 B=512 R=1": (512, 8, 2, "standard"),
}


def main():
    L = _lib.lib()
    dev = torch.device("cuda:0")
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
    model.load_state_dict(init_unet_params(seed=0))
    dif.to(dev)
    dif.precision, dif.conv_engine = "fp16", "tcgen05"
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    eng = model.engine()
    st = _lib.stream_ptr(dev)
    out = {}
    for name, (B, n, nc, guidance) in CONFIGS.items():
        T = 24 + nc * 10
        x = torch.empty(B, T, 4 * n, device=dev)
        _lib.check(L.cindm_fill_initial_noise(_lib.ptr(x), B, T, n, 0, 0, 1000, st))
        steps = 40

        def run(t0, k):
            cfg = dif._sample_config(B, nc, 10, n, "mean-inside", fn, guidance, t0, t0 - k + 1, True)
            _lib.check(L.cindm_sample(eng.handle, ctypes.byref(cfg), _lib.ptr(x), None, None, st))

        run(999, 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(995, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"ms_per_ddpm_step": ms, "designs_per_sec": B / ms, "slices_per_evaluation": (nc + 1) * n * (n - 1) // 2 * B}
    # C5: score 1e5 generated-shaped 8-body 44-step designs (rollout of 172 steps + fused MAE / objective)
    rng = np.random.default_rng(0)
    b = 100000
    pred = torch.from_numpy(rng.uniform(0.12, 0.88, size=(b, 44, 32)).astype(np.float32)).to(dev)
    pred[..., 2::4] = torch.from_numpy(rng.uniform(-0.5, 0.5, size=(b, 44, 8)).astype(np.float32)).to(dev)
    pred[..., 3::4] = torch.from_numpy(rng.uniform(-0.5, 0.5, size=(b, 44, 8)).astype(np.float32)).to(dev)
    score_designs(pred[:1000])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mae, obj = score_designs(pred)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["C5 score 1e5 8-body 44-step designs (fused MAE + objective)"] = {"seconds": dt, "designs_per_sec": b / dt,
                                                                        "nan_designs": int(torch.isnan(mae).sum())}
    s0 = (pred[:, 0].reshape(b, 8, 4) * 200.0).double()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    traj = simulation(s0, 172, stride=4, device=dev)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["C5 rollout with trajectory write-out"] = {"seconds": dt, "designs_per_sec": b / dt}
    # CPU oracle for the same rollout on a bounded sample
    from oracle import nbody_ref
    sample = s0[:2000].cpu().numpy()
    t0 = time.perf_counter()
    nbody_ref.rollout(sample, 172, 4)
    dt = time.perf_counter() - t0
    out["C5 CPU oracle (1 thread, 2000 designs)"] = {"seconds": dt, "designs_per_sec": 2000 / dt}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
