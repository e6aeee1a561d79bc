"""Rate of the reference's other model shapes (44-step rollout, Unet_dim 64 / 96) on the generic fp32 CUDA kernels, next to
the 24-step model on the same fp32 kernels and on the 16-bit tensor-core path.  Secondary figures for DESIGN.md.

    python profiles/other_models_bench.py > gpurun_out/other_models.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from cindm_b200 import _lib
from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D, get_design_fn
from cindm_b200.model.params import init_unet_params, unet_param_shapes

# name: (horizon, dim, precision, engine)
MODELS = {
    "24-step dim 64, fp16 tcgen05": (24, 64, "fp16", "tcgen05"),
    "24-step dim 64, fp32 simt": (24, 64, "fp32", "simt"),
    "44-step dim 64, fp32 simt": (44, 64, "fp32", "simt"),
    "44-step dim 96, fp32 simt": (44, 96, "fp32", "simt"),
}
B, N_BODIES, GUIDANCE, R = 50, 2, "standard-recurrence-10", 10


def main():
    L = _lib.lib()
    dev = torch.device("cuda:0")
    st = _lib.stream_ptr(dev)
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    out = {}
    for name, (hor, dim, prec, engine) in MODELS.items():
        model = TemporalUnet1D(horizon=hor, transition_dim=8, cond_dim=False, dim=dim, dim_mults=(1, 2, 4, 8), attention=True)
        dif = GaussianDiffusion1D(model, image_size=hor, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
        model.load_state_dict(init_unet_params(unet_param_shapes(hor, 8, dim), seed=0))
        dif.to(dev)
        dif.precision, dif.conv_engine = prec, engine
        eng = model.engine()
        x = torch.empty(B, hor, 4 * N_BODIES, device=dev)
        _lib.check(L.cindm_fill_initial_noise(_lib.ptr(x), B, hor, N_BODIES, 0, 0, 1000, st))

        def run(t0, k):
            cfg = dif._sample_config(B, 0, 10, N_BODIES, "mean-inside", fn, GUIDANCE, t0, t0 - k + 1, True)
            _lib.check(L.cindm_sample(eng.handle, ctypes.byref(cfg), _lib.ptr(x), None, None, st))

        run(999, 3)
        torch.cuda.synchronize()
        steps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(996, steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[name] = {"candidates": B, "recurrence": R, "ms_per_ddpm_step": ms, "designs_per_s": B / ms,
                     "finite": bool(torch.isfinite(x).all())}
        model._drop_engine()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
