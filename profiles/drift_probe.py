"""How far do the 16-bit paths land from the fp32 path after a long chain with identical Philox draws?  Prints the
per-candidate mean final distance to the target for fp32 / fp16 / bf16 and, for scale, fp32 with another seed.

    python profiles/drift_probe.py > gpurun_out/drift.txt
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D, get_design_fn
from cindm_b200.model.params import init_unet_params


def main():
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
    model.load_state_dict(init_unet_params(seed=0))
    dif.to("cuda:0")
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    kw = dict(n_composed=1, compose_start_step=10, compose_n_bodies=4, compose_mode="mean-inside", design_fn=fn,
              design_guidance="standard-recurrence-2")
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dif.num_timesteps = steps

    def dist(x):           # [B]: mean over bodies of the last frame's distance to (0.5, 0.5)
        p = x[:, -1].reshape(x.shape[0], -1, 4)[..., :2].double()
        return (p - 0.5).norm(dim=-1).mean(-1)

    runs = {}
    for name, prec, eng, seed in (("fp32", "fp32", "simt", 31), ("fp16", "fp16", "tcgen05", 31), ("bf16", "bf16", "tcgen05", 31),
                                  ("fp32 seed 32", "fp32", "simt", 32)):
        dif.precision, dif.conv_engine, dif.seed = prec, eng, seed
        model.precision, model.conv_engine = prec, eng
        runs[name] = dif.p_sample_loop((64, 24, 8), None, **kw).cpu()
    ref = runs["fp32"]
    print(f"{steps} DDPM steps, R = 2, 64 candidates, 4 bodies, 2 windows")
    for name, x in runs.items():
        d = dist(x)
        rel = ((x - ref).double().norm() / ref.double().norm()).item()
        print(f"{name:14s} rel-L2 vs fp32 {rel:.3e}   mean final distance {d.mean():.5f} +- {d.std() / 8:.5f} (s.e.)   "
              f"mean |d - d_fp32| {(d - dist(ref)).abs().mean():.5f}   mean (d - d_fp32) {(d - dist(ref)).mean():+.5f}")


if __name__ == "__main__":
    main()
