"""Small driver for ncu captures of single kernels: N un-graphed composed-epsilon evaluations of the C4 shape
(43 008 slices) and/or one fused scoring call over 1e5 designs.

    ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
        -k "regex:conv_tc_kernel<__half, 64, 8, 1" -s 6 -c 1 -o gpurun_out/gn64 python profiles/layer_probe.py --evals 2
    ncu --set full --clock-control none --import-source on -k regex:score_designs -c 1 -o gpurun_out/score \
        python profiles/layer_probe.py --evals 0 --score 100000
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from cindm_b200 import _lib
from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
from cindm_b200.model.params import init_unet_params
from cindm_b200.utils import score_designs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--evals", type=int, default=2)
    ap.add_argument("--candidates", type=int, default=512)
    ap.add_argument("--bodies", type=int, default=8)
    ap.add_argument("--composed", type=int, default=2)
    ap.add_argument("--score", type=int, default=0)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--engine", default="tcgen05")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    if a.evals:
        model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
        dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
        model.load_state_dict(init_unet_params(seed=0))
        dif.to(dev)
        dif.precision, dif.conv_engine = a.precision, a.engine
        x = torch.randn(a.candidates, 24 + 10 * a.composed, 4 * a.bodies, device=dev)
        for _ in range(a.evals):
            dif.composed_eps(x, 500, a.composed, 10, a.bodies, "mean-inside")
        torch.cuda.synchronize()
    if a.score:
        rng = np.random.default_rng(0)
        b = a.score
        pred = torch.from_numpy(rng.uniform(0.12, 0.88, size=(b, 44, 32)).astype(np.float32)).to(dev)
        pred[..., 2::4] = torch.from_numpy(rng.uniform(-0.5, 0.5, size=(b, 44, 8)).astype(np.float32)).to(dev)
        pred[..., 3::4] = torch.from_numpy(rng.uniform(-0.5, 0.5, size=(b, 44, 8)).astype(np.float32)).to(dev)
        score_designs(pred)
        torch.cuda.synchronize()
    print("probe done", _lib.lib().cindm_launch_count())


if __name__ == "__main__":
    main()
