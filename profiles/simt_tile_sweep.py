"""fp32 SIMT conv tile dispatch: time of one U-Net evaluation per slice count and (MIN128, MIN64) CTA thresholds, and a check
that every dispatch gives bit-identical epsilon.  The switches are read per launch, so one process sweeps them.

    python profiles/simt_tile_sweep.py > gpurun_out/simt_tile_sweep.json
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
from cindm_b200.model.params import init_unet_params, unet_param_shapes

SETTINGS = {                       # name: environment
    "round-1 dispatch (no 32x32)": {"CINDM_SIMT_TILE32": "0"},
    "default (min128=74 min64=148)": {},
    "min128=37 min64=74": {"CINDM_SIMT_MIN128": "37", "CINDM_SIMT_MIN64": "74"},
    "min128=74 min64=148": {"CINDM_SIMT_MIN128": "74", "CINDM_SIMT_MIN64": "148"},
    "min128=148 min64=296": {"CINDM_SIMT_MIN128": "148", "CINDM_SIMT_MIN64": "296"},
}
# large slice counts: position-major tiles of the 128-row kernel (taps that are all zero padding skipped) on / off
SETTINGS_BIG = {
    "row-major tiles (CINDM_SIMT_POSMAJOR=0)": {"CINDM_SIMT_POSMAJOR": "0"},
    "default (position-major at H <= 8)": {},
}
KEYS = ("CINDM_SIMT_TILE32", "CINDM_SIMT_MIN128", "CINDM_SIMT_MIN64", "CINDM_SIMT_POSMAJOR")


def main():
    out = {}
    for hor, dim, sizes in ((24, 64, (50, 150, 500, 1500, 3000, 43008)), (44, 64, (50, 500, 3000)), (44, 96, (50,))):
        model = TemporalUnet1D(horizon=hor, transition_dim=8, cond_dim=False, dim=dim, dim_mults=(1, 2, 4, 8), attention=True)
        dif = GaussianDiffusion1D(model, image_size=hor, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
        model.load_state_dict(init_unet_params(unet_param_shapes(hor, 8, dim), seed=0, randomize_affine=True))
        dif.to("cuda:0")
        for S in sizes:
            x = torch.randn(S, hor, 8, generator=torch.Generator().manual_seed(S)).cuda()
            t = torch.full((S,), 420, dtype=torch.long)
            ref = None
            row = {}
            for name, env in (SETTINGS_BIG if S >= 3000 else SETTINGS).items():
                for k in KEYS:
                    os.environ.pop(k, None)
                os.environ.update(env)
                for _ in range(2):
                    y = model(x, t, None)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    y = model(x, t, None)
                e1.record()
                torch.cuda.synchronize()
                y = y.cpu()
                if ref is None:
                    ref = y
                row[name] = {"ms_per_evaluation": e0.elapsed_time(e1) / 5, "bit_identical_to_first": bool(torch.equal(y, ref))}
            out[f"horizon {hor} dim {dim} S={S}"] = row
        model._drop_engine()
    for k in KEYS:
        os.environ.pop(k, None)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
