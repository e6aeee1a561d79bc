"""Per-kernel CUDA-event timings of ONE composed evaluation at a small slice count (C1: 50 slices by default).

    python profiles/small_batch_profile.py [batch] [n_bodies] [n_composed] [precision engine]     (default fp16 tcgen05)
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cindm_b200 import _lib
from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
from cindm_b200.model.params import init_unet_params


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nc = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
    model.load_state_dict(init_unet_params(seed=0))
    dif.to("cuda:0")
    prec = sys.argv[4] if len(sys.argv) > 4 else "fp16"
    engine = sys.argv[5] if len(sys.argv) > 5 else "tcgen05"
    dif.precision, dif.conv_engine = prec, engine
    model.precision, model.conv_engine = prec, engine
    L = _lib.lib()
    x = torch.randn(b, 24 + 10 * nc, 4 * n, device="cuda")
    for rep in range(3):
        if rep == 2:
            _lib.check(L.cindm_profile_enable(1))
        dif.composed_eps(x, 500, nc, 10, n)
    torch.cuda.synchronize()
    buf = ctypes.create_string_buffer(1 << 18)
    L.cindm_profile_report(buf, len(buf))
    _lib.check(L.cindm_profile_enable(0))
    rows = [r.split(",") for r in buf.value.decode().strip().splitlines()]
    rows = [(r[0], int(r[1]), float(r[2])) for r in rows if len(r) >= 3]
    total = sum(r[2] for r in rows)
    print(f"{b} candidates, {n} bodies, {nc + 1} windows: {sum(r[1] for r in rows)} launch groups, {total:.3f} ms (event-timed, un-graphed)")
    for tag, cnt, ms in sorted(rows, key=lambda r: -r[2]):
        print(f"  {tag:48s} {cnt:3d} x {1e3 * ms / cnt:8.2f} us")


if __name__ == "__main__":
    main()
