"""Host-side mirror of the reference helpers the sampling driver uses (utils.py), backed by the CUDA
rollout kernel instead of pymunk.  Names and argument meaning follow the reference:
simulation (utils.py:1071-1125), eval_simu (:1127-1148), caculate_confidence_interval (:1215-1239),
setup_seed (:1257-1262), get_item_1d layout (:203-223)."""
import random

import numpy as np
import torch

from .. import _lib


def setup_seed(seed):
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)


def simulation(features, n_steps, filename=None, width=200, height=200, radius=20, mass=1, stride=1, device=None):
    """features: [B, n_bodies, 4] (x, y, vx, vy in pixel units) -> float64 tensor [B, n_steps // stride, n_bodies, 4].

    With stride == 1 this is the reference's `simulation`: entry k is the state after k steps.  The
    rendering / .npy side effects of `filename` are not reproduced."""
    if filename is not None:
        raise NotImplementedError("rendering / saving trajectories is outside the CUDA scoring path")
    if (width, height, radius, mass) != (200, 200, 20, 1):
        raise NotImplementedError("the rollout kernel is built for the reference's world: 200x200 box, r=20, m=1")
    dev = torch.device(device) if device is not None else (features.device if features.is_cuda else torch.device("cuda"))
    state0 = features.detach().to(dev, torch.float64).contiguous()
    b, n, four = state0.shape
    assert four == 4
    traj = torch.empty((b, n_steps // stride, n, 4), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().cindm_nbody_rollout(_lib.ptr(state0), _lib.ptr(traj), b, n, n_steps, stride, _lib.stream_ptr(dev)))
    return traj


def eval_simu(cond_design, design_fn, n_bodies, rollout_steps, time_interval=4):
    """cond_design: [B, conditioned_steps, n_bodies*4] -> (pred_simu [B, rollout_steps, n_bodies*4] float64, design_fn(pred_simu)).

    As the reference: the LAST conditioning frame x200 is simulated for rollout_steps*time_interval steps and
    the states after time_interval-1, 2*time_interval-1, ... steps are kept (traj[:, time_interval-1::time_interval])."""
    assert cond_design.shape[-1] // 4 == n_bodies
    cond_simu = cond_design[:, -1, :] * 200.0
    cond_simu = cond_simu.reshape(cond_simu.shape[0], n_bodies, -1)
    pred_simu = simulation(cond_simu, rollout_steps * time_interval, stride=time_interval)
    pred_simu = pred_simu.reshape(pred_simu.shape[0], pred_simu.shape[1], -1)
    # true fp64 division (torch's CUDA scalar-divisor fast path multiplies by the reciprocal instead)
    pred_simu = torch.div(pred_simu.to(cond_design.device), torch.full((), 200.0, dtype=torch.float64, device=pred_simu.device))
    return pred_simu, design_fn(pred_simu)


def score_designs(pred, pos_target=(0.5, 0.5)):
    """Fused scoring of generated designs pred [B, T, 4n] fp32: per-candidate (MAE over all T*4n entries of
    cat(frame0, simulated) vs pred, mean over bodies of the last simulated frame's distance to the target),
    i.e. the per-sample terms of the driver's MAE and design_obj_simu (inverse_design_diffusion_1d.py:316-337)."""
    pred = pred.detach().to(torch.float32).contiguous()
    assert pred.is_cuda
    b, t, f = pred.shape
    mae = torch.empty(b, dtype=torch.float64, device=pred.device)
    obj = torch.empty(b, dtype=torch.float64, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(_lib.lib().cindm_score_designs(_lib.ptr(pred), _lib.ptr(mae), _lib.ptr(obj), b, t, f // 4,
                                                  float(pos_target[0]), float(pos_target[1]), _lib.stream_ptr(pred.device)))
    return mae, obj


def caculate_confidence_interval(data):
    """(mean, std, 95% margin, min) of per-sample values (reference spelling kept)."""
    per = data if data.dim() <= 1 else data.mean(dim=tuple(range(1, data.dim())))
    mean, std = per.mean(), per.std()
    margin = std * 1.96 / torch.sqrt(torch.tensor(len(data), dtype=torch.float64))
    return mean, std, margin, per.min()


def get_item_1d_from_array(traj, time_stride=4):
    """Dataset layout of get_item_1d: pixel-unit trajectories [B, steps, n, 4] -> [B, steps/stride, n*4] / 200."""
    x = torch.as_tensor(traj)[:, ::time_stride] / 200.0
    return x.reshape(x.shape[0], x.shape[1], -1)
