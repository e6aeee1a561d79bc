"""ctypes wrapper of oracle/nbody_ref.c plus the driver-side scoring arithmetic (TEST INFRASTRUCTURE).

eval_simu / the metrics of inference/inverse_design_diffusion_1d.py:316-337 restated with numpy:
frame 0 x 200 (fp32 product) -> rollout of (T-1)*4 steps -> frames [3::4] / 200 -> design objective
(mean over bodies of the last frame's distance to the target) and MAE over all T*4n entries.
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.nbody_ref_rollout.restype = None
        _lib.nbody_ref_rollout.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib.nbody_ref_set_order.restype = None
        _lib.nbody_ref_set_order.argtypes = [ctypes.c_int, ctypes.c_ulonglong]
        _lib.nbody_ref_census.restype = None
        _lib.nbody_ref_census.argtypes = [ctypes.POINTER(ctypes.c_long), ctypes.POINTER(ctypes.c_long)]
    return _lib


ORDERS = {"walls-first lexicographic (default)": 0, "pairs before walls": 1, "reverse": 2, "random permutation per step": 3,
          "body-major": 4}


def set_contact_order(mode=0, seed=1):
    """Contact solve order of subsequent rollouts (see nbody_ref.c); 0 is what the CUDA kernel implements."""
    _load().nbody_ref_set_order(int(mode), int(seed))


def census():
    """(steps with two or more active contacts sharing a body, steps with any contact) of the last rollout call."""
    a, b = ctypes.c_long(), ctypes.c_long()
    _load().nbody_ref_census(ctypes.byref(a), ctypes.byref(b))
    return a.value, b.value


def rollout(state0, n_steps, stride=1):
    """state0: [B, n, 4] float64 pixel units -> [B, n_steps // stride, n, 4] (states after stride-1, 2*stride-1, ... steps)."""
    state0 = np.ascontiguousarray(state0, dtype=np.float64)
    b, n, _ = state0.shape
    out = np.empty((b, n_steps // stride, n, 4), dtype=np.float64)
    _load().nbody_ref_rollout(state0.ctypes.data, out.ctypes.data, b, n, n_steps, stride)
    return out


def score_designs(pred, target=(0.5, 0.5)):
    """pred: [B, T, 4n] float32 (normalised units) -> (pred_simu [B, T-1, 4n] f64, mae [B], objective [B])."""
    pred = np.asarray(pred, dtype=np.float32)
    b, t, f = pred.shape
    n = f // 4
    state0 = (pred[:, 0, :] * np.float32(200.0)).astype(np.float64).reshape(b, n, 4)
    sim = rollout(state0, (t - 1) * 4, 4).reshape(b, t - 1, f) / 200.0
    full = np.concatenate([pred[:, :1].astype(np.float64), sim], axis=1)
    mae = np.abs(full - pred.astype(np.float64)).mean(axis=(1, 2))
    last = sim[:, -1].reshape(b, n, 4)[:, :, :2]
    obj = np.sqrt(((last - np.asarray(target, dtype=np.float64)) ** 2).sum(-1)).mean(-1)
    return sim, mae, obj
