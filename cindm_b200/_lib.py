"""ctypes binding of libcindm_b200.so (include/cindm_b200.h).  No CPU fallback: if the library
is missing the import of any compute entry point fails loudly."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# CINDM_B200_LIB points at another build of the same library (A/B measurements of two kernel versions in one job)
LIB_PATH = os.environ.get("CINDM_B200_LIB") or os.path.join(HERE, "lib", "libcindm_b200.so")

PREC_F32, PREC_F16, PREC_BF16 = 0, 1, 2
CONV_SIMT, CONV_TCGEN05 = 0, 1
COMPOSE_MEAN_INSIDE, COMPOSE_SUM_INSIDE, COMPOSE_MEAN_OUTSIDE, COMPOSE_NOISE_SUM, COMPOSE_EBM = 0, 1, 2, 3, 4
OBJ_L2, OBJ_L2SQUARE = 0, 1
GUIDE_NONE, GUIDE_STANDARD, GUIDE_STANDARD_ALPHA = 0, 1, 2

PRECISIONS = {"fp32": PREC_F32, "f32": PREC_F32, "fp16": PREC_F16, "f16": PREC_F16, "bf16": PREC_BF16}


class Config(Structure):
    _fields_ = [("horizon", c_int), ("transition_dim", c_int), ("dim", c_int), ("timesteps", c_int)]


class Objective(Structure):
    _fields_ = [("target_x", c_double), ("target_y", c_double), ("coef", c_float), ("consistency_coef", c_float),
                ("mode", c_int), ("guidance", c_int)]


class SampleConfig(Structure):
    _fields_ = [("batch", c_int), ("n_bodies", c_int), ("n_composed", c_int), ("compose_start_step", c_int),
                ("compose_mode", c_int), ("recurrence", c_int), ("precision", c_int), ("conv_engine", c_int),
                ("t_start", c_int), ("t_end", c_int), ("seed", c_uint64), ("candidate_offset", c_int64),
                ("use_graph", c_int), ("objective", Objective), ("cond_rows", c_int), ("chain_blocks", c_int),
                ("ebm_uncond_coef", c_float)]


class CindmError(RuntimeError):
    pass


_SIGNATURES = {
    "cindm_last_error": (c_char_p, []),
    "cindm_version": (c_int, []),
    "cindm_launch_count": (ctypes.c_longlong, []),
    "cindm_profile_enable": (c_int, [c_int]),
    "cindm_profile_report": (c_int, [c_char_p, c_int]),
    "cindm_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "cindm_destroy": (c_int, [c_void_p]),
    "cindm_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, POINTER(c_int64), c_int]),
    "cindm_finalize_weights": (c_int, [c_void_p, c_void_p]),
    "cindm_set_schedule": (c_int, [c_void_p, c_void_p, c_int]),
    "cindm_reserve": (c_int, [c_void_p, c_int64, c_int]),
    "cindm_workspace_bytes": (c_int64, [c_int64, c_int]),
    "cindm_model_workspace_bytes": (c_int64, [c_int, c_int, c_int64, c_int]),
    "cindm_schedule_tables": (c_int, [c_int, c_void_p]),
    "cindm_build_index_maps": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32),
                                       POINTER(c_int32), POINTER(c_int32)]),
    "cindm_compose_gather": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cindm_compose_scatter_mean": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cindm_unet_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int, c_void_p]),
    "cindm_unet_read_tap": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, POINTER(c_int64), POINTER(c_int64),
                                    POINTER(c_int64)]),
    "cindm_unet_enable_taps": (c_int, [c_void_p, c_int]),
    "cindm_composed_eps": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p]),
    "cindm_design_grad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(Objective), c_void_p]),
    "cindm_posterior_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, POINTER(Objective), c_void_p]),
    "cindm_attach_unconditioned": (c_int, [c_void_p, c_void_p]),
    "cindm_ebm_eps": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p]),
    "cindm_ula_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float, c_uint64, c_int64,
                               c_int, c_int, c_void_p]),
    "cindm_predict_start": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "cindm_sample": (c_int, [c_void_p, POINTER(SampleConfig), c_void_p, c_void_p, c_void_p, c_void_p]),
    "cindm_composed_posterior": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_void_p]),
    "cindm_set_initial_state_overwrite": (c_int, [c_void_p, c_void_p, c_int]),
    "cindm_sample_ddim": (c_int, [c_void_p, POINTER(SampleConfig), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p]),
    "cindm_fill_initial_noise": (c_int, [c_void_p, c_int, c_int, c_int, c_uint64, c_int64, c_int, c_void_p]),
    "cindm_nbody_rollout": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "cindm_score_designs": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p]),
}

_lib = None


def exported_symbols():
    """Names include/cindm_b200.h declares (used by the CPU test that checks the .so exports them)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CindmError(
                f"{LIB_PATH} is missing: build it with `python -m cindm_b200.build` "
                "(there is no CPU or PyTorch fallback for the sampling path)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().cindm_last_error()
        raise CindmError(f"cindm_b200 error {rc}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device / host pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous()
    return c_void_p(t.data_ptr())


def stream_ptr(device=None):
    import torch
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)
