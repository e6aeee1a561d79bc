"""The ctypes stub INTEGRATION.md section 2 tells a reference maintainer to add is EXECUTED here, verbatim from the
document: on CPU against the reference's real GaussianDiffusion1D (through oracle/ref_shim.py, when /root/reference is
mounted) up to the first call that needs a device, and on the GPU end to end against the golden composed epsilon."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    section = text.split("## 2.")[1].split("## 3.")[0]
    blocks = re.findall(r"```python\n(.*?)```", section, flags=re.S)
    assert len(blocks) == 1, "INTEGRATION.md section 2 holds exactly one python block: the stub"
    code = blocks[0].replace('"cindm_b200/lib/libcindm_b200.so"', repr(os.path.join(ROOT, "cindm_b200", "lib", "libcindm_b200.so")))
    ns = {}
    exec(compile(code, "INTEGRATION.md#2", "exec"), ns)
    return ns


class _StandIn:
    """What attach_b200 touches of a reference GaussianDiffusion1D: .model.state_dict(), .num_timesteps, the 13 buffers."""

    def __init__(self, weights):
        from cindm_b200.model.diffusion_1d import cosine_beta_schedule, schedule_buffers
        self.model = types.SimpleNamespace(state_dict=lambda: weights)
        self.num_timesteps = 1000
        for k, v in schedule_buffers(cosine_beta_schedule(1000)).items():
            setattr(self, k, v)


def reference_or_stand_in(weights):
    from oracle import ref_shim
    if ref_shim.available():
        from oracle import make_golden
        _, _, dif = make_golden.build_reference(weights)
        return dif, True
    return _StandIn(weights), False


def test_stub_binds_the_reference_classes_up_to_the_device_boundary(test_weights):
    ns = stub_namespace()
    dif, real = reference_or_stand_in(test_weights)
    L = ns["L"]
    loaded = []
    real_load = L.cindm_load_weight

    def counting_load(*a):
        rc = real_load(*a)
        loaded.append(rc)
        return rc

    L.cindm_load_weight = counting_load
    try:
        if torch.cuda.is_available():
            ns["attach_b200"](dif)
            assert dif._b200
        else:
            with pytest.raises(RuntimeError) as err:           # cindm_finalize_weights needs a device; nothing falls back
                ns["attach_b200"](dif)
            assert "cuda" in str(err.value).lower() or "device" in str(err.value).lower()
    finally:
        L.cindm_load_weight = real_load
    # every U-Net state-dict entry of the (real) reference module was accepted by name and shape
    assert len(loaded) == 234 and not any(loaded)
    if real:
        assert type(dif).__name__ == "GaussianDiffusion1D" and len(dif.state_dict()) == 247


@pytest.mark.gpu
def test_stub_composed_eps_matches_the_reference_golden(test_weights, golden):
    ns = stub_namespace()
    dif, _ = reference_or_stand_in(test_weights)
    ns["attach_b200"](dif)
    meta = __import__("json").load(open(os.path.join(ROOT, "tests", "golden", "meta.json")))
    g = golden("composed_eps.npz")
    for name in ("c4_8body_w3", "c3_4body_w1"):
        n, nc, start, mode, b, t = meta["compose_cases"][name]
        x = torch.from_numpy(g[name + ":x"]).cuda()
        eps = ns["composed_eps"](dif, x, t, nc, start, n, mean_inside=(mode == "mean-inside")).cpu().double()
        ref = torch.from_numpy(g[name + ":eps"]).double()
        assert ((eps - ref).norm() / ref.norm()).item() < 1e-2          # fp16 operands, tcgen05 convs (the stub's choice)
