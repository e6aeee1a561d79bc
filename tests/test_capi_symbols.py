"""CPU: the C-ABI library builds with nvcc for sm_100a, loads, and exports every symbol that
include/cindm_b200.h declares.  No compute entry point is called here (there is no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    from cindm_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cindm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cindm_[a-z0-9_]+)\s*\(", text)))


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/cindm_b200.h must compile as C99 (no C++, no torch types) and a C program that
    links the library must resolve every entry point it names."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "use_header.c"
    calls = "\n".join(f"    p[{i}] = (void*){name};" for i, name in enumerate(declared_symbols()))
    src.write_text('#include "cindm_b200.h"\n#include <stdio.h>\nint main(void) {\n'
                   f"    void* p[{len(declared_symbols())}];\n{calls}\n"
                   '    cindm_config c = {24, 8, 64, 1000};\n    printf("%d %p\\n", c.horizon + cindm_version(), p[0]);\n    return 0;\n}\n')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-Wno-pedantic", "-I", inc, "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_program_links_and_runs_the_host_entry_points(library, tmp_path):
    """examples/host_entry_points.c: a C99 consumer of the shared library (schedule tables, index maps, error reporting:
    the entry points that need no GPU) builds against include/cindm_b200.h and passes its own known-answer checks."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not on PATH")
    exe = tmp_path / "host_entry_points"
    libdir = os.path.dirname(library)
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "host_entry_points.c"), "-o", str(exe), "-L", libdir, "-lcindm_b200",
                        f"-Wl,-rpath,{libdir}", "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok:"), r.stdout + r.stderr


def test_header_and_binding_agree():
    from cindm_b200 import _lib
    assert declared_symbols() == _lib.exported_symbols()


def test_library_exports_every_declared_symbol(library):
    handle = ctypes.CDLL(library)
    for name in declared_symbols():
        assert hasattr(handle, name), name
    handle.cindm_version.restype = ctypes.c_int
    assert handle.cindm_version() >= 100


def test_host_only_entry_points(library):
    """Schedule tables and index maps are pure host code: check them against the golden vectors."""
    import numpy as np
    import torch
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import SCHEDULE_KEYS, schedule_buffers, cosine_beta_schedule
    g = np.load(os.path.join(ROOT, "tests", "golden", "schedule.npz"))
    out = torch.empty(13, 1000, dtype=torch.float32)
    _lib.check(_lib.lib().cindm_schedule_tables(1000, ctypes.c_void_p(out.data_ptr())))
    bufs = schedule_buffers(cosine_beta_schedule(1000))
    for i, k in enumerate(SCHEDULE_KEYS):
        ref = torch.from_numpy(g[k])
        assert torch.equal(bufs[k], ref), k                       # host mirror: bit-exact
        ulp = (out[i].view(torch.int32) - ref.view(torch.int32)).abs().max().item()
        assert ulp <= 1, (k, ulp)                                  # C libm vs torch fp64: at most one fp32 ulp
    gi = np.load(os.path.join(ROOT, "tests", "golden", "index_maps.npz"))
    cover = (ctypes.c_int32 * 44)()
    _lib.check(_lib.lib().cindm_build_index_maps(8, 2, 10, 24, None, None, None, cover))
    assert list(cover) == gi["n8_nc2_s10:cover"].tolist()
    assert list(cover) == [1] * 10 + [2] * 10 + [3] * 4 + [2] * 10 + [1] * 10


def test_index_maps_host_entry_point_vs_oracle_property(library):
    """cindm_build_index_maps is pure host code: window starts, lexicographic pair list and cover counts equal the oracle's
    (itself pinned to maps recovered by probing the reference, tests/golden/index_maps.npz) for ANY body count, window count,
    window stride and model horizon -- 24, the 44-step models, the single-step model."""
    from hypothesis import given, settings, strategies as st
    from cindm_b200 import _lib
    from oracle import sampler_ref
    L = _lib.lib()

    @settings(max_examples=200, deadline=None)
    @given(n=st.integers(2, 10), nc=st.integers(0, 6), horizon=st.sampled_from([8, 10, 20, 24, 44, 48]), data=st.data())
    def check(n, nc, horizon, data):
        start = data.draw(st.integers(1, horizon - 1))                    # the reference asserts compose_start_step < horizon (:1679)
        ref = sampler_ref.index_maps(n, nc, start, horizon)
        W, P, T = nc + 1, n * (n - 1) // 2, horizon + nc * start
        win, pi, pj, cover = (ctypes.c_int32 * W)(), (ctypes.c_int32 * P)(), (ctypes.c_int32 * P)(), (ctypes.c_int32 * T)()
        _lib.check(L.cindm_build_index_maps(n, nc, start, horizon, win, pi, pj, cover))
        assert list(win) == ref["win_t0"]
        assert list(zip(pi, pj)) == ref["pairs"]
        assert list(cover) == ref["cover"] and ref["t_total"] == T
        assert sum(cover) == W * horizon                                  # every window row is counted exactly once

    check()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from cindm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.CindmError):
        _lib.lib()


def test_host_logic_parsers():
    from cindm_b200 import _lib
    from cindm_b200.model.diffusion_1d import parse_design_guidance, get_design_fn
    assert parse_design_guidance("standard") == (_lib.GUIDE_STANDARD, 0)
    assert parse_design_guidance("standard-recurrence-10") == (_lib.GUIDE_STANDARD, 10)
    assert parse_design_guidance("standard-alpha-recurrence-3") == (_lib.GUIDE_STANDARD_ALPHA, 3)
    with pytest.raises(ValueError):
        parse_design_guidance("standard-recurrence-__import__('os')")
    with pytest.raises(NotImplementedError):
        parse_design_guidance("universal-forward-recurrence-5")
    import torch
    fn = get_design_fn(torch.tensor([0.5, 0.5], dtype=torch.float64), 1, coef=0.2, time_consistency_coef=0.2)
    s = fn.as_struct(_lib.GUIDE_STANDARD)
    assert (s.target_x, s.target_y, s.mode) == (0.5, 0.5, _lib.OBJ_L2)
    assert abs(s.coef - 0.2) < 1e-7


def test_state_dict_layout_matches_reference_keys():
    from cindm_b200.model.diffusion_1d import GaussianDiffusion1D, TemporalUnet1D
    model = TemporalUnet1D(horizon=24, transition_dim=8, cond_dim=False, dim=64, dim_mults=(1, 2, 4, 8), attention=True)
    dif = GaussianDiffusion1D(model, image_size=24, conditioned_steps=0, timesteps=1000, sampling_timesteps=1000)
    sd = dif.state_dict()
    assert len(sd) == 247                                           # 234 U-Net entries + 13 buffers
    assert tuple(sd["model.downs.0.0.blocks.0.block.0.weight"].shape) == (64, 8, 5)
    assert tuple(sd["model.downs.0.2.fn.norm.g"].shape) == (1, 64, 1)
    assert tuple(sd["model.ups.2.3.conv.weight"].shape) == (64, 64, 4)
    dif.load_state_dict(sd)
    with pytest.raises(RuntimeError):
        dif.load_state_dict({k: v for k, v in sd.items() if k != "model.final_conv.1.bias"})


def test_workspace_size_follows_the_level_structure(library):
    """Host-only: a horizon % 8 model holds horizon * dim values per slice on every level; the 44-step models keep 11 positions
    on the 512-channel level (11 * 8 * dim > 44 * dim), so their activation buffers are sized by that level."""
    from cindm_b200 import _lib
    L = _lib.lib()
    assert L.cindm_model_workspace_bytes(24, 64, 1000, 0) == L.cindm_workspace_bytes(1000, 0)
    per_slice = lambda h, d: (L.cindm_model_workspace_bytes(h, d, 4096, 0) - L.cindm_model_workspace_bytes(h, d, 2048, 0)) / 2048
    # 9 activation-sized fp32 buffers + qkv (384) + att (128) + slices / eps_pair (2 x 8) per position
    assert per_slice(24, 64) == pytest.approx(4 * (9 * 24 * 64 + 24 * (384 + 128 + 16)), rel=1e-3)
    assert per_slice(44, 64) == pytest.approx(4 * (9 * 11 * 512 + 44 * (384 + 128 + 16)), rel=1e-3)
    assert per_slice(44, 96) == pytest.approx(4 * (9 * 11 * 768 + 44 * (384 + 128 + 16)), rel=1e-3)


def test_c_abi_reports_errors_instead_of_throwing(library):
    """Argument validation happens before any CUDA call: negative code + message, nothing thrown across the boundary."""
    from cindm_b200 import _lib
    L = _lib.lib()
    handle = ctypes.c_void_p()
    # odd horizon (undefined in the reference, :549-554), horizon / dim out of range, a dim GroupNorm(8) cannot split, ...
    for bad in (_lib.Config(45, 8, 64, 1000), _lib.Config(64, 8, 64, 1000), _lib.Config(24, 8, 100, 1000), _lib.Config(24, 8, 256, 1000),
                _lib.Config(24, 16, 64, 1000), _lib.Config(24, 8, 64, 0)):
        rc = L.cindm_create(ctypes.byref(bad), ctypes.byref(handle))
        assert rc < 0 and L.cindm_last_error()
    assert L.cindm_create(None, ctypes.byref(handle)) < 0
    assert L.cindm_build_index_maps(1, 0, 10, 24, None, None, None, None) < 0          # fewer than two bodies
    assert L.cindm_build_index_maps(4, -1, 10, 24, None, None, None, None) < 0
    assert L.cindm_schedule_tables(0, None) < 0
    with pytest.raises(_lib.CindmError):
        _lib.check(L.cindm_schedule_tables(0, None))
    assert L.cindm_destroy(None) == 0
