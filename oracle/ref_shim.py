"""Import the UNMODIFIED reference `model/diffusion_1d.py` for validation only.

TEST INFRASTRUCTURE.  Only `oracle/make_golden.py` and the CPU tests that pin the
oracle use this module, and only in the build container where `/root/reference`
is mounted (the GPU box never has it).  Seven third-party imports that the
reference pulls in at module import time (model/diffusion_1d.py:1-31) but never
touches on the sampling path are replaced with empty stand-ins in `sys.modules`.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CINDM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "diffusion_1d.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        mod = sys.modules[name]
    else:
        mod = types.ModuleType(name)
        mod.__path__ = []  # behave like a package so that sub-imports resolve
        sys.modules[name] = mod
    for k, v in attrs.items():
        setattr(mod, k, v)
    return mod


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return None

    def __getattr__(self, item):
        return _Anything()


_loaded = None


def load():
    """Return the reference module object (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise FileNotFoundError("reference tree not present at " + REFERENCE_ROOT)
    for name in ("accelerate", "ema_pytorch", "imageio"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name, Accelerator=_Anything, EMA=_Anything, imwrite=_Anything())
    try:
        importlib.import_module("torch_geometric.data.dataloader")
    except Exception:
        _stub("torch_geometric")
        _stub("torch_geometric.data")
        _stub("torch_geometric.data.dataloader", DataLoader=_Anything)
    try:
        importlib.import_module("matplotlib.pyplot")
        importlib.import_module("matplotlib.backends.backend_pdf")
    except Exception:
        plt = _Anything()
        _stub("matplotlib", pyplot=plt, pylab=plt)
        _stub("matplotlib.pyplot")
        _stub("matplotlib.pylab")
        _stub("matplotlib.backends")
        _stub("matplotlib.backends.backend_pdf", PdfPages=_Anything)
    # the reference expects to live in a package called `cindm`
    _stub("cindm")
    _stub("cindm.data")
    _stub("cindm.data.nbody_dataset", NBodyDataset=_Anything)
    _stub("cindm.utils", p=_Anything(), get_item_1d=_Anything(), COLOR_LIST=[], CustomLoss=_Anything,
          Printer=_Anything, CustomSampler=_Anything, visulization=_Anything())
    _stub("cindm.filepath", EXP_PATH="/tmp")
    path = os.path.join(REFERENCE_ROOT, "model", "diffusion_1d.py")
    spec = importlib.util.spec_from_file_location("_cindm_reference_diffusion_1d", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded = mod
    return mod
