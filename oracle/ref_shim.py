"""Import the UNMODIFIED reference modules from /root/reference (this container
only -- the path does not exist on the GPU box).  TEST INFRASTRUCTURE.

``histogram_classes.RGBuvHistBlock`` imports directly.  ``histoGAN.histoGAN``
needs three un-vendored packages and a CUDA assert bypassed
(histoGAN/histoGAN.py:28,32,33,46); we register empty stub modules and patch
``torch.cuda.is_available`` only for the duration of the import.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("HISTOGAN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "histogram_classes",
                                       "RGBuvHistBlock.py"))


def _ensure_path():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def ref_hist_module():
    _ensure_path()
    return importlib.import_module("histogram_classes.RGBuvHistBlock")


def ref_gan_module():
    """``histoGAN.histoGAN`` of the reference, importable on a CPU-only box."""
    import torch
    _ensure_path()
    if "histoGAN.histoGAN" in sys.modules and getattr(
            sys.modules["histoGAN.histoGAN"], "__file__", "").startswith(REF_ROOT):
        return sys.modules["histoGAN.histoGAN"]

    def _stub(name, **attrs):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(m, k, v)
            sys.modules[name] = m

    class _Missing:  # placeholder classes; never instantiated by the tests
        def __init__(self, *a, **k):
            raise RuntimeError("stubbed third-party class")

    _stub("torch_optimizer", DiffGrad=_Missing)
    _stub("vector_quantize_pytorch", VectorQuantize=_Missing)
    _stub("linear_attention_transformer", ImageLinearAttention=_Missing)
    saved = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        mod = importlib.import_module("histoGAN.histoGAN")
    finally:
        torch.cuda.is_available = saved
    return mod


def ref_rehisto_module():
    """``ReHistoGAN.rehistoGAN`` of the reference (needs the same stubs as histoGAN.histoGAN)."""
    import torch
    ref_gan_module()
    saved = torch.cuda.is_available
    torch.cuda.is_available = lambda: True
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return importlib.import_module("ReHistoGAN.rehistoGAN")
    finally:
        torch.cuda.is_available = saved
