"""ctypes binding of libhistogan_b200.so (the C ABI in include/histogan_b200.h).

There is no CPU or PyTorch fallback: if the shared library cannot be loaded the
import of any operator raises, and every call checks the returned status.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libhistogan_b200.so"

HG_ABI_VERSION = 4

RESIZE_IDS = {"interpolation": 0, "sampling": 1}
METHOD_IDS = {"thresholding": 0, "RBF": 1, "inverse-quadratic": 2}


class HistParams(C.Structure):
    """struct hg_hist_params (include/histogan_b200.h)."""
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("sb", C.c_int64), ("sc", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64),
        ("h", C.c_int32), ("insz", C.c_int32), ("resizing", C.c_int32), ("method", C.c_int32),
        ("sigma", C.c_double), ("lo", C.c_double), ("hi", C.c_double),
        ("intensity_scale", C.c_int32), ("green_only", C.c_int32), ("projection", C.c_int32),
    ]


class ConvParams(C.Structure):
    """struct hg_conv_params."""
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("Cout", C.c_int32), ("KH", C.c_int32), ("KW", C.c_int32), ("stride", C.c_int32),
                ("pad", C.c_int32), ("OH", C.c_int32), ("OW", C.c_int32)]


class ConvEpilogue(C.Structure):
    """struct hg_conv_epilogue."""
    _fields_ = [("scale", C.c_void_p), ("bias", C.c_void_p), ("noise", C.c_void_p),
                ("noise_w", C.c_void_p), ("noise_b", C.c_void_p), ("residual", C.c_void_p),
                ("noise_size", C.c_int32), ("flags", C.c_int32), ("lrelu_slope", C.c_float),
                ("reserved_", C.c_int32), ("out_img_stride", C.c_int64),
                ("out_row_stride", C.c_int64), ("out_pix_stride", C.c_int64)]


_SIGNATURES = {
    "hg_abi_version": (C.c_int, []),
    "hg_last_error": (C.c_char_p, []),
    "hg_launch_count": (C.c_uint64, []),
    "hg_device_check": (C.c_int, [C.c_int]),
    "hg_hist_num_pixels": (C.c_int64, [C.POINTER(HistParams)]),
    "hg_hist_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(HistParams)]),
    "hg_hist_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(HistParams)]),
    "hg_hist_fwd": (C.c_int, [C.c_void_p, C.POINTER(HistParams), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_size_t, C.c_void_p]),
    "hg_hist_bwd": (C.c_int, [C.c_void_p, C.POINTER(HistParams), C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hg_hist_preprocess": (C.c_int, [C.c_void_p, C.POINTER(HistParams), C.c_void_p, C.c_void_p]),
    "hg_debug_logf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hg_conv2d_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ConvParams),
                                C.POINTER(ConvEpilogue), C.c_void_p]),
    "hg_conv2d_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(ConvParams),
                                  C.c_void_p]),
    "hg_unpack_conv_wgrad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_void_p]),
    "hg_modulate_round": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_int32, C.c_int32, C.c_void_p]),
    "hg_channel_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p]),
    "hg_modconv_epilogue_bwd": (C.c_int, [C.c_void_p] * 10 + [C.c_int32] * 5 + [C.c_float, C.c_void_p]),
    "hg_modulate_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p]),
    "hg_upsample_modulate_round": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 4 + [C.c_void_p]),
    "hg_upsample_modulate_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "hg_torgb_fwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p]),
    "hg_torgb_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_void_p]),
    "hg_grouped_linear_fwd": (C.c_int, [C.c_int32] + [C.c_void_p] * 6 + [C.c_int32, C.c_int32, C.c_float,
                                                                       C.c_float, C.c_void_p]),
    "hg_grouped_linear_bwd": (C.c_int, [C.c_int32] + [C.c_void_p] * 8 + [C.c_int32, C.c_int32, C.c_void_p]),
    "hg_weight_sqsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hg_demod_bwd": (C.c_int, [C.c_void_p] * 8 + [C.c_int32] * 4 + [C.c_void_p]),
    "hg_conv_small_fwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 7 + [C.c_int64] * 4 + [C.c_int32, C.c_float,
                                                                                      C.c_void_p]),
    "hg_conv_small_dgrad": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 7 + [C.c_int64] * 4 + [C.c_void_p]),
    "hg_conv_small_wgrad_workspace_bytes": (C.c_size_t, [C.c_int32] * 3),
    "hg_conv_small_wgrad": (C.c_int, [C.c_void_p] * 4 + [C.c_size_t] + [C.c_int32] * 7 + [C.c_int64] * 4 +
                            [C.c_void_p]),
    "hg_pad_round_nhwc": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 5 + [C.c_int64] * 4 + [C.c_void_p]),
    "hg_upsample2x_planar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_void_p]),
    "hg_diffgrad_step": (C.c_int, [C.c_int32] + [C.c_void_p] * 7 + [C.c_float] * 5 + [C.c_void_p]),
    "hg_ema_update": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "hg_bias_act_bwd": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_float, C.c_void_p]),
    "hg_pack_conv_weight": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p]),
    "hg_instnorm_lrelu_fwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_float, C.c_float,
                                                                         C.c_int32, C.c_void_p]),
    "hg_instnorm_lrelu_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 3 + [C.c_float, C.c_float,
                                                                         C.c_int32, C.c_void_p]),
    "hg_laplacian_l1_workspace_bytes": (C.c_size_t, []),
    "hg_laplacian_l1_fwd": (C.c_int, [C.c_void_p] * 5 + [C.c_size_t] + [C.c_int32] * 3 + [C.c_void_p]),
    "hg_laplacian_l1_bwd": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p]),
    "hg_depthwise_conv": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 6 + [C.c_void_p]),
    "hg_hellinger_workspace_bytes": (C.c_size_t, []),
    "hg_hellinger_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hg_hellinger_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class HistoganLibraryError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load (building first if nvcc is available and the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        try:
            from . import build as _build
            _build.build()
        except Exception as e:  # no silent fallback
            raise HistoganLibraryError(
                f"{LIB_PATH} is missing and could not be built ({e}); run "
                f"`python -m histogan_b200.build`. There is no CPU fallback.") from e
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:
        raise HistoganLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.hg_abi_version() != HG_ABI_VERSION:
        raise HistoganLibraryError(
            f"ABI mismatch: library {lib.hg_abi_version()} vs binding {HG_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().hg_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): "
                           f"{msg.decode() if msg else 'unknown error'}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def current_stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(
            f"{what}: expected a CUDA tensor, got device {t.device}. histogan_b200 has no "
            f"CPU path (the CPU restatement lives in oracle/ and is test-only).")
