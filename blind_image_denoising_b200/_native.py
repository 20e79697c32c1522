"""ctypes binding of libbfcnn_b200.so (include/bfcnn_b200.h).

There is deliberately no Python/CPU implementation behind these calls: if the shared
library is missing, or no B200 is visible, every compute entry point raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint32,
                    c_uint64, c_void_p)
from pathlib import Path

from .arch import CArch

_LIB_PATH = Path(__file__).resolve().parent / "libbfcnn_b200.so"

PREC_FP32, PREC_F16, PREC_F16X3 = 0, 1, 2
# no "bf16" alias: the tensor-core arm computes with fp16 operands (10 mantissa bits, not bf16's 7)
PRECISIONS = {"fp32": PREC_FP32, "f16": PREC_F16, "fp16": PREC_F16, "f16x3": PREC_F16X3, "fp16x3": PREC_F16X3}
FLAG_IN_DEVICE, FLAG_OUT_DEVICE, FLAG_NO_PAD_POW2, FLAG_IN_F32 = 1, 2, 4, 8

# every symbol include/bfcnn_b200.h declares (tests check the library exports all of them)
EXPORTED_SYMBOLS = [
    "bfcnn_abi_version", "bfcnn_last_error", "bfcnn_device_count", "bfcnn_num_weights",
    "bfcnn_num_trainable", "bfcnn_create", "bfcnn_destroy", "bfcnn_release_workspaces", "bfcnn_set_weights",
    "bfcnn_get_weights", "bfcnn_denoise_u8", "bfcnn_denoise_f32", "bfcnn_launch_count",
    "bfcnn_last_stack_ms", "bfcnn_set_kernel_timing", "bfcnn_kernel_times", "bfcnn_corrupt", "bfcnn_loss", "bfcnn_train_step", "bfcnn_train_losses", "bfcnn_saved_activation", "bfcnn_downscale2x",
    "bfcnn_allreduce_grads", "bfcnn_adam_step", "bfcnn_conv3x3", "bfcnn_set_train_engine",
    "bfcnn_generic_prepare", "bfcnn_generic_conv2d", "bfcnn_generic_finish",
]


class NoiseCfg(ctypes.Structure):
    _fields_ = [("additive_min", c_float), ("additive_max", c_float),
                ("multiplicative_min", c_float), ("multiplicative_max", c_float),
                ("random_left_right", c_int32), ("random_up_down", c_int32),
                ("subsample", c_int32), ("round_values", c_int32), ("draw_group", c_int32)]


class LossCfg(ctypes.Structure):
    _fields_ = [("hinge", c_float), ("cutoff", c_float), ("mae_multiplier", c_float),
                ("mse_multiplier", c_float), ("regularization", c_float), ("ssim_multiplier", c_float)]


class AdamCfg(ctypes.Structure):
    _fields_ = [("learning_rate", c_float), ("beta_1", c_float), ("beta_2", c_float),
                ("epsilon", c_float), ("global_clipnorm", c_float)]


class NativeError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libbfcnn_b200 error {status}: {message}")
        self.status = status


_lib = None


def library_path() -> Path:
    return _LIB_PATH


def load_library() -> ctypes.CDLL:
    """dlopen the in-tree library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} is missing: build it with `python -m blind_image_denoising_b200.build` "
            "(needs nvcc). This package has no CPU or pure-Python fallback.")
    lib = ctypes.CDLL(str(_LIB_PATH), mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
    H = c_void_p
    lib.bfcnn_abi_version.restype = c_int
    lib.bfcnn_last_error.restype = c_char_p
    lib.bfcnn_device_count.restype = c_int
    lib.bfcnn_num_weights.argtypes = [POINTER(CArch)]
    lib.bfcnn_num_weights.restype = c_int64
    lib.bfcnn_num_trainable.argtypes = [POINTER(CArch)]
    lib.bfcnn_num_trainable.restype = c_int64
    lib.bfcnn_create.argtypes = [POINTER(CArch), c_void_p, c_size_t, c_int, POINTER(H)]
    lib.bfcnn_create.restype = c_int
    lib.bfcnn_destroy.argtypes = [H]
    lib.bfcnn_destroy.restype = None
    lib.bfcnn_release_workspaces.argtypes = [H]
    lib.bfcnn_release_workspaces.restype = c_int
    lib.bfcnn_set_weights.argtypes = [H, c_void_p, c_size_t]
    lib.bfcnn_set_weights.restype = c_int
    lib.bfcnn_get_weights.argtypes = [H, c_void_p, c_size_t]
    lib.bfcnn_get_weights.restype = c_int
    for f in (lib.bfcnn_denoise_u8, lib.bfcnn_denoise_f32):
        f.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_uint32, c_void_p]
        f.restype = c_int
    lib.bfcnn_launch_count.argtypes = [H]
    lib.bfcnn_launch_count.restype = c_int64
    lib.bfcnn_last_stack_ms.argtypes = [H, POINTER(c_float)]
    lib.bfcnn_last_stack_ms.restype = c_int
    lib.bfcnn_set_kernel_timing.argtypes = [H, c_int]
    lib.bfcnn_set_kernel_timing.restype = c_int
    lib.bfcnn_kernel_times.argtypes = [H, POINTER(c_float), POINTER(c_int), c_int, POINTER(c_int)]
    lib.bfcnn_kernel_times.restype = c_int
    lib.bfcnn_corrupt.argtypes = [H, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_uint64,
                                  c_uint64, POINTER(NoiseCfg), c_void_p]
    lib.bfcnn_corrupt.restype = c_int
    lib.bfcnn_loss.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(LossCfg),
                               POINTER(c_float), c_void_p]
    lib.bfcnn_loss.restype = c_int
    lib.bfcnn_train_step.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_int, POINTER(LossCfg),
                                     c_void_p, c_void_p, c_int, c_void_p]
    lib.bfcnn_train_step.restype = c_int
    lib.bfcnn_downscale2x.argtypes = [H, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.bfcnn_downscale2x.restype = c_int
    lib.bfcnn_train_losses.argtypes = [H, POINTER(c_float), c_void_p]
    lib.bfcnn_train_losses.restype = c_int
    lib.bfcnn_saved_activation.argtypes = [H, c_int, c_int, c_void_p, c_void_p]
    lib.bfcnn_saved_activation.restype = c_int
    lib.bfcnn_set_train_engine.argtypes = [H, c_int]
    lib.bfcnn_set_train_engine.restype = c_int
    lib.bfcnn_conv3x3.argtypes = [H, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.bfcnn_conv3x3.restype = c_int
    lib.bfcnn_allreduce_grads.argtypes = [H, c_void_p, c_void_p, c_void_p]
    lib.bfcnn_allreduce_grads.restype = c_int
    lib.bfcnn_generic_prepare.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.bfcnn_generic_prepare.restype = c_int
    lib.bfcnn_generic_conv2d.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                         c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.bfcnn_generic_conv2d.restype = c_int
    lib.bfcnn_generic_finish.argtypes = [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]
    lib.bfcnn_generic_finish.restype = c_int
    lib.bfcnn_adam_step.argtypes = [H, c_void_p, c_float, POINTER(AdamCfg), c_int64, c_void_p]
    lib.bfcnn_adam_step.restype = c_int
    if lib.bfcnn_abi_version() != 3:
        raise ImportError("libbfcnn_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load_library().bfcnn_last_error()
        raise NativeError(status, msg.decode("utf-8", "replace") if msg else "")
