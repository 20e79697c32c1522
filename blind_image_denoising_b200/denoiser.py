"""Host-side mirror of the reference's inference callable.

`Denoiser.__call__(uint8[N,H,W,3]) -> uint8[N,H,W,3]` has the signature and semantics of
`DenoiserModule.__call__` (reference bfcnn/module_denoiser.py:39-75), which is what
`bfcnn.load_model(name)` returns in the reference (bfcnn/__init__.py:81-97).  All arithmetic
runs in libbfcnn_b200.so on a B200; this file only moves pointers.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _native
from .arch import Arch
from .weights import flatten_variables, unflatten_variables


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class Denoiser:
    """One model instance bound to one GPU (one handle per device)."""

    def __init__(self, arch: Arch, variables: Sequence[np.ndarray], *, device: int = 0,
                 precision: str = "f16x3", pad_pow2: bool = True, name: str = ""):
        self._lib = _native.load_library()
        self.arch = arch
        self.name = name
        self.device = int(device)
        self.precision = precision
        self.pad_pow2 = bool(pad_pow2)
        if precision != "auto" and precision not in _native.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_native.PRECISIONS) + ['auto']}")
        flat = np.ascontiguousarray(flatten_variables(arch, variables), dtype=np.float32)
        carch = arch.to_c()
        h = ctypes.c_void_p()
        _native.check(self._lib.bfcnn_create(ctypes.byref(carch), flat.ctypes.data, flat.size,
                                             self.device, ctypes.byref(h)))
        self._h = h
        self.calibration = None
        if precision == "auto":
            self.precision = self._calibrate()

    def _calibrate(self, size: int = 192, max_abs: float = 0.25, mean_abs: float = 0.025) -> str:
        """precision="auto": the fast f16 arithmetic when, with THESE weights, it stays within half the fp32 gate (max-abs
        0.5 / mean-abs 0.05 on the 0-255 scale) of the fp32-grade f16x3 arithmetic on two probe images -- uniform
        noise and a smooth ramp with noise --, else f16x3.  Trained denoisers pass (f16 is off by ~0.3 / 0.03 at 1x18,
        tests/test_pretrained_gpu.py); glorot-scale random weights do not (0.86 / 0.07).  A heuristic on probe inputs, not a
        bound: the default of load_model stays f16x3."""
        rng = np.random.default_rng(0)
        yy, xx = np.mgrid[0:size, 0:size]
        ramp = np.stack([(yy + xx) * 255.0 / (2 * size - 2), yy * 255.0 / (size - 1), xx * 255.0 / (size - 1)], -1)
        probes = np.stack([rng.integers(0, 256, size=(size, size, 3)).astype(np.float64),
                           np.clip(ramp + rng.normal(0.0, 20.0, ramp.shape), 0, 255)]).round().astype(np.uint8)
        a = self(probes, precision="f16", return_float=True).astype(np.float64)
        b = self(probes, precision="f16x3", return_float=True).astype(np.float64)
        d = np.abs(a - b)
        ok = bool(d.max() <= max_abs and d.mean() <= mean_abs)
        self.calibration = {"max_abs": float(d.max()), "mean_abs": float(d.mean()), "chosen": "f16" if ok else "f16x3"}
        return self.calibration["chosen"]

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bfcnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def get_weights(self):
        flat = np.empty(self.arch.num_weights(), dtype=np.float32)
        _native.check(self._lib.bfcnn_get_weights(self._h, flat.ctypes.data, flat.size))
        return unflatten_variables(self.arch, flat)

    def set_weights(self, variables: Sequence[np.ndarray]):
        flat = np.ascontiguousarray(flatten_variables(self.arch, variables), dtype=np.float32)
        _native.check(self._lib.bfcnn_set_weights(self._h, flat.ctypes.data, flat.size))

    def launch_count(self) -> int:
        return int(self._lib.bfcnn_launch_count(self._h))

    def release_workspaces(self):
        """Give the feature-map / staging workspaces back to the driver (they grow to the largest call seen)."""
        _native.check(self._lib.bfcnn_release_workspaces(self._h))

    def last_stack_ms(self) -> float:
        ms = ctypes.c_float()
        _native.check(self._lib.bfcnn_last_stack_ms(self._h, ctypes.byref(ms)))
        return float(ms.value)

    def set_kernel_timing(self, on: bool):
        """Record CUDA events around every conv-stack launch of the following calls (bench.py's per-kernel roofline)."""
        _native.check(self._lib.bfcnn_set_kernel_timing(self._h, int(bool(on))))

    def kernel_times(self):
        """[(kind, ms)] of the last timed call; kind 0 = base conv, 1 = pass, 2 = last pass, -1 = idle gap between two
        launches."""
        ms = (ctypes.c_float * 128)()
        kinds = (ctypes.c_int * 128)()
        cnt = ctypes.c_int()
        _native.check(self._lib.bfcnn_kernel_times(self._h, ms, kinds, 128, ctypes.byref(cnt)))
        return [(int(kinds[i]), float(ms[i])) for i in range(cnt.value)]

    # ------------------------------------------------------------------
    def __call__(self, image, *, precision: Optional[str] = None, pad_pow2: Optional[bool] = None,
                 return_float: bool = False, out=None):
        """uint8 [N,H,W,3] (numpy, torch-CPU or torch-CUDA) -> same container type.

        return_float=True returns the float32 prediction before round/cast (0..255)."""
        prec = _native.PRECISIONS[precision or self.precision]
        flags = 0 if (self.pad_pow2 if pad_pow2 is None else pad_pow2) else _native.FLAG_NO_PAD_POW2
        fn = self._lib.bfcnn_denoise_f32 if return_float else self._lib.bfcnn_denoise_u8

        if _is_torch(image):
            import torch
            if image.dtype != torch.uint8:
                raise TypeError("image must be uint8")  # tf.TensorSpec(dtype=tf.uint8), module_denoiser.py:44
            if image.dim() != 4 or image.shape[-1] != 3:
                raise ValueError("image must have shape [N,H,W,3]")
            x = image.contiguous()
            n, h, w, _ = x.shape
            odt = torch.float32 if return_float else torch.uint8
            if out is None:
                out = torch.empty((n, h, w, 3), dtype=odt, device=x.device)
            stream = None
            if x.is_cuda:
                if x.device.index != self.device:
                    raise ValueError(f"tensor is on cuda:{x.device.index}, model on cuda:{self.device}")
                flags |= _native.FLAG_IN_DEVICE | _native.FLAG_OUT_DEVICE
                stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
            _native.check(fn(self._h, x.data_ptr(), out.data_ptr(), n, h, w, prec, flags, stream))
            return out

        x = np.asarray(image)
        if x.dtype != np.uint8:
            raise TypeError("image must be uint8")
        if x.ndim != 4 or x.shape[-1] != 3:
            raise ValueError("image must have shape [N,H,W,3]")
        x = np.ascontiguousarray(x)
        n, h, w, _ = x.shape
        if out is None:
            out = np.empty((n, h, w, 3), dtype=np.float32 if return_float else np.uint8)
        _native.check(fn(self._h, x.ctypes.data, out.ctypes.data, n, h, w, prec, flags, None))
        return out


class PipelinedDenoiser:
    """`depth` model instances on ONE GPU fed from a small thread pool.

    A `Denoiser` call on host arrays returns only after its result has been copied back, so a single instance leaves
    the GPU idle while the first chunk goes up and the last one comes down.  ctypes releases the GIL and every handle
    owns its streams and workspaces, so with two instances the copies of one batch overlap the conv stack of the
    other.  `map` keeps the order of its inputs; `__call__` is the plain synchronous call on instance 0."""

    def __init__(self, factory, depth: int = 2):
        from concurrent.futures import ThreadPoolExecutor
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.models = [factory() for _ in range(depth)]
        # one worker thread per instance: a handle is not re-entrant, calls on it stay in order
        self._pools = [ThreadPoolExecutor(max_workers=1) for _ in range(depth)]

    def __call__(self, image, **kwargs):
        return self.models[0](image, **kwargs)

    def map(self, images, outs=None, **kwargs):
        """Denoise a sequence of host batches; yields the results in order (outs: optional matching output buffers)."""
        futures = []
        for i, x in enumerate(images):
            k = i % len(self.models)
            kw = dict(kwargs)
            if outs is not None:
                kw["out"] = outs[i]
            futures.append(self._pools[k].submit(self.models[k], x, **kw))
            if len(futures) >= 2 * len(self.models):   # bounded look-ahead
                yield futures.pop(0).result()
        for f in futures:
            yield f.result()

    def close(self):
        for pl in self._pools:
            pl.shutdown(wait=True)
        for m in self.models:
            m.close()
