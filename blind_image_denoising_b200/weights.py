"""Weight containers for the resnet denoiser hot path.

The unit that crosses the C ABI is ONE flat float32 vector holding
``hydra.variables`` in Keras order: base kernel [k0,k0,3,16] (HWIO),
then per block W_a, W_b [3,3,16,16], gamma, moving_mean, moving_var [16], then
the head kernels [1,1,16,F] and [1,1,F,3].  BN folding, head collapse and the
packing into kernel layouts happen inside the native library (csrc/host_pack.cuh),
so that set_weights / get_weights speak the reference's own format.

The three pretrained resnet directories named by the reference README are not in
the reference snapshot, so `synthetic_variables` produces deterministic
stand-ins of the same shapes; they are used for parity tests and benchmarks and
are labelled as synthetic wherever they are written to disk.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .arch import Arch

GLOROT_TRUNC_FACTOR = 0.87962566103423978  # keras VarianceScaling truncated-normal correction


def _glorot_truncated_normal(rng: np.random.Generator, shape) -> np.ndarray:
    """Keras ``glorot_normal``: truncated normal (+-2 sigma) with
    std = sqrt(2/(fan_in+fan_out))/0.8796 (reference default initializer,
    `bfcnn/backbone_resnet.py:36`)."""
    receptive = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * receptive, shape[-1] * receptive
    std = np.sqrt(2.0 / (fan_in + fan_out)) / GLOROT_TRUNC_FACTOR
    out = rng.standard_normal(size=shape)
    bad = np.abs(out) >= 2.0
    while bad.any():
        out[bad] = rng.standard_normal(size=int(bad.sum()))
        bad = np.abs(out) >= 2.0
    return (out * std).astype(np.float32)


def synthetic_variables(arch: Arch, seed: int = 0,
                        residual_gain: float = 0.5,
                        head_gain: float = 0.8) -> List[np.ndarray]:
    """Deterministic synthetic ``hydra.variables`` (SURVEY 8d "Synthetic weights").

    gamma ~ U(0.5,1.5)*residual_gain, moving_mean ~ N(0,0.05^2), moving_var ~ U(0.5,1.5)
    so that BN folding AND its constant term are exercised.  The two
    gains keep the 2N-conv residual stream and the pre-tanh head output inside the
    un-saturated range of ``tanh(2y)*0.51`` for N up to 18; with gain 1 every output
    pixel clips to 0/255 and a parity test could not see an error.
    """
    rng = np.random.default_rng(seed)
    shapes = arch.variable_shapes()
    out: List[np.ndarray] = []
    out.append(_glorot_truncated_normal(rng, shapes[0]))
    for _ in range(arch.no_layers):
        out.append(_glorot_truncated_normal(rng, (3, 3, arch.filters, arch.filters)))
        out.append(_glorot_truncated_normal(rng, (3, 3, arch.filters, arch.filters)))
        out.append((rng.uniform(0.5, 1.5, arch.filters) * residual_gain).astype(np.float32))
        out.append((rng.standard_normal(arch.filters) * 0.05).astype(np.float32))
        out.append(rng.uniform(0.5, 1.5, arch.filters).astype(np.float32))
    out.append(_glorot_truncated_normal(rng, shapes[-2]))
    out.append((_glorot_truncated_normal(rng, shapes[-1]) * head_gain).astype(np.float32))
    return out


def initial_variables(arch: Arch, seed: int = 0) -> List[np.ndarray]:
    """``hydra.variables`` as Keras initialises them (what `model_builder` yields before any training):
    glorot_normal kernels (backbone_resnet.py:36, model.py:276), BN gamma = 1, moving_mean = 0, moving_var = 1."""
    rng = np.random.default_rng(seed)
    shapes = arch.variable_shapes()
    out: List[np.ndarray] = [_glorot_truncated_normal(rng, shapes[0])]
    for _ in range(arch.no_layers):
        out.append(_glorot_truncated_normal(rng, (3, 3, arch.filters, arch.filters)))
        out.append(_glorot_truncated_normal(rng, (3, 3, arch.filters, arch.filters)))
        out.append(np.ones(arch.filters, np.float32))
        out.append(np.zeros(arch.filters, np.float32))
        out.append(np.ones(arch.filters, np.float32))
    out.append(_glorot_truncated_normal(rng, shapes[-2]))
    out.append(_glorot_truncated_normal(rng, shapes[-1]))
    return out


def flatten_variables(arch: Arch, variables: Sequence[np.ndarray]) -> np.ndarray:
    shapes = arch.variable_shapes()
    if len(variables) != len(shapes):
        raise ValueError(f"expected {len(shapes)} variables, got {len(variables)}")
    parts = []
    for v, s in zip(variables, shapes):
        v = np.asarray(v, dtype=np.float32)
        if tuple(v.shape) != tuple(s):
            raise ValueError(f"variable shape {v.shape} != expected {s}")
        parts.append(np.ascontiguousarray(v).reshape(-1))
    return np.concatenate(parts).astype(np.float32)


def unflatten_variables(arch: Arch, flat: np.ndarray) -> List[np.ndarray]:
    flat = np.asarray(flat, dtype=np.float32).reshape(-1)
    if flat.size != arch.num_weights():
        raise ValueError(f"flat size {flat.size} != {arch.num_weights()}")
    out, o = [], 0
    for s in arch.variable_shapes():
        n = int(np.prod(s))
        out.append(flat[o:o + n].reshape(s).copy())
        o += n
    return out


def trainable_offsets(arch: Arch):
    """[(offset_in_flat_variables, size, offset_in_flat_trainables)] for every
    trainable variable, in Keras ``trainable_variables`` order."""
    res, o, t = [], 0, 0
    for s, tr in zip(arch.variable_shapes(), arch.trainable_mask()):
        n = int(np.prod(s))
        if tr:
            res.append((o, n, t))
            t += n
        o += n
    return res


def gather_trainables(arch: Arch, flat: np.ndarray) -> np.ndarray:
    flat = np.asarray(flat).reshape(-1)
    return np.concatenate([flat[o:o + n] for o, n, _ in trainable_offsets(arch)])
