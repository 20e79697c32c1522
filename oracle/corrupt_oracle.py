"""CPU ORACLE for the training-input corruption -- TEST INFRASTRUCTURE ONLY (see bfcnn_oracle.py).

Restates `dataset_builder.prepare_data_fn` (/root/reference/bfcnn/dataset.py:120-238):
  geometric_augmentation_fn :141-158  flip left/right and up/down, each w.p. 1/2
  noise_augmentation_fn     :161-230  use_add, use_mul ~ Bernoulli(1/2) (:170-177), sigma ~ U(min,max)
                                      (:178-187), FIRST x*TN(1,sigma_mul) (:190-206), THEN x+TN(0,sigma_add)
                                      (:209-225), tf.round (:228)
  clean batch               :233-235  round, cast float32
plus the sub-sampling corruption the README lists (README.md:49-55) but HEAD does not
implement (our spec: w.p. 1/2 each 2x2 cell takes its top-left pixel).

PARITY UNPINNED with respect to TensorFlow's own random stream: `tf.random.*` cannot run here
 and its op-seed bookkeeping is not reproducible outside TF.  What is restated
exactly is the *algorithm* of tf.random.truncated_normal / tf.random.uniform as published
(Philox4x32-10, 23-bit uniforms, Box-Muller with u1 clamped to 1e-7, rejection of |z| >= 2);
the stream layout (which counter feeds which value) is ours and is documented in DESIGN.md.
Distribution tests in tests/test_oracle.py pin the sampler to scipy.stats.truncnorm.

This file is an independent numpy restatement: it shares no code with the CUDA kernel.  Every
float32 operation below is a single IEEE round-to-nearest op (numpy float32 arithmetic), the
kernel uses the matching __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn, hence bit-identical output.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

F32 = np.float32
U32 = np.uint32
U64 = np.uint64


@dataclass
class NoiseConfig:
    additive_min: float = 5.0          # configs/resnet_color_1x6_...json:70 "additional_noise"
    additive_max: float = 40.0
    multiplicative_min: float = 0.05   # configs/...json:71 "multiplicative_noise"
    multiplicative_max: float = 0.1
    random_left_right: bool = True
    random_up_down: bool = True
    subsample: bool = False
    round_values: bool = True          # dataset.py:228 always rounds; False exists for tests of the un-rounded values
    draw_group: int = 0                # who shares the call-level draws of dataset.py:141-187: 0/1 every sample its own,
                                       # k > 1 runs of k consecutive global sample indices (the crops of one image,
                                       # dataset.py:276-297), < 0 the whole call


# ----------------------------------------------------------------------------------
def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Philox4x32-10 (Salmon et al. 2011; the generator behind tf.random.*). Vectorised over counters."""
    c0 = np.asarray(c0, U32).astype(U64); c1 = np.asarray(c1, U32).astype(U64)
    c2 = np.asarray(c2, U32).astype(U64); c3 = np.asarray(c3, U32).astype(U64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF; k1 = int(k1) & 0xFFFFFFFF
    M0, M1 = U64(0xD2511F53), U64(0xCD9E8D57)
    mask = U64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> U64(32), p0 & mask
        hi1, lo1 = p1 >> U64(32), p1 & mask
        n0 = hi1 ^ c1 ^ U64(k0)
        n2 = hi0 ^ c3 ^ U64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0.astype(U32), c1.astype(U32), c2.astype(U32), c3.astype(U32)


def u01(r):
    """tf.random.uniform float32: 23 mantissa bits -> [0, 1)."""
    return (np.asarray(r, U32) >> U32(9)).astype(F32) * F32(2.0 ** -23)


def ln_det(u):
    u = np.asarray(u, F32)
    b = u.view(U32)
    e = (b >> U32(23)).astype(np.int32) - 127
    m = ((b & U32(0x007FFFFF)) | U32(0x3F800000)).view(F32)
    big = m > F32(1.41421354)
    m = np.where(big, m * F32(0.5), m).astype(F32)
    e = np.where(big, e + 1, e)
    s = (m + F32(-1.0)) / (m + F32(1.0))
    s2 = s * s
    p = np.full_like(s, F32(0.222222224))
    p = p * s2 + F32(0.285714298)
    p = p * s2 + F32(0.400000006)
    p = p * s2 + F32(0.666666687)
    p = p * s2 + F32(2.0)
    return e.astype(F32) * F32(0.693147182) + s * p


def sincos_turns_det(v):
    v = np.asarray(v, F32)
    q = (v * F32(4.0) + F32(0.5)).astype(np.int32)
    f = v + (-(q.astype(F32) * F32(0.25)))
    a = f * F32(6.28318548)
    a2 = a * a
    ps = np.full_like(a, F32(2.75573188e-06))
    ps = ps * a2 + F32(-1.98412701e-04)
    ps = ps * a2 + F32(8.33333377e-03)
    ps = ps * a2 + F32(-1.66666672e-01)
    ps = ps * a2 + F32(1.0)
    s0 = a * ps
    pc = np.full_like(a, F32(-2.75573200e-07))
    pc = pc * a2 + F32(2.48015876e-05)
    pc = pc * a2 + F32(-1.38888892e-03)
    pc = pc * a2 + F32(4.16666679e-02)
    pc = pc * a2 + F32(-0.5)
    pc = pc * a2 + F32(1.0)
    k = q & 3
    sn = np.select([k == 0, k == 1, k == 2], [s0, pc, -s0], -pc).astype(F32)
    cs = np.select([k == 0, k == 1, k == 2], [pc, -s0, -pc], s0).astype(F32)
    return sn, cs


def box_muller_det(r0, r1):
    """TF BoxMullerFloat: u1 clamped to 1e-7; z0 = sin(2 pi u2) r, z1 = cos(2 pi u2) r."""
    u1 = np.maximum(u01(r0), F32(1.0e-7)).astype(F32)
    rad = np.sqrt(F32(-2.0) * ln_det(u1)).astype(F32)
    sn, cs = sincos_turns_det(u01(r1))
    return sn * rad, cs * rad


def truncated_normal_det(pix, slot: int, g: int, seed: int):
    """Standard normal truncated to (-2, 2) by rejection: value slot `slot` of pixels `pix` of global
    sample g; attempt a uses Philox counter (pix, slot | a<<8, g_lo, g_hi), key (seed_lo, seed_hi)."""
    pix = np.asarray(pix, U32)
    out = np.zeros(pix.shape, F32)
    todo = np.ones(pix.shape, bool)
    g_lo, g_hi = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a = 0
    while todo.any():
        idx = np.nonzero(todo)
        r0, r1, r2, r3 = philox4x32_10(pix[idx], U32(slot | (a << 8)), U32(g_lo), U32(g_hi), k0, k1)
        z0, z1 = box_muller_det(r0, r1)
        z2, z3 = box_muller_det(r2, r3)
        val = np.zeros(z0.shape, F32)
        found = np.zeros(z0.shape, bool)
        for z in (z0, z1, z2, z3):
            ok = (~found) & (np.abs(z) < F32(2.0))
            val[ok] = z[ok]
            found |= ok
        sub = tuple(i[found] for i in idx)
        out[sub] = val[found]
        todo[sub] = False
        a += 1
    return out


def sample_parameters(g: int, seed: int, cfg: NoiseConfig):
    """The per-sample draws of dataset.py:141-142,170-187."""
    g_lo, g_hi = g & 0xFFFFFFFF, (g >> 32) & 0xFFFFFFFF
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    ra = [int(x) for x in philox4x32_10(U32(0), U32(0xFFFFFFFF), U32(g_lo), U32(g_hi), k0, k1)]
    rb = [int(x) for x in philox4x32_10(U32(1), U32(0xFFFFFFFF), U32(g_lo), U32(g_hi), k0, k1)]
    add_on, mul_on = cfg.additive_max > 0, cfg.multiplicative_max > 0
    amin, amax = F32(cfg.additive_min), F32(cfg.additive_max)
    mmin, mmax = F32(cfg.multiplicative_min), F32(cfg.multiplicative_max)
    return {
        "flip_lr": bool(cfg.random_left_right and u01(ra[0]) > F32(0.5)),
        "flip_ud": bool(cfg.random_up_down and u01(ra[1]) > F32(0.5)),
        "use_add": bool(add_on and u01(ra[2]) > F32(0.5)),
        "use_mul": bool(mul_on and u01(ra[3]) > F32(0.5)),
        "sigma_add": F32(amin + F32(F32(amax + (-amin)) * u01(rb[0]))),
        "sigma_mul": F32(mmin + F32(F32(mmax + (-mmin)) * u01(rb[1]))),
        "subsample": bool(cfg.subsample and u01(rb[2]) > F32(0.5)),
    }


def corrupt(clean_u8: np.ndarray, seed: int, sample_offset: int, cfg: NoiseConfig) -> Tuple[np.ndarray, np.ndarray]:
    """clean uint8 [n,h,w,3] -> (clean float32, noisy float32), sample s on stream (seed, sample_offset + s)."""
    clean_u8 = np.asarray(clean_u8)
    assert clean_u8.dtype == np.uint8 and clean_u8.ndim == 4 and clean_u8.shape[-1] == 3
    n, h, w, _ = clean_u8.shape
    clean = np.empty((n, h, w, 3), F32)
    noisy = np.empty((n, h, w, 3), F32)
    pix = np.arange(h * w, dtype=np.uint32).reshape(h, w)
    for s in range(n):
        g = sample_offset + s
        gd = g
        if cfg.draw_group > 1:
            gd = g - g % cfg.draw_group
        elif cfg.draw_group < 0:
            gd = sample_offset
        p = sample_parameters(gd, seed, cfg)
        img = clean_u8[s]
        if p["flip_lr"]:
            img = img[:, ::-1]
        if p["flip_ud"]:
            img = img[::-1]
        c = img.astype(F32)                       # dataset.py:233-235
        v = c
        if p["subsample"]:
            v = c[(np.arange(h) & ~1)][:, (np.arange(w) & ~1)]
        v = v.copy()
        if p["use_mul"]:                          # dataset.py:190-206
            for ch in range(3):
                z = truncated_normal_det(pix, ch, g, seed)
                v[..., ch] = v[..., ch] * (F32(1.0) + p["sigma_mul"] * z)
        if p["use_add"]:                          # dataset.py:209-225
            for ch in range(3):
                z = truncated_normal_det(pix, 3 + ch, g, seed)
                v[..., ch] = v[..., ch] + p["sigma_add"] * z
        if cfg.round_values:
            v = np.rint(v)                        # tf.round, half to even (dataset.py:228)
        clean[s], noisy[s] = c, v.astype(F32)
    return clean, noisy


# ----------------------------------------------------------------------------------
# Adam + global-norm clip (optimizer.py:145-224 -> keras.optimizers.Adam(global_clipnorm=...))
# ----------------------------------------------------------------------------------
def adam_step(w, g, m, v, step: int, *, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7,
              global_clipnorm=0.0, grad_scale=1.0):
    """One Keras-2.13 Adam update in float64 on flat vectors; returns (w, m, v).
    tf.clip_by_global_norm: g * clip / max(norm, clip)."""
    w = np.asarray(w, np.float64); g = np.asarray(g, np.float64) * grad_scale
    m = np.asarray(m, np.float64); v = np.asarray(v, np.float64)
    if global_clipnorm and global_clipnorm > 0:
        norm = np.sqrt((g * g).sum())
        g = g * (global_clipnorm / max(norm, global_clipnorm))
    alpha = learning_rate * np.sqrt(1.0 - beta_2 ** step) / (1.0 - beta_1 ** step)
    m = m + (g - m) * (1.0 - beta_1)
    v = v + (g * g - v) * (1.0 - beta_2)
    w = w - m * alpha / (np.sqrt(v) + epsilon)
    return w, m, v
