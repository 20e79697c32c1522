"""CPU tests of the corruption / optimiser oracles (no GPU): known-answer vectors of the generator,
distribution of the truncated-normal sampler, semantics of dataset.py:120-238, committed golden vectors."""
import os

import numpy as np
import pytest

from oracle import corrupt_oracle as C

U32 = np.uint32
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "training_golden.npz")


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10 (the generator tf.random.* is built on)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = C.philox4x32_10(*(U32(c) for c in ctr), key[0], key[1])
        assert tuple(int(g) for g in got) == want


def test_uniform_23_bits():
    assert C.u01(0) == 0.0 and C.u01(0xFFFFFFFF) == np.float32(1.0 - 2.0 ** -23)
    assert C.u01(1 << 31) == 0.5


def test_elementary_functions_accuracy():
    u = np.maximum(np.random.default_rng(0).random(200000).astype(np.float32), 1e-7)
    assert np.abs(C.ln_det(u) - np.log(u.astype(np.float64))).max() < 2e-6
    s, c = C.sincos_turns_det(u)
    assert np.abs(s - np.sin(2 * np.pi * u.astype(np.float64))).max() < 5e-7
    assert np.abs(c - np.cos(2 * np.pi * u.astype(np.float64))).max() < 5e-7


def test_truncated_normal_distribution():
    from scipy import stats
    z = C.truncated_normal_det(np.arange(300000, dtype=np.uint32), 2, 11, 99)
    assert np.abs(z).max() < 2.0                                # tf.random.truncated_normal: |z| < 2 sigma
    assert abs(float(z.mean())) < 5e-3
    assert abs(float(z.std()) - 0.87962566) < 5e-3              # std of N(0,1) truncated to +-2
    assert stats.kstest(z.astype(np.float64), stats.truncnorm(-2, 2).cdf).pvalue > 1e-3
    # independent slots / samples are uncorrelated
    z2 = C.truncated_normal_det(np.arange(300000, dtype=np.uint32), 3, 11, 99)
    assert abs(float(np.corrcoef(z, z2)[0, 1])) < 1e-2


def test_corrupt_semantics():
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(16, 20, 24, 3), dtype=np.uint8)
    cfg = C.NoiseConfig(subsample=True)
    clean, noisy = C.corrupt(x, seed=5, sample_offset=100, cfg=cfg)
    assert clean.dtype == np.float32 and noisy.dtype == np.float32
    assert clean.min() >= 0 and clean.max() <= 255               # reference tests/bfcnn/test_dataset.py
    assert np.array_equal(noisy, np.rint(noisy))                 # dataset.py:228
    seen = set()
    for s in range(16):
        p = C.sample_parameters(100 + s, 5, cfg)
        seen.add((p["flip_lr"], p["flip_ud"], p["use_add"], p["use_mul"]))
        ref = x[s]
        if p["flip_lr"]:
            ref = ref[:, ::-1]
        if p["flip_ud"]:
            ref = ref[::-1]
        assert np.array_equal(clean[s], ref.astype(np.float32))  # flips are applied to the clean target too
        assert 5.0 <= p["sigma_add"] <= 40.0 and 0.05 <= p["sigma_mul"] <= 0.1
        if not (p["use_add"] or p["use_mul"] or p["subsample"]):
            assert np.array_equal(noisy[s], clean[s])
        if p["use_add"] and not p["use_mul"] and not p["subsample"]:
            d = noisy[s] - clean[s]
            assert np.abs(d).max() <= 2 * p["sigma_add"] + 0.51   # truncated at 2 sigma, then rounded
            assert abs(d.std() / p["sigma_add"] - 0.8796) < 0.08
    assert len(seen) > 4                                          # the Bernoulli(1/2) switches all fire
    # stream position is (seed, global sample index): a shifted batch reproduces the same samples
    c2, n2 = C.corrupt(x[4:], seed=5, sample_offset=104, cfg=cfg)
    assert np.array_equal(n2, noisy[4:]) and np.array_equal(c2, clean[4:])
    _, n3 = C.corrupt(x, seed=6, sample_offset=100, cfg=cfg)
    assert not np.array_equal(n3, noisy)


def test_corrupt_disabled_noise_is_identity():
    x = np.random.default_rng(0).integers(0, 256, size=(3, 8, 8, 3), dtype=np.uint8)
    cfg = C.NoiseConfig(additive_min=0, additive_max=0, multiplicative_min=0, multiplicative_max=0,
                        random_left_right=False, random_up_down=False)
    clean, noisy = C.corrupt(x, 1, 0, cfg)
    assert np.array_equal(clean, x.astype(np.float32)) and np.array_equal(noisy, clean)


def test_corrupt_golden():
    z = np.load(GOLDEN)
    cfg = C.NoiseConfig(subsample=True)
    clean, noisy = C.corrupt(z["corrupt_x"], int(z["corrupt_seed"]), int(z["corrupt_offset"]), cfg)
    assert np.array_equal(noisy, z["corrupt_noisy"]) and np.array_equal(clean, z["corrupt_clean"])


def test_adam_matches_torch_adam_when_eps_vanishes():
    """Keras Adam puts epsilon outside the bias correction; with eps -> 0 it equals torch.optim.Adam."""
    import torch
    rng = np.random.default_rng(0)
    w0 = rng.standard_normal(50)
    w, m, v = w0.copy(), np.zeros(50), np.zeros(50)
    p = torch.tensor(w0, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([p], lr=1e-2, betas=(0.9, 0.999), eps=1e-30)
    for step in range(1, 6):
        g = rng.standard_normal(50)
        w, m, v = C.adam_step(w, g, m, v, step, learning_rate=1e-2, epsilon=1e-30)
        p.grad = torch.tensor(g)
        opt.step()
    assert np.allclose(w, p.detach().numpy(), rtol=1e-10, atol=1e-12)


def test_adam_global_clipnorm():
    g = np.array([3.0, 4.0])
    w, m, v = C.adam_step(np.zeros(2), g, np.zeros(2), np.zeros(2), 1, global_clipnorm=1.0)
    w2, _, _ = C.adam_step(np.zeros(2), g / 5.0, np.zeros(2), np.zeros(2), 1)
    assert np.allclose(w, w2)
