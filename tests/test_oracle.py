"""CPU checks of the oracle itself (no GPU): the torch-conv restatement against an independent
direct numpy convolution, BN folding / head collapse identities, golden fixtures, loss
semantics and the autograd backward against finite differences."""
import os

import numpy as np
import pytest
import torch

from blind_image_denoising_b200 import Arch, synthetic_variables
from oracle import bfcnn_oracle as O


def test_conv_same_matches_direct():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 9, 11, 5))
    for k in (1, 3, 7):
        w = rng.standard_normal((k, k, 5, 4))
        a = O._conv_same(torch.as_tensor(x).permute(0, 3, 1, 2), torch.as_tensor(w)).permute(0, 2, 3, 1).numpy()
        b = O.conv_same_direct(x, w)
        assert np.abs(a - b).max() < 1e-12


def test_forward_matches_handwritten_numpy():
    """hydra_forward vs a forward written with conv_same_direct + folded BN + collapsed head."""
    arch = Arch(no_layers=3)
    v = synthetic_variables(arch, 1)
    rng = np.random.default_rng(2)
    x = rng.integers(0, 256, size=(1, 12, 10, 3)).astype(np.float64)
    ref = O.hydra_forward(v, x)
    base, blocks, h0, h1 = O.split_variables(v)
    f = O.conv_same_direct(x / 255.0 - 0.5, base)
    for wa, wb, g, m, var in blocks:
        t = np.maximum(O.conv_same_direct(f, wa), 0.0)
        wf, bf = O.fold_bn(wb, g, m, var)
        f = f + O.conv_same_direct(t, wf) + bf.reshape(1, 1, 1, -1)
    y = np.tanh(2.0 * (f @ O.collapse_head(h0, h1))) * 0.51
    y = (np.clip(y, -0.5, 0.5) + 0.5) * 255.0
    assert np.abs(y - ref).max() < 1e-9


def test_golden_fixture_reproduces():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inference_golden.npz"))
    for n in (6, 12, 18):
        y, u8 = O.denoise(synthetic_variables(Arch(no_layers=n), 0), g[f"x_{n}"], pad_pow2=True)
        assert np.abs(y - g[f"y_{n}"]).max() < 1e-3
        assert np.array_equal(u8, g[f"u8_{n}"])
        # outputs are not degenerate: they span the range and touch the clamps only rarely
        assert g[f"y_{n}"].std() > 20 and ((g[f"u8_{n}"] == 0) | (g[f"u8_{n}"] == 255)).mean() < 0.2


def test_pow2_canvas_semantics():
    """module_denoiser.py:56-68: raw zeros bottom/right change the border band only."""
    arch = Arch(no_layers=2)
    v = synthetic_variables(arch, 0)
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(1, 20, 24, 3), dtype=np.uint8)
    yp, _ = O.denoise(v, x, pad_pow2=True)
    yn, _ = O.denoise(v, x, pad_pow2=False)
    R = arch.receptive_radius
    assert np.abs(yp[:, :20 - R, :24 - R] - yn[:, :20 - R, :24 - R]).max() < 1e-9
    assert np.abs(yp - yn).max() > 1e-3
    assert O.next_pow2(1) == 1 and O.next_pow2(256) == 256 and O.next_pow2(257) == 512 and O.next_pow2(2160) == 4096


def test_round_half_even():
    assert list(O.round_half_even_u8(np.array([0.5, 1.5, 2.5, 254.5, 255.4, -0.2]))) == [0, 2, 2, 254, 255, 0]


def test_hinged_mae_semantics():
    """keras relu(x, threshold) passes values above the hinge UNCHANGED."""
    e = torch.tensor([[[[0.2, -0.5, 0.5001, -3.0, 300.0, 0.0]]]], dtype=torch.float64)
    d = O.keras_relu_threshold(torch.abs(e), 0.5, 255.0)
    assert np.allclose(d.numpy().ravel(), [0, 0, 0.5001, 3.0, 255.0, 0])
    assert abs(float(O.mae_diff(e, 0.5, 255.0)) - (0.5001 + 3 + 255) / 6) < 1e-12


def test_train_step_gradients_finite_difference():
    arch = Arch(no_layers=2)
    v = [a.astype(np.float64) for a in synthetic_variables(arch, 3)]
    rng = np.random.default_rng(4)
    clean = rng.integers(0, 256, size=(2, 8, 8, 3)).astype(np.float64)
    noisy = np.rint(clean + rng.standard_normal(clean.shape) * 20)
    r = O.train_step(v, clean, noisy, hinge=0.5)
    assert len(r["grads"]) == 3 + 3 * arch.no_layers
    train_idx = [i for i, t in enumerate(arch.trainable_mask()) if t]
    eps = 1e-6
    for gi, vi in list(enumerate(train_idx))[::2]:
        flat_idx = np.random.default_rng(gi).integers(0, v[vi].size, size=2)
        for fi in flat_idx:
            vp = [a.copy() for a in v]
            vm = [a.copy() for a in v]
            vp[vi].reshape(-1)[fi] += eps
            vm[vi].reshape(-1)[fi] -= eps
            fd = (O.train_step(vp, clean, noisy)["total"] - O.train_step(vm, clean, noisy)["total"]) / (2 * eps)
            an = r["grads"][gi].reshape(-1)[fi]
            assert abs(fd - an) <= 1e-4 * max(1.0, abs(an)) + 1e-6, (gi, fi, fd, an)


def test_train_step_bn_moving_update():
    arch = Arch(no_layers=1)
    v = synthetic_variables(arch, 0)
    rng = np.random.default_rng(0)
    clean = rng.integers(0, 256, size=(2, 6, 6, 3)).astype(np.float64)
    r = O.train_step(v, clean, clean)
    m, var = r["new_moving"][0]
    assert m.shape == (16,) and var.shape == (16,)
    # momentum 0.995: the moving stats move by at most 0.5 % of the gap per step
    assert np.all(np.abs(m - v[4] * 0.995) < 1.0) and np.all(var > 0)


def test_ssim_matches_independent_formula():
    """oracle.ssim_tf (tf.image.ssim restated) against the textbook SSIM with a separable Gaussian from scipy:
    ((2 mx my + c1)(2 sxy + c2)) / ((mx^2 + my^2 + c1)(sx^2 + sy^2 + c2)), VALID windows, mean over windows, channels."""
    import torch
    from scipy.ndimage import correlate1d
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(2, 19, 23, 3)).astype(np.float64)
    y = np.clip(x + rng.normal(0, 25, size=x.shape), 0, 255)
    got = O.ssim_tf(torch.tensor(x), torch.tensor(y), max_val=255.0, filter_size=7).numpy()
    c = np.arange(7) - 3.0
    g = np.exp(-0.5 * c * c / 1.5 ** 2); g /= g.sum()

    def filt(a):   # VALID 7x7 separable correlation over H, W
        a = correlate1d(a, g, axis=1, mode="constant")[:, 3:-3]
        return correlate1d(a, g, axis=2, mode="constant")[:, :, 3:-3]
    mx, my = filt(x), filt(y)
    sxx, syy, sxy = filt(x * x) - mx * mx, filt(y * y) - my * my, filt(x * y) - mx * my
    c1, c2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    ssim = ((2 * mx * my + c1) * (2 * sxy + c2)) / ((mx * mx + my * my + c1) * (sxx + syy + c2))
    want = ssim.mean(axis=(1, 2)).mean(axis=1)
    assert np.allclose(got, want, rtol=1e-10)
    assert np.allclose(O.ssim_tf(torch.tensor(x), torch.tensor(x)).numpy(), 1.0)
