"""CPU ORACLE for the bfcnn resnet denoiser hot path -- TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  The product package
(`blind_image_denoising_b200/`) never does; it fails loudly without its CUDA library.

PARITY UNPINNED: the reference is TensorFlow/Keras Python (tensorflow==2.13.1,
`/root/reference/requirements.txt:9`), TensorFlow is not installable here, and the
reference's tests hold no golden vectors for this path.  This file
restates the reference's algorithm line by line from its sources, and encodes the
TF/Keras 2.13 op semantics listed in SURVEY 8c from their published behaviour.
It is cross-checked against an independent pure-numpy direct convolution
(tests/test_oracle.py) and, for the backward pass, against finite differences.

All citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPSILON = 1e-3     # bfcnn/constants.py:9
BN_MOMENTUM = 0.995   # bfcnn/constants.py:11
L1_COEFF = 0.01       # keras string regularizer "l1" -> L1(0.01) (backbone_resnet.py:35,145,162)
L2_COEFF = 0.01       # keras string regularizer "l2" -> L2(0.01) (model.py:275)


# ----------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------
def _t(a, dtype):
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def _conv_same(x: torch.Tensor, w_hwio: torch.Tensor) -> torch.Tensor:
    """Keras Conv2D(padding="same", strides=1, use_bias=False): cross-correlation,
    HWIO kernel, symmetric zero pad (k-1)/2 (utilities.py:195-196). x is NCHW."""
    k = w_hwio.shape[0]
    w = w_hwio.permute(3, 2, 0, 1).contiguous()  # HWIO -> OIHW
    return F.conv2d(x, w, bias=None, stride=1, padding=(k - 1) // 2)


def next_pow2(n: int) -> int:
    """utilities.py:740-741: 2**ceil(log2(n)) evaluated in float32 (exact for the
    sizes of interest; 1 -> 1)."""
    return 1 << max(0, int(math.ceil(math.log2(n)))) if n > 1 else 1


def split_variables(variables: Sequence[np.ndarray]):
    """Keras variable order -> (base, [(wa, wb, gamma, mean, var)], h0, h1)."""
    n = (len(variables) - 3) // 5
    base = variables[0]
    blocks = [tuple(variables[1 + 5 * i: 1 + 5 * i + 5]) for i in range(n)]
    return base, blocks, variables[-2], variables[-1]


# ----------------------------------------------------------------------------
# forward (inference): module_denoiser.py:53-73
# ----------------------------------------------------------------------------
def hydra_forward(variables: Sequence[np.ndarray], x_nhwc: np.ndarray, *,
                  dtype=torch.float64, head_literal: bool = False,
                  return_features: bool = False):
    """model.py:100-116,136-140 in inference mode.  x_nhwc is float in 0..255.
    Returns the float prediction in NHWC (0..255 scale unless head_literal)."""
    base, blocks, h0, h1 = split_variables(variables)
    x = _t(x_nhwc, dtype).permute(0, 3, 1, 2)
    # layer_normalize, utilities.py:449-461
    x = torch.clamp(x, 0.0, 255.0)
    x = (x - 0.0) / (255.0 - 0.0) - 0.5
    # base conv, backbone_resnet.py:258-262 (linear, no BN)
    x = _conv_same(x, _t(base, dtype))
    feats = [x]
    # residual blocks, backbone_blocks.py:167-246
    for wa, wb, gamma, mean, var in blocks:
        prev = x
        t = torch.relu(_conv_same(x, _t(wa, dtype)))               # :174-178, no BN
        u = _conv_same(t, _t(wb, dtype))                           # :191-196
        g, m, v = _t(gamma, dtype), _t(mean, dtype), _t(var, dtype)
        # BatchNormalization(center=False) inference: (u-mean)*gamma/sqrt(var+eps)  (F6)
        u = (u - m.view(1, -1, 1, 1)) * (g / torch.sqrt(v + BN_EPSILON)).view(1, -1, 1, 1)
        x = u + prev                                               # :240-242 Add([x, previous])
        feats.append(x)
    # denoiser head, model.py:297-342
    y = _conv_same(x, _t(h0, dtype))
    y = _conv_same(y, _t(h1, dtype))
    y = torch.tanh(2.0 * y) * 0.51
    if not head_literal:
        # layer_denormalize, utilities.py:435-443
        y = (torch.clamp(y, -0.5, 0.5) + 0.5) * 255.0
    out = y.permute(0, 2, 3, 1).contiguous().numpy()
    if return_features:
        return out, [f.permute(0, 2, 3, 1).contiguous().numpy() for f in feats]
    return out


def round_half_even_u8(y: np.ndarray) -> np.ndarray:
    """tf.round (half to even) then tf.cast(uint8)  (module_denoiser.py:71-73)."""
    return np.clip(np.rint(y), 0, 255).astype(np.uint8)


def denoise(variables: Sequence[np.ndarray], image_u8: np.ndarray, *,
            pad_pow2: bool = True, dtype=torch.float64,
            head_literal: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """DenoiserModule.__call__ (module_denoiser.py:39-75).

    Returns (pre-round float NHWC, uint8 NHWC).  With pad_pow2 the image is
    embedded in the 2^k x 2^k' canvas with raw zeros bottom/right
    (utilities.py:736-751), run whole, then cropped (utilities.py:755-764)."""
    image_u8 = np.asarray(image_u8)
    assert image_u8.dtype == np.uint8 and image_u8.ndim == 4
    n, h, w, c = image_u8.shape
    x = image_u8.astype(np.float64)
    if pad_pow2:
        hc, wc = next_pow2(h), next_pow2(w)
        canvas = np.zeros((n, hc, wc, c), dtype=np.float64)
        canvas[:, :h, :w, :] = x
        x = canvas
    y = hydra_forward(variables, x, dtype=dtype, head_literal=head_literal)
    y = y[:, :h, :w, :]
    return y, round_half_even_u8(y)


# ----------------------------------------------------------------------------
# BN folding restated (what the native packer must reproduce)
# ----------------------------------------------------------------------------
def fold_bn(wb: np.ndarray, gamma, mean, var, eps: float = BN_EPSILON):
    """w' = w*gamma/sqrt(var+eps) per output channel; b' = -mean*gamma/sqrt(var+eps) (F6)."""
    s = np.asarray(gamma, np.float64) / np.sqrt(np.asarray(var, np.float64) + eps)
    return np.asarray(wb, np.float64) * s.reshape(1, 1, 1, -1), -np.asarray(mean, np.float64) * s


def collapse_head(h0: np.ndarray, h1: np.ndarray) -> np.ndarray:
    """The two linear 1x1 convs 16->F->3 multiply into one [16,3] matrix."""
    return np.asarray(h0, np.float64)[0, 0] @ np.asarray(h1, np.float64)[0, 0]


# ----------------------------------------------------------------------------
# independent direct convolution (numpy loops) to check _conv_same itself
# ----------------------------------------------------------------------------
def conv_same_direct(x_nhwc: np.ndarray, w_hwio: np.ndarray) -> np.ndarray:
    x = np.asarray(x_nhwc, np.float64)
    w = np.asarray(w_hwio, np.float64)
    n, h, wd, ci = x.shape
    k = w.shape[0]
    r = (k - 1) // 2
    xp = np.zeros((n, h + 2 * r, wd + 2 * r, ci))
    xp[:, r:r + h, r:r + wd, :] = x
    out = np.zeros((n, h, wd, w.shape[3]))
    for dy in range(k):
        for dx in range(k):
            out += np.einsum("nhwc,co->nhwo", xp[:, dy:dy + h, dx:dx + wd, :], w[dy, dx])
    return out


# ----------------------------------------------------------------------------
# hinged MAE + metrics: loss.py:40-65, 71-86, 92-131, 190-247
# ----------------------------------------------------------------------------
def keras_relu_threshold(x: torch.Tensor, threshold: float, max_value: float) -> torch.Tensor:
    """keras.activations.relu(x, threshold=t, max_value=m) (Keras 2.13 backend.relu):
    t != 0: x*[x > t] (strict) ; t == 0: relu(x); then clip to [0, m]."""
    if threshold != 0.0:
        x = x * (x > threshold).to(x.dtype)
    else:
        x = torch.relu(x)
    return torch.clamp(x, 0.0, max_value)


def mae_diff(error: torch.Tensor, hinge: float = 0.0, cutoff: float = 255.0) -> torch.Tensor:
    d = keras_relu_threshold(torch.abs(error), hinge, cutoff)   # loss.py:53-57
    d = d.mean(dim=(1, 2, 3))                                    # :60 (the 1-tuple wrapper only adds a unit axis)
    return d.mean()                                              # :63


def rmse_diff(error: torch.Tensor, hinge: float = 0.0, cutoff: float = 255.0 * 255.0) -> torch.Tensor:
    d = keras_relu_threshold(error, hinge, cutoff)               # loss.py:104-108 (relu of the SIGNED error)
    d = d * d
    d = d.mean(dim=(1, 2, 3))
    d = torch.sqrt(d + 1e-3)                                     # DEFAULT_EPSILON constants.py:7
    return d.mean()


def ssim_tf(img1: torch.Tensor, img2: torch.Tensor, max_val: float = 255.0, filter_size: int = 7,
            filter_sigma: float = 1.5, k1: float = 0.01, k2: float = 0.03) -> torch.Tensor:
    """tf.image.ssim (TF 2.13 image_ops_impl: _fspecial_gauss, _ssim_helper, _ssim_per_channel) on NHWC tensors, as
    called by loss.py:217-224 (filter_size=7, max_val=255).  Returns one value per image:
    mean over channels of the spatial mean (VALID windows) of luminance * contrast-structure."""
    dtype = img1.dtype
    coords = torch.arange(filter_size, dtype=dtype) - (filter_size - 1.0) / 2.0
    g = -0.5 * coords * coords / (filter_sigma * filter_sigma)
    g2 = torch.softmax((g.view(1, -1) + g.view(-1, 1)).reshape(-1), dim=0).view(1, 1, filter_size, filter_size)
    c = img1.shape[-1]
    k = g2.repeat(c, 1, 1, 1)

    def reducer(t):   # depthwise_conv2d, strides 1, padding VALID
        return F.conv2d(t.permute(0, 3, 1, 2), k, groups=c)

    c1, c2 = (k1 * max_val) ** 2, (k2 * max_val) ** 2
    mean0, mean1 = reducer(img1), reducer(img2)
    num0 = mean0 * mean1 * 2.0
    den0 = mean0 * mean0 + mean1 * mean1
    luminance = (num0 + c1) / (den0 + c1)
    num1 = reducer(img1 * img2) * 2.0
    den1 = reducer(img1 * img1 + img2 * img2)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    return (luminance * cs).mean(dim=(2, 3)).mean(dim=1)


def denoiser_loss(gt: torch.Tensor, pred: torch.Tensor, *, hinge=0.0, cutoff=255.0,
                  mae_multiplier=1.0, mse_multiplier=0.0, ssim_multiplier=0.0) -> Dict[str, torch.Tensor]:
    """loss.py:190-247 (the reference's default ssim_multiplier is 1.0, loss.py:171; the _l1_ recipes use 0)."""
    e = gt - pred
    mae_actual = mae_diff(e, 0.0, 255.0)
    mse_actual = rmse_diff(e, 0.0, 255.0)
    total = torch.zeros((), dtype=gt.dtype)
    if mae_multiplier > 0.0:
        total = total + mae_diff(e, hinge, cutoff) * mae_multiplier
    ssim_loss = torch.zeros((), dtype=gt.dtype)
    if ssim_multiplier > 0.0:                                    # loss.py:217-225
        ssim_loss = 1.0 - ssim_tf(gt, pred, max_val=255.0, filter_size=7).mean()
        total = total + ssim_loss * ssim_multiplier
    if mse_multiplier > 0.0:
        total = total + rmse_diff(e, hinge, cutoff * cutoff) * mse_multiplier
    return {"total_loss": total, "mae_loss": mae_actual, "mse_loss": mse_actual, "ssim_loss": ssim_loss}


# ----------------------------------------------------------------------------
# multi-scale ground truth: utilities.py:625-685 (multiscales_generator_fn), used at train_loop.py:239-247,274
# ----------------------------------------------------------------------------
def multiscales(x_nhwc: np.ndarray, no_scales: int, clip_values: bool = True, round_values: bool = True,
                dtype=np.float32) -> List[np.ndarray]:
    """[n, pool(n), pool(pool(n)), ...]: tf.nn.avg_pool2d(ksize 2x2, strides 2, VALID), then clip [0,255], then tf.round
    (half to even) per level (utilities.py:655-668)."""
    n = np.asarray(x_nhwc, dtype)
    scales = [n]
    for _ in range(no_scales):
        b, h, w, c = n.shape
        ho, wo = h // 2, w // 2
        v = n[:, :2 * ho, :2 * wo, :].reshape(b, ho, 2, wo, 2, c)
        n = ((v[:, :, 0, :, 0] + v[:, :, 0, :, 1]) + (v[:, :, 1, :, 0] + v[:, :, 1, :, 1])) * dtype(0.25)
        if clip_values:
            n = np.clip(n, 0.0, 255.0)
        if round_values:
            n = np.rint(n)
        n = n.astype(dtype)
        scales.append(n)
    return scales


# ----------------------------------------------------------------------------
# training step: train_loop.py:263-312
# ----------------------------------------------------------------------------
def train_step(variables: Sequence[np.ndarray], clean_nhwc: np.ndarray, noisy_nhwc: np.ndarray, *,
               hinge: float = 0.5, cutoff: float = 255.0, mae_multiplier: float = 1.0,
               mse_multiplier: float = 0.0, regularization: float = 0.01, ssim_multiplier: float = 0.0,
               dtype=torch.float64, head_literal: bool = False, relu_flips: Optional[Dict[int, np.ndarray]] = None,
               return_preact: bool = False):
    """One tape step: hydra(noisy, training=True) -> denoiser loss (+ L1/L2 weight
    regularisation * lambda) -> gradients w.r.t. trainable variables.

    BN in training mode: normalise with the biased batch variance; moving stats
    <- m*old + (1-m)*batch (moving_var uses the unbiased variance).
    Returns dict(total, denoiser_total, mae, mse, reg, grads[list in trainable order],
    prediction, new_moving[list of (mean,var)]).

    relu_flips / return_preact serve the gradient-parity tests only: `preact[i]` is block i's conv_a output before
    the ReLU (NCHW); `relu_flips[i]` (bool, NCHW) inverts the ReLU's 0/1 derivative mask of those units -- two
    FP32-grade implementations may disagree on the sign of a pre-activation that is below their rounding error, which
    moves that unit's whole share of the gradient while leaving the forward values unchanged to ~1e-6."""
    base, blocks, h0, h1 = split_variables(variables)
    params = {}

    def P(name, a):
        p = _t(a, dtype).clone().requires_grad_(True)
        params[name] = p
        return p

    wbase = P("base", base)
    x = _t(noisy_nhwc, dtype).permute(0, 3, 1, 2)
    x = torch.clamp(x, 0.0, 255.0) / 255.0 - 0.5
    x = _conv_same(x, wbase)
    new_moving = []
    preacts = []
    reg = torch.abs(wbase).sum() * L1_COEFF
    order = ["base"]
    for i, (wa, wb, gamma, mean, var) in enumerate(blocks):
        pa, pb, pg = P(f"a{i}", wa), P(f"b{i}", wb), P(f"g{i}", gamma)
        order += [f"a{i}", f"b{i}", f"g{i}"]
        prev = x
        pre = _conv_same(x, pa)
        preacts.append(pre.detach().numpy())
        if relu_flips is not None and i in relu_flips:
            mask = (pre.detach() > 0) ^ torch.as_tensor(np.asarray(relu_flips[i], bool))
            t = pre * mask.to(dtype)
        else:
            t = torch.relu(pre)
        u = _conv_same(t, pb)
        bm = u.mean(dim=(0, 2, 3))
        bv = ((u - bm.view(1, -1, 1, 1)) ** 2).mean(dim=(0, 2, 3))       # biased
        cnt = u.shape[0] * u.shape[2] * u.shape[3]
        u = (u - bm.view(1, -1, 1, 1)) / torch.sqrt(bv + BN_EPSILON).view(1, -1, 1, 1) * pg.view(1, -1, 1, 1)
        x = u + prev
        ub = bv.detach() * (cnt / max(cnt - 1, 1))
        new_moving.append((
            (_t(mean, dtype) * BN_MOMENTUM + bm.detach() * (1 - BN_MOMENTUM)).numpy(),
            (_t(var, dtype) * BN_MOMENTUM + ub * (1 - BN_MOMENTUM)).numpy()))
        reg = reg + (torch.abs(pa).sum() + torch.abs(pb).sum()) * L1_COEFF
    ph0, ph1 = P("h0", h0), P("h1", h1)
    order += ["h0", "h1"]
    y = _conv_same(_conv_same(x, ph0), ph1)
    y = torch.tanh(2.0 * y) * 0.51
    if not head_literal:
        y = (torch.clamp(y, -0.5, 0.5) + 0.5) * 255.0
    reg = reg + ((ph0 * ph0).sum() + (ph1 * ph1).sum()) * L2_COEFF
    pred = y.permute(0, 2, 3, 1)
    gt = _t(clean_nhwc, dtype)
    dl = denoiser_loss(gt, pred, hinge=hinge, cutoff=cutoff, mae_multiplier=mae_multiplier,
                       mse_multiplier=mse_multiplier, ssim_multiplier=ssim_multiplier)
    total = dl["total_loss"] + reg * regularization            # train_loop.py:297-301
    grads = torch.autograd.grad(total, [params[k] for k in order])
    extra = {"preact": preacts} if return_preact else {}
    return {
        **extra,
        "total": float(total.detach()), "denoiser_total": float(dl["total_loss"].detach()),
        "mae": float(dl["mae_loss"].detach()), "mse": float(dl["mse_loss"].detach()), "reg": float(reg.detach()),
        "ssim": float(dl["ssim_loss"].detach()),
        "grads": [g.numpy() for g in grads],
        "prediction": pred.detach().numpy(), "new_moving": new_moving,
    }


# ----------------------------------------------------------------------------
# CPU baseline leg (bench.py): fp32 torch-CPU restatement, all host threads
# ----------------------------------------------------------------------------
def denoise_fp32_cpu(variables, image_u8, pad_pow2=True):
    return denoise(variables, image_u8, pad_pow2=pad_pow2, dtype=torch.float32)
