"""CPU ORACLE for the general `type: "resnet"` backbone configurations -- TEST INFRASTRUCTURE ONLY (see bfcnn_oracle.py;
PARITY UNPINNED for the same reason: TensorFlow cannot run here).

Restates, layer by layer and WITHOUT any folding, what the reference builds for a resnet config:
  /root/reference/bfcnn/backbone_resnet.py:93-298   argument fixing, conv params, base conv, [initial BN], blocks,
                                                     [final BN], [ChannelwiseMultiplier], [Multiplier]
  /root/reference/bfcnn/backbone_blocks.py:167-246   per block: conv1 (no BN) -> conv2 (BN) -> conv3 (BN) ->
                                                     [ChannelwiseMultiplier] -> [Multiplier] -> Add
  /root/reference/bfcnn/utilities.py:195-215         conv2d_wrapper: Conv2D / DepthwiseConv2D (linear) -> BN -> activation
  /root/reference/bfcnn/custom_layers.py:1028-1162   Multiplier / ChannelwiseMultiplier: activation(w0 + w1) * x
  /root/reference/bfcnn/model.py:100-151,297-342     normalise -> backbone -> head -> tanh(2y)*0.51 -> denormalise
  /root/reference/bfcnn/module_denoiser.py:53-73     uint8 -> pow2 canvas -> hydra -> crop -> round -> uint8
It shares no code with blind_image_denoising_b200/generic.py (which FOLDS the normalisations into the convs).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .bfcnn_oracle import BN_EPSILON, next_pow2, round_half_even_u8


def _conv(x: torch.Tensor, kernel: np.ndarray, groups: int, depth_multiplier: int) -> torch.Tensor:
    """Keras Conv2D(groups) [kh,kw,cin/groups,cout] or DepthwiseConv2D [kh,kw,cin,dm], "same", stride 1, no bias; x NCHW."""
    k = torch.as_tensor(np.asarray(kernel), dtype=x.dtype)
    kh = k.shape[0]
    if depth_multiplier > 0:
        cin, dm = k.shape[2], k.shape[3]
        # tf.nn.depthwise_conv2d: output channel ci * dm + m  ==  torch grouped conv with groups = cin, weight [cin*dm, 1, kh, kw]
        w = k.permute(2, 3, 0, 1).reshape(cin * dm, 1, kh, kh)
        return F.conv2d(x, w, padding=(kh - 1) // 2, groups=cin)
    w = k.permute(3, 2, 0, 1).contiguous()
    return F.conv2d(x, w, padding=(kh - 1) // 2, groups=groups)


def _bn(x, gamma, mean, var):
    g, m, v = (torch.as_tensor(np.asarray(a), dtype=x.dtype).view(1, -1, 1, 1) for a in (gamma, mean, var))
    return (x - m) * g / torch.sqrt(v + BN_EPSILON)


def _mult(x, w0, w1):
    s = torch.relu(torch.as_tensor(np.asarray(w0), dtype=x.dtype) + torch.as_tensor(np.asarray(w1), dtype=x.dtype))
    return x * s.view(1, -1, 1, 1)


def denoise_generic(model_config: Dict, variables: Sequence[np.ndarray], image_u8: np.ndarray, *, pad_pow2: bool = True,
                    dtype=torch.float64) -> Tuple[np.ndarray, np.ndarray]:
    """(pre-round float NHWC, uint8 NHWC) of DenoiserModule.__call__ for a resnet `model` config section."""
    bb, dn = model_config["backbone"], model_config.get("denoiser", {})
    kernels = list(bb.get("block_kernels", [3, 3]))
    nb = len(kernels)
    depthwise = list(bb.get("block_depthwise") or [-1] * nb)
    groups = list(bb.get("block_groups") or [1] * nb)
    act = str(bb.get("activation", "relu")).strip().lower()
    acts = [str(a).strip().lower() for a in (bb.get("block_activation") or [act] * nb)]
    acts[-1] = str(bb.get("base_activation", "linear")).strip().lower()         # backbone_resnet.py:178
    use_bn = bool(bb.get("use_bn", True))
    it = iter(variables)

    image_u8 = np.asarray(image_u8)
    n, h, w, _ = image_u8.shape
    x = image_u8.astype(np.float64)
    if pad_pow2:                                                                 # utilities.py:736-751
        canvas = np.zeros((n, next_pow2(h), next_pow2(w), 3))
        canvas[:, :h, :w] = x
        x = canvas
    x = torch.as_tensor(x, dtype=dtype).permute(0, 3, 1, 2)
    x = torch.clamp(x, 0.0, 255.0) / 255.0 - 0.5                                 # utilities.py:449-461
    x = _conv(x, next(it), 1, 0)                                                 # base conv, linear, no BN
    if bb.get("add_initial_bn", False):
        x = _bn(x, next(it), next(it), next(it))
    for _ in range(int(bb["no_layers"])):
        prev = x
        for i in range(nb):
            x = _conv(x, next(it), int(groups[i]), int(depthwise[i]) if int(depthwise[i]) != -1 else 0)
            if use_bn and i > 0:                                                 # backbone_blocks.py:174-213
                x = _bn(x, next(it), next(it), next(it))
            if acts[i] == "relu":
                x = torch.relu(x)
        if bb.get("add_channelwise_scaling", False):
            x = _mult(x, next(it), next(it))
        if bb.get("add_learnable_multiplier", False):
            x = _mult(x, next(it), next(it))
        x = x + prev                                                             # backbone_blocks.py:240-242
    if bb.get("add_final_bn", False):
        x = _bn(x, next(it), next(it), next(it))
    if bb.get("add_channelwise_scaling", False):
        x = _mult(x, next(it), next(it))
    if bb.get("add_learnable_multiplier", False):
        x = _mult(x, next(it), next(it))
    y = _conv(_conv(x, next(it), 1, 0), next(it), 1, 0)                          # model.py:297-340
    assert next(it, None) is None, "variables left over"
    y = torch.tanh(2.0 * y) * 0.51
    y = (torch.clamp(y, -0.5, 0.5) + 0.5) * 255.0                                # utilities.py:435-443
    y = y.permute(0, 2, 3, 1).contiguous().numpy()[:, :h, :w, :]
    return y, round_half_even_u8(y)
