"""Every other `type: "resnet"` configuration of the reference's builder, on the FP32 layer kernels of
csrc/generic.cu.

The tcgen05 stacks are specialised on the 16-channel, two-3x3-convs-per-block family the north star names
(`Arch` / `Denoiser`).  The reference's resnet builder accepts more (reference bfcnn/backbone_resnet.py:19-298,
bfcnn/backbone_blocks.py:74-246): 1 to 3 convs per block with any odd kernel size and filter count, grouped convs, a
depthwise middle conv (`block_depthwise`, the only in-tree resnet JSON:
bfcnn/configs/resnet_color_1x6_bn_32x128x32_1x3x1_128x128_depthwise_l1_relu.json), BatchNormalization right after the base
conv / after the last block (`add_initial_bn`, `add_final_bn`), and the learnable scalings ChannelwiseMultiplier /
Multiplier (`add_channelwise_scaling`, `add_learnable_multiplier`; bfcnn/custom_layers.py:1028-1162).  `ResnetSpec`
parses such a config, knows the Keras order and shapes of `hydra.variables` for it, and `GenericDenoiser` folds every
BatchNormalization and multiplier into a per-channel (scale, bias) of the conv before it and runs the layers one launch
each.  Same call semantics as `Denoiser` (uint8 [N,H,W,3] in, uint8 out, pow2 canvas); inference only, no CPU path.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .arch import BN_EPSILON


def _is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


@dataclass(frozen=True)
class ConvSpec:
    kernel: int
    cin: int
    cout: int
    groups: int = 1
    depth_multiplier: int = 0        # > 0: DepthwiseConv2D
    relu: bool = False
    bn: bool = False                 # BatchNormalization(center=False) after the conv (utilities.py:195-215)

    def kernel_shape(self) -> Tuple[int, ...]:
        if self.depth_multiplier > 0:
            return (self.kernel, self.kernel, self.cin, self.depth_multiplier)
        return (self.kernel, self.kernel, self.cin // self.groups, self.cout)


@dataclass(frozen=True)
class ResnetSpec:
    """Structure of a `type: "resnet"` hydra model: base conv, N identical blocks, head."""
    no_layers: int
    base: ConvSpec
    block: Tuple[ConvSpec, ...]
    head_filters: int = 32
    add_initial_bn: bool = False
    add_final_bn: bool = False
    channelwise: bool = False        # ChannelwiseMultiplier after every block's last conv and once after the blocks
    multiplier: bool = False         # Multiplier, likewise
    bn_epsilon: float = BN_EPSILON

    @property
    def filters(self) -> int:
        return self.base.cout

    @property
    def receptive_radius(self) -> int:
        return (self.base.kernel - 1) // 2 + self.no_layers * sum((c.kernel - 1) // 2 for c in self.block)

    def variable_shapes(self) -> List[Tuple[int, ...]]:
        """`hydra.variables` in Keras order: per layer, trainable then non-trainable; layers in the order
        the builder creates them (backbone_resnet.py:254-287, backbone_blocks.py:167-221, model.py:297-340)."""
        f = self.filters
        shapes: List[Tuple[int, ...]] = [self.base.kernel_shape()]
        if self.add_initial_bn:
            shapes += [(f,)] * 3
        for _ in range(self.no_layers):
            for c in self.block:
                shapes.append(c.kernel_shape())
                if c.bn:
                    shapes += [(c.cout,)] * 3                      # gamma, moving_mean, moving_variance
            if self.channelwise:
                shapes += [(f,), (1,)]                             # w0 [C] (trainable), w1 [1]
            if self.multiplier:
                shapes += [(1,), (1,)]
        if self.add_final_bn:
            shapes += [(f,)] * 3
        if self.channelwise:
            shapes += [(f,), (1,)]
        if self.multiplier:
            shapes += [(1,), (1,)]
        shapes += [(1, 1, f, self.head_filters), (1, 1, self.head_filters, 3)]
        return shapes


def spec_from_config(config: Dict) -> ResnetSpec:
    """Parse the `model` section of a pipeline config the way backbone_resnet.builder does (:19-49, :93-190)."""
    model = config.get("model", config)
    bb, dn = model["backbone"], model.get("denoiser", {})
    if str(bb.get("type", "resnet")).strip().lower() != "resnet":
        raise ValueError("only type=resnet backbones are on the hot path")
    for flag in ("add_gates", "add_gelu", "add_concat_input", "add_mean_sigma_normalization", "add_gradient_dropout"):
        if bb.get(flag, False):
            raise ValueError(f"backbone option {flag} is not implemented")
    if bb.get("selector_params", None) or bb.get("dropout_rate", -1) not in (-1, -1.0):
        raise ValueError("selector_params / dropout_rate are not implemented")
    if bb.get("use_bias", False) or dn.get("use_bias", False):
        raise ValueError("bias terms are not implemented (bias-free models only)")
    if dn.get("use_bn", False) or dn.get("use_ln", False) or str(dn.get("activation", "linear")).strip().lower() != "linear":
        raise ValueError("denoiser head options use_bn / use_ln / activation are not implemented")
    if list(bb.get("value_range", [0, 255])) != [0, 255]:
        raise ValueError("value_range must be [0, 255]")
    ishape = bb.get("input_shape", ["?", "?", 3])
    if int(ishape[-1]) != 3 or int(dn.get("output_channels", 3)) != 3:
        raise ValueError("only colour (3-channel) models are implemented")
    filters, k0 = int(bb["filters"]), int(bb["kernel_size"])
    kernels = list(bb.get("block_kernels", [3, 3]))
    bfilters = list(bb.get("block_filters", [32, 32]))
    nb = len(kernels)
    if not (1 <= nb <= 3) or len(bfilters) != nb:          # backbone_resnet.py:113-120
        raise ValueError("len(block_kernels) must be 1..3 and equal len(block_filters)")
    depthwise = list(bb.get("block_depthwise") or [-1] * nb)
    groups = list(bb.get("block_groups") or [1] * nb)
    act = str(bb.get("activation", "relu")).strip().lower()
    acts = [str(a).strip().lower() for a in (bb.get("block_activation") or [act] * nb)]
    base_act = str(bb.get("base_activation", "linear")).strip().lower()
    if len(depthwise) != nb or len(groups) != nb or len(acts) != nb:
        raise ValueError("block_depthwise / block_groups / block_activation must match block_kernels in length")
    acts[-1] = base_act                                       # backbone_resnet.py:178
    for a in acts + [base_act]:
        if a not in ("relu", "linear"):
            raise ValueError(f"activation [{a}] is not implemented (relu / linear)")
    if base_act != "linear":
        raise ValueError("base_activation must be linear")
    use_bn = bool(bb.get("use_bn", True))
    convs, cin = [], filters
    for i in range(nb):
        k = int(kernels[i])
        if k % 2 != 1:
            raise ValueError("block kernels must be odd")
        if int(depthwise[i]) != -1:
            dm = int(depthwise[i])
            c = ConvSpec(k, cin, cin * dm, 1, dm, acts[i] == "relu", use_bn and i > 0)
        else:
            c = ConvSpec(k, cin, int(bfilters[i]), int(groups[i]), 0, acts[i] == "relu", use_bn and i > 0)
            if c.cin % c.groups or c.cout % c.groups:
                raise ValueError("block_groups must divide the channel counts")
        # the first conv of a block carries no BN (backbone_blocks.py:174-178, bn_first_conv_params=False); a one-conv
        # block therefore has none at all
        convs.append(c)
        cin = c.cout
    if cin != filters:
        raise ValueError("the last conv of a block must return to the backbone's filter count (Add with the skip)")
    return ResnetSpec(no_layers=int(bb["no_layers"]), base=ConvSpec(k0, 3, filters), block=tuple(convs),
                      head_filters=int(dn.get("filters", 32)), add_initial_bn=bool(bb.get("add_initial_bn", False)),
                      add_final_bn=bool(bb.get("add_final_bn", False)),
                      channelwise=bool(bb.get("add_channelwise_scaling", False)),
                      multiplier=bool(bb.get("add_learnable_multiplier", False)))


def initial_variables(spec: ResnetSpec, seed: int = 0, trained_like: bool = True) -> List[np.ndarray]:
    """Deterministic variables of the right shapes (glorot-scaled kernels).  trained_like: BN statistics / gammas and
    multipliers away from their initial values so that every folded term is exercised."""
    from .weights import _glorot_truncated_normal
    rng = np.random.default_rng(seed)
    out: List[np.ndarray] = []

    def bn(c):
        if trained_like:
            return [(rng.uniform(0.5, 1.5, c) * 0.5).astype(np.float32), (rng.standard_normal(c) * 0.05).astype(np.float32),
                    rng.uniform(0.5, 1.5, c).astype(np.float32)]
        return [np.ones(c, np.float32), np.zeros(c, np.float32), np.ones(c, np.float32)]

    def mult(c):
        w0 = (rng.uniform(-0.3, 0.3, c)).astype(np.float32) if trained_like else np.zeros(c, np.float32)
        return [w0, np.ones(1, np.float32)]

    f = spec.filters
    out.append(_glorot_truncated_normal(rng, spec.base.kernel_shape()))
    if spec.add_initial_bn:
        out += bn(f)
    for _ in range(spec.no_layers):
        for c in spec.block:
            shp = c.kernel_shape()
            if c.depth_multiplier > 0:   # Keras glorot on a depthwise kernel: fan_in = k*k*cin ... keep the values O(1/k)
                out.append((rng.standard_normal(shp) / c.kernel).astype(np.float32) * 0.5)
            else:
                out.append(_glorot_truncated_normal(rng, shp))
            if c.bn:
                out += bn(c.cout)
        if spec.channelwise:
            out += mult(f)
        if spec.multiplier:
            out += mult(1)
    if spec.add_final_bn:
        out += bn(f)
    if spec.channelwise:
        out += mult(f)
    if spec.multiplier:
        out += mult(1)
    out.append(_glorot_truncated_normal(rng, (1, 1, f, spec.head_filters)))
    out.append((_glorot_truncated_normal(rng, (1, 1, spec.head_filters, 3)) * 0.8).astype(np.float32))
    assert [v.shape for v in out] == [tuple(s) for s in spec.variable_shapes()]
    return out


@dataclass
class _Layer:
    conv: ConvSpec
    weights: np.ndarray
    scale: Optional[np.ndarray] = None
    bias: Optional[np.ndarray] = None
    residual: bool = False           # add the block input after this layer (backbone_blocks.py:240-242)
    block_start: bool = False
    dev: Dict = field(default_factory=dict)


def fold_layers(spec: ResnetSpec, variables: Sequence[np.ndarray]) -> List[_Layer]:
    """Walk the variables in Keras order and fold every BatchNormalization (inference: (x - mean) * gamma / sqrt(var +
    eps), no beta) and every multiplier (relu(w0 + w1) * x, custom_layers.py:1081,1152) into the (scale, bias)
    of the conv before it.  An initial BN folds into the base conv; a final BN / final multipliers, which follow an Add,
    fold into the first head conv (scale on its input channels, and the BN shift through the head as a bias)."""
    shapes = spec.variable_shapes()
    if [tuple(np.shape(v)) for v in variables] != [tuple(s) for s in shapes]:
        raise ValueError("variables do not match the architecture of the config")
    it = iter([np.asarray(v, np.float64) for v in variables])
    eps = spec.bn_epsilon

    def take_bn():
        g, m, v = next(it), next(it), next(it)
        s = g / np.sqrt(v + eps)
        return s, -m * s

    def take_mult():
        w0, w1 = next(it), next(it)
        return np.maximum(w0 + w1, 0.0)              # activation="relu" (backbone_resnet.py:192-204)

    f = spec.filters
    layers: List[_Layer] = []
    base = _Layer(spec.base, next(it))
    if spec.add_initial_bn:
        base.scale, base.bias = take_bn()
    layers.append(base)
    for _ in range(spec.no_layers):
        blk: List[_Layer] = []
        for c in spec.block:
            ly = _Layer(c, next(it))
            if c.bn:
                ly.scale, ly.bias = take_bn()
            blk.append(ly)
        last = blk[-1]
        for on in (spec.channelwise, spec.multiplier):
            if on:
                m = take_mult() * np.ones(f)
                last.scale = m if last.scale is None else last.scale * m
                last.bias = None if last.bias is None else last.bias * m
        blk[0].block_start = True
        last.residual = True
        layers += blk
    a, b = np.ones(f), np.zeros(f)               # what sits between the last Add and the head: x -> a * x + b
    if spec.add_final_bn:
        a, b = take_bn()
    for on in (spec.channelwise, spec.multiplier):
        if on:
            m = take_mult() * np.ones(f)
            a, b = a * m, b * m
    h0, h1 = next(it), next(it)
    h0_folded = h0 * a.reshape(1, 1, f, 1)         # head conv on (a * x + b) = (h0 * a) x + h0^T b
    head0 = _Layer(ConvSpec(1, f, spec.head_filters), h0_folded)
    shift = (b.reshape(1, f) @ h0[0, 0]).reshape(-1)
    if np.any(shift != 0):
        head0.bias = shift
    layers += [head0, _Layer(ConvSpec(1, spec.head_filters, 3), h1)]
    return layers


class GenericDenoiser:
    """`DenoiserModule.__call__` (module_denoiser.py:39-75) for any supported resnet spec, one FP32 launch per layer."""

    def __init__(self, spec: ResnetSpec, variables: Sequence[np.ndarray], *, device: int = 0, pad_pow2: bool = True, name: str = "",
                 precision: str = "fp32"):
        import torch
        if precision not in ("fp32",):
            raise ValueError("the generic resnet path computes in fp32 only (the tcgen05 stacks cover the "
                             "resnet_color_1xN_bn_16x3x3 family)")
        self._lib = _native.load_library()
        if not torch.cuda.is_available():
            raise _native.NativeError(-2, "no CUDA device visible: libbfcnn_b200 has no CPU path")
        self.spec, self.name, self.device, self.pad_pow2, self.precision = spec, name, int(device), bool(pad_pow2), precision
        self._variables = [np.asarray(v, np.float32) for v in variables]
        self.layers = fold_layers(spec, self._variables)
        dev = torch.device("cuda", self.device)
        for ly in self.layers:
            ly.dev["w"] = torch.from_numpy(np.ascontiguousarray(ly.weights, np.float32)).to(dev)
            ly.dev["s"] = None if ly.scale is None else torch.from_numpy(np.ascontiguousarray(ly.scale, np.float32)).to(dev)
            ly.dev["b"] = None if ly.bias is None else torch.from_numpy(np.ascontiguousarray(ly.bias, np.float32)).to(dev)
        self._launches = 0

    def get_weights(self) -> List[np.ndarray]:
        return [v.copy() for v in self._variables]

    def launch_count(self) -> int:
        return self._launches

    def close(self):
        self.layers = []

    def __call__(self, image, *, pad_pow2: Optional[bool] = None, return_float: bool = False, out=None, precision=None):
        import torch
        if precision not in (None, "fp32"):
            raise ValueError("the generic resnet path computes in fp32 only")
        as_numpy = not _is_torch(image)
        x = torch.from_numpy(np.ascontiguousarray(image)) if as_numpy else image
        if x.dtype != torch.uint8:
            raise TypeError("image must be uint8")          # tf.TensorSpec(dtype=tf.uint8), module_denoiser.py:44
        if x.dim() != 4 or x.shape[-1] != 3:
            raise ValueError("image must have shape [N,H,W,3]")
        was_cuda = x.is_cuda
        dev = torch.device("cuda", self.device)
        x = x.to(dev).contiguous()
        n, h, w, _ = x.shape
        odt = torch.float32 if return_float else torch.uint8
        res = torch.empty((n, h, w, 3), dtype=odt, device=dev)
        if n * h * w > 0:
            pad = self.pad_pow2 if pad_pow2 is None else bool(pad_pow2)
            hc = (1 << max(0, (h - 1).bit_length())) if pad else h          # utilities.py:740-741
            wc = (1 << max(0, (w - 1).bit_length())) if pad else w
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            cur = torch.empty((n, hc, wc, 3), dtype=torch.float32, device=dev)
            _native.check(self._lib.bfcnn_generic_prepare(self.device, x.data_ptr(), cur.data_ptr(), n, h, w, hc, wc, st))
            skip = None
            for ly in self.layers:
                c = ly.conv
                if ly.block_start:
                    skip = cur
                nxt = torch.empty((n, hc, wc, c.cout), dtype=torch.float32, device=dev)
                _native.check(self._lib.bfcnn_generic_conv2d(
                    self.device, cur.data_ptr(), nxt.data_ptr(), ly.dev["w"].data_ptr(),
                    None if ly.dev["s"] is None else ly.dev["s"].data_ptr(), None if ly.dev["b"] is None else ly.dev["b"].data_ptr(),
                    skip.data_ptr() if ly.residual else None, n, hc, wc, c.cin, c.cout, c.kernel, c.groups, c.depth_multiplier,
                    int(c.relu), st))
                cur = nxt
            _native.check(self._lib.bfcnn_generic_finish(self.device, cur.data_ptr(), res.data_ptr(), n, h, w, hc, wc,
                                                         0 if return_float else 1, st))
            self._launches += len(self.layers) + 2
        if out is not None:
            out.copy_(res) if _is_torch(out) else np.copyto(out, res.cpu().numpy())
            return out
        if as_numpy:
            return res.cpu().numpy()
        return res if was_cuda else res.cpu()
