"""Architecture description of the bias-free ResNet denoiser hot path.

Mirrors the hyper-parameters that the reference passes to its resnet builder
(`bfcnn/backbone_resnet.py:19-49`) and to the denoiser head
(`bfcnn/model.py:267-275`), restricted to the family named by the north star:
``resnet_color_1xN_bn_16x3x3`` (two 3x3 16-channel convs per block, BN after
the second, ReLU after the first, linear block output, additive skip).

The same struct crosses the C ABI as ``bfcnn_arch`` (include/bfcnn_b200.h).
"""
from __future__ import annotations

import ctypes
import re
from dataclasses import dataclass
from typing import Dict, List, Tuple

BN_EPSILON = 1e-3    # reference bfcnn/constants.py:9  (DEFAULT_BN_EPSILON)
BN_MOMENTUM = 0.995  # reference bfcnn/constants.py:11 (DEFAULT_BN_MOMENTUM)


class CArch(ctypes.Structure):
    """ctypes image of ``bfcnn_arch`` (include/bfcnn_b200.h)."""
    _fields_ = [
        ("no_layers", ctypes.c_int32),
        ("base_kernel", ctypes.c_int32),
        ("filters", ctypes.c_int32),
        ("head_filters", ctypes.c_int32),
        ("in_channels", ctypes.c_int32),
        ("out_channels", ctypes.c_int32),
        ("bn_epsilon", ctypes.c_float),
        ("bn_momentum", ctypes.c_float),
    ]


@dataclass(frozen=True)
class Arch:
    no_layers: int = 6          # N residual blocks (backbone_resnet.py: no_layers)
    base_kernel: int = 3        # k0 of the base conv (backbone_resnet.py: kernel_size)
    filters: int = 16           # channels of every backbone conv
    head_filters: int = 32      # model.py:268 default "filters"
    in_channels: int = 3
    out_channels: int = 3
    bn_epsilon: float = BN_EPSILON
    bn_momentum: float = BN_MOMENTUM

    def __post_init__(self):
        if self.filters != 16:
            raise ValueError("only the 16-channel resnet family is on the hot path")
        if self.base_kernel not in (1, 3, 5, 7):
            raise ValueError("base_kernel must be an odd value in {1,3,5,7}")
        if not (0 <= self.no_layers <= 64):
            raise ValueError("no_layers must be in [0, 64]")
        if self.in_channels != 3 or self.out_channels != 3:
            raise ValueError("only colour (3-channel) models are on the hot path")
        if not (1 <= self.head_filters <= 64):
            raise ValueError("head_filters must be in [1, 64]")

    # ------------------------------------------------------------------
    @property
    def receptive_radius(self) -> int:
        """R = (k0-1)/2 + 2N."""
        return (self.base_kernel - 1) // 2 + 2 * self.no_layers

    def variable_shapes(self) -> List[Tuple[int, ...]]:
        """Shapes of ``hydra.variables`` in Keras order:
        base kernel, per block [W_a, W_b, gamma, moving_mean, moving_var], head W0, W1."""
        c, k0 = self.filters, self.base_kernel
        shapes: List[Tuple[int, ...]] = [(k0, k0, self.in_channels, c)]
        for _ in range(self.no_layers):
            shapes += [(3, 3, c, c), (3, 3, c, c), (c,), (c,), (c,)]
        shapes += [(1, 1, c, self.head_filters), (1, 1, self.head_filters, self.out_channels)]
        return shapes

    def trainable_mask(self) -> List[bool]:
        """True for variables that receive gradients (moving stats do not)."""
        mask = [True]
        for _ in range(self.no_layers):
            mask += [True, True, True, False, False]
        mask += [True, True]
        return mask

    def num_variables(self) -> int:
        return 3 + 5 * self.no_layers

    def num_weights(self) -> int:
        n = 0
        for s in self.variable_shapes():
            m = 1
            for d in s:
                m *= d
            n += m
        return n

    def num_trainable(self) -> int:
        n = 0
        for s, t in zip(self.variable_shapes(), self.trainable_mask()):
            if t:
                m = 1
                for d in s:
                    m *= d
                n += m
        return n

    def flops_per_pixel(self) -> int:
        """Algorithmic forward FLOPs/px, head un-collapsed."""
        c, k0, f = self.filters, self.base_kernel, self.head_filters
        return (2 * k0 * k0 * self.in_channels * c
                + self.no_layers * 2 * (2 * 9 * c * c)
                + 2 * c * f + 2 * f * self.out_channels)

    def to_c(self) -> CArch:
        return CArch(self.no_layers, self.base_kernel, self.filters, self.head_filters,
                     self.in_channels, self.out_channels, self.bn_epsilon, self.bn_momentum)


# ----------------------------------------------------------------------
def arch_from_variable_shapes(shapes: List[Tuple[int, ...]]) -> Arch:
    """Infer (k0, N, F) from the shape sequence of a checkpoint."""
    shapes = [tuple(int(d) for d in s) for s in shapes]
    if len(shapes) < 3 or (len(shapes) - 3) % 5 != 0:
        raise ValueError(f"variable count {len(shapes)} is not 3 + 5*N")
    n = (len(shapes) - 3) // 5
    b = shapes[0]
    if len(b) != 4 or b[0] != b[1] or b[2] != 3:
        raise ValueError(f"unexpected base kernel shape {b}")
    h0, h1 = shapes[-2], shapes[-1]
    if len(h0) != 4 or h0[:2] != (1, 1) or len(h1) != 4 or h1[:2] != (1, 1):
        raise ValueError(f"unexpected head shapes {h0} {h1}")
    arch = Arch(no_layers=n, base_kernel=b[0], filters=b[3], head_filters=h0[3],
                in_channels=b[2], out_channels=h1[3])
    if arch.variable_shapes() != shapes:
        raise ValueError("variable shapes do not match the resnet_color_1xN_bn_16x3x3 family")
    return arch


def arch_from_config(config: Dict) -> Arch:
    """Build an Arch from a reference pipeline config (the ``model`` section of
    ``pipeline.json``; keys as read by `bfcnn/model.py:58-66` and
    `bfcnn/backbone_resnet.py:19-49`)."""
    model = config.get("model", config)
    bb = model["backbone"]
    dn = model.get("denoiser", {})
    if bb.get("type", "resnet").strip().lower() != "resnet":
        raise ValueError("only type=resnet backbones are on the hot path")
    bk = list(bb.get("block_kernels", [3, 3]))
    bf = list(bb.get("block_filters", [16, 16]))
    if bk != [3, 3] or bf != [bb.get("filters", 16)] * 2:
        raise ValueError(f"unsupported block layout kernels={bk} filters={bf}")
    for flag in ("add_gates", "add_final_bn", "add_initial_bn", "add_concat_input",
                 "add_channelwise_scaling", "add_learnable_multiplier",
                 "add_mean_sigma_normalization", "add_gelu"):
        if bb.get(flag, False):
            raise ValueError(f"backbone option {flag} is not on the hot path")
    if bb.get("use_bias", False) or dn.get("use_bias", False):
        raise ValueError("bias terms are not on the hot path (bias-free models only)")
    if not bb.get("use_bn", True):
        raise ValueError("use_bn=False is not on the hot path")
    # Keys that change the arithmetic the kernels hard-code.  Most of them leave every variable SHAPE unchanged, so the
    # checkpoint/shape cross-check of load_model cannot catch them: reject anything but the reference defaults
    # (backbone_resnet.py:19-49, model.py:267-275) instead of computing something else silently.
    def _same(v, allowed):
        return (v.strip().lower() if isinstance(v, str) else v) in allowed

    def _need(section, cfg, key, allowed, default):
        v = cfg.get(key, default)
        if not _same(v, allowed):
            raise ValueError(f"{section} option {key}={v!r} is not on the hot path (the kernels implement {sorted(map(str, allowed))[0]!s})")

    _need("backbone", bb, "activation", {"relu"}, "relu")                  # ReLU after conv_a (backbone_blocks.py:174-178)
    _need("backbone", bb, "base_activation", {"linear"}, "linear")         # base conv and block outputs are linear (:178)
    _need("backbone", bb, "kernel_regularizer", {"l1"}, "l1")              # train.cu: L1(0.01) on every backbone kernel
    _need("backbone", bb, "dropout_rate", {-1, -1.0}, -1)
    _need("backbone", bb, "add_gradient_dropout", {False}, False)
    v = bb.get("block_activation", None)   # the last entry is overridden by base_activation (backbone_resnet.py:178)
    if v and (len(v) != len(bk) or [str(a).strip().lower() for a in v[:-1]] != ["relu"] * (len(bk) - 1)):
        raise ValueError(f"backbone option block_activation={v!r} is not on the hot path")
    v = bb.get("block_regularizer", None)
    if v and [str(a).strip().lower() for a in v] != ["l1"] * len(bk):
        raise ValueError(f"backbone option block_regularizer={v!r} is not on the hot path")
    v = bb.get("block_groups", None)
    if v and list(v) != [1] * len(bk):
        raise ValueError(f"backbone option block_groups={v!r} is not on the hot path")
    v = bb.get("block_depthwise", None)
    if v and list(v) != [-1] * len(bk):
        raise ValueError(f"backbone option block_depthwise={v!r} is not on the hot path")
    if bb.get("selector_params", None):
        raise ValueError("backbone option selector_params is not on the hot path")
    if list(bb.get("value_range", [0, 255])) != [0, 255]:
        raise ValueError("backbone option value_range must be [0, 255] (the normaliser is fused into the base conv)")
    # denoiser head (model.py:267-275): two LINEAR 1x1 convs without normalisation collapse into one [16,3] matrix
    _need("denoiser", dn, "activation", {"linear"}, "linear")
    _need("denoiser", dn, "kernel_regularizer", {"l2"}, "l2")              # train.cu: L2(0.01) on the head kernels
    _need("denoiser", dn, "use_bn", {False}, False)
    _need("denoiser", dn, "use_ln", {False}, False)
    ishape = bb.get("input_shape", ["?", "?", 3])
    return Arch(no_layers=int(bb["no_layers"]),
                base_kernel=int(bb.get("kernel_size", 3)),
                filters=int(bb.get("filters", 16)),
                head_filters=int(dn.get("filters", 32)),
                in_channels=int(ishape[-1]),
                out_channels=int(dn.get("output_channels", 3)))


_NAME_RE = re.compile(r"resnet_color_1x(\d+)_bn_16x3x3")


def arch_from_name(name: str) -> Arch:
    """``resnet_color_1x18_bn_16x3x3_256x256_l1_relu`` -> Arch(no_layers=18)."""
    m = _NAME_RE.search(name)
    if not m:
        raise ValueError(f"[{name}] is not a resnet_color_1xN_bn_16x3x3 model name")
    return Arch(no_layers=int(m.group(1)))


def default_pipeline_config(arch: Arch, name: str = "") -> Dict:
    """A pipeline.json body in the reference's schema
    (`bfcnn/configs/resnet_color_1x6_...json`) for this arch."""
    return {
        "model": {
            "backbone": {
                "type": "resnet", "input_shape": ["?", "?", arch.in_channels],
                "no_layers": arch.no_layers, "kernel_size": arch.base_kernel,
                "filters": arch.filters, "block_kernels": [3, 3],
                "block_filters": [arch.filters, arch.filters],
                "value_range": [0, 255], "activation": "relu", "use_bn": True,
                "use_bias": False, "kernel_regularizer": "l1",
                "kernel_initializer": "glorot_normal",
            },
            "denoiser": {
                "filters": arch.head_filters, "use_bias": False,
                "output_channels": arch.out_channels, "kernel_regularizer": "l2",
                "kernel_initializer": "glorot_normal",
            },
        },
        "loss": {"hinge": 0.5, "cutoff": 255.0, "mae_multiplier": 1.0,
                 "ssim_multiplier": 0.0, "mse_multiplier": 0.0, "regularization": 0.01},
        "dataset": {"batch_size": 32, "input_shape": [256, 256, 3],
                    "additional_noise": [5, 40], "multiplicative_noise": [0.05, 0.1],
                    "random_up_down": True, "random_left_right": True,
                    "value_range": [0, 255], "round_values": True},
        "name": name,
    }
