"""every other `type: "resnet"` configuration (3-conv blocks with a depthwise middle, grouped convs, initial /
final BatchNormalization, ChannelwiseMultiplier / Multiplier).  CPU: the config parser, the Keras variable order, and the
host-side folding of normalisations / multipliers into per-conv (scale, bias), evaluated with torch and held against the
un-folded oracle.  GPU: the FP32 layer kernels through `GenericDenoiser` / `bfcnn.load_model` against the same oracle."""
import json

import numpy as np
import pytest

import blind_image_denoising_b200 as bf
from blind_image_denoising_b200.generic import ConvSpec, fold_layers, initial_variables, spec_from_config

DEPTHWISE = "resnet_color_1x6_bn_32x128x32_1x3x1_128x128_depthwise_l1_relu"


def _cfg(**bb):
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT[DEPTHWISE]))
    cfg["model"]["backbone"].update(bb)
    return cfg


CASES = {
    "depthwise_1x3x1": _cfg(no_layers=2),                                            # the reference's in-tree resnet config
    "all_extras": _cfg(no_layers=2, kernel_size=3, filters=8, block_kernels=[3, 3], block_filters=[8, 8], block_depthwise=[],
                       block_groups=[], block_activation=[], block_regularizer=[], add_initial_bn=True, add_final_bn=True,
                       add_channelwise_scaling=True, add_learnable_multiplier=True),
    "one_conv_blocks": _cfg(no_layers=3, kernel_size=5, filters=12, block_kernels=[3], block_filters=[12], block_depthwise=[],
                            block_groups=[], block_activation=[], block_regularizer=[]),
    "grouped_no_bn": _cfg(no_layers=1, kernel_size=1, filters=16, block_kernels=[3, 1], block_filters=[32, 16], block_depthwise=[],
                          block_groups=[4, 2], block_activation=["relu", "relu"], block_regularizer=[], use_bn=False),
}


def test_spec_and_variable_order():
    spec = spec_from_config(bf.CONFIGS_DICT[DEPTHWISE])
    assert spec.no_layers == 6 and spec.base == ConvSpec(7, 3, 32) and spec.receptive_radius == 3 + 6
    assert [c.kernel_shape() for c in spec.block] == [(1, 1, 32, 32), (3, 3, 32, 4), (1, 1, 64, 32)]   # Keras kernel shapes
    shapes = spec.variable_shapes()
    # base; per block: conv1, depthwise + BN(128), conv3 + BN(32); head
    assert shapes[:9] == [(7, 7, 3, 32), (1, 1, 32, 32), (3, 3, 32, 4), (128,), (128,), (128,), (1, 1, 64, 32), (32,), (32,)]
    assert len(shapes) == 1 + 6 * 9 + 2 and shapes[-2:] == [(1, 1, 32, 32), (1, 1, 32, 3)]
    extras = spec_from_config(CASES["all_extras"]).variable_shapes()
    assert extras[:4] == [(3, 3, 3, 8), (8,), (8,), (8,)]                                   # initial BN after the base conv
    assert extras[4:13] == [(3, 3, 8, 8), (3, 3, 8, 8), (8,), (8,), (8,), (8,), (1,), (1,), (1,)]   # block: convs, BN, channelwise (w0, w1), multiplier
    assert extras[-9:] == [(8,), (8,), (8,), (8,), (1,), (1,), (1,), (1, 1, 8, 32), (1, 1, 32, 3)]
    # the fast family parses to the same structure as Arch
    fast = spec_from_config(bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"])
    assert [tuple(s) for s in fast.variable_shapes()] == [tuple(s) for s in bf.Arch(no_layers=6).variable_shapes()]
    for bad in (dict(add_gates=True), dict(block_activation=["gelu", "relu", "linear"]), dict(block_kernels=[3, 3, 3, 3], block_filters=[8] * 4),
                dict(block_filters=[32, 128, 16]), dict(use_bias=True), dict(dropout_rate=0.5), dict(block_kernels=[1, 2, 1])):
        with pytest.raises(ValueError):
            spec_from_config(_cfg(**bad))


def _run_folded_with_torch(spec, variables, x_u8):
    """The folded layer list evaluated with torch fp64 (what csrc/generic.cu computes in fp32)."""
    import torch
    import torch.nn.functional as F
    from oracle.bfcnn_oracle import next_pow2
    n, h, w, _ = x_u8.shape
    canvas = np.zeros((n, next_pow2(h), next_pow2(w), 3))
    canvas[:, :h, :w] = x_u8
    cur = torch.as_tensor(canvas).permute(0, 3, 1, 2) / 255.0 - 0.5
    skip = None
    for ly in fold_layers(spec, variables):
        c = ly.conv
        k = torch.as_tensor(np.asarray(ly.weights, np.float64))
        if ly.block_start:
            skip = cur
        if c.depth_multiplier > 0:
            y = F.conv2d(cur, k.permute(2, 3, 0, 1).reshape(c.cout, 1, c.kernel, c.kernel), padding=(c.kernel - 1) // 2, groups=c.cin)
        else:
            y = F.conv2d(cur, k.permute(3, 2, 0, 1).contiguous(), padding=(c.kernel - 1) // 2, groups=c.groups)
        if ly.scale is not None:
            y = y * torch.as_tensor(ly.scale).view(1, -1, 1, 1)
        if ly.bias is not None:
            y = y + torch.as_tensor(ly.bias).view(1, -1, 1, 1)
        if c.relu:
            y = torch.relu(y)
        if ly.residual:
            y = y + skip
        cur = y
    y = (torch.clamp(torch.tanh(2.0 * cur) * 0.51, -0.5, 0.5) + 0.5) * 255.0
    return y.permute(0, 2, 3, 1).numpy()[:, :h, :w]


@pytest.mark.parametrize("case", sorted(CASES))
def test_folding_matches_the_unfolded_oracle(case):
    from oracle.generic_oracle import denoise_generic
    cfg = CASES[case]
    spec = spec_from_config(cfg)
    v = initial_variables(spec, seed=3)
    x = np.random.default_rng(1).integers(0, 256, size=(2, 21, 13, 3), dtype=np.uint8)
    yref, _ = denoise_generic(cfg["model"], v, x, pad_pow2=True)
    y = _run_folded_with_torch(spec, v, x)
    assert np.abs(y - yref).max() <= 1e-8 and 5.0 < yref.std()          # folding is exact; the output is not saturated
    with pytest.raises(ValueError):
        fold_layers(spec, v[:-1])


def test_generic_oracle_agrees_with_the_16x3x3_oracle():
    """On the fast family the general restatement and the dedicated one are the same function."""
    from oracle import bfcnn_oracle as O
    from oracle.generic_oracle import denoise_generic
    cfg = bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]
    v = bf.synthetic_variables(bf.Arch(no_layers=6), 0)
    x = np.random.default_rng(0).integers(0, 256, size=(1, 19, 30, 3), dtype=np.uint8)
    a, _ = denoise_generic(cfg["model"], v, x)
    b, _ = O.denoise(v, x)
    assert np.abs(a - b).max() <= 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_generic_denoiser_matches_oracle(native_lib, case):
    from oracle.generic_oracle import denoise_generic
    cfg = CASES[case]
    spec = spec_from_config(cfg)
    v = initial_variables(spec, seed=3)
    m = bf.GenericDenoiser(spec, v)
    rng = np.random.default_rng(2)
    for shape in [(2, 45, 70, 3), (1, 64, 64, 3), (1, 1, 1, 3)]:
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        yref, u8ref = denoise_generic(cfg["model"], v, x, pad_pow2=True)
        y, u8 = m(x, return_float=True), m(x)
        d = np.abs(y.astype(np.float64) - yref)
        print(f"[{case}] {shape}: max-abs {d.max():.5f} mean-abs {d.mean():.6f}")
        assert d.max() <= 0.5 and d.mean() <= 0.05 and d.max() <= 0.02          # the fp32 gate, and what FP32 FFMA achieves
        du = np.abs(u8.astype(int) - u8ref.astype(int))
        assert u8.dtype == np.uint8 and du.max() <= 1 and (du > 0).mean() < 0.01
    assert m(np.zeros((0, 8, 8, 3), np.uint8)).shape == (0, 8, 8, 3)
    x = rng.integers(0, 256, size=(1, 33, 20, 3), dtype=np.uint8)
    yref, _ = denoise_generic(cfg["model"], v, x, pad_pow2=False)
    assert np.abs(m(x, return_float=True, pad_pow2=False) - yref).max() <= 0.02
    import torch
    xt = torch.from_numpy(x).cuda()
    assert m(xt).is_cuda and np.array_equal(m(xt).cpu().numpy(), m(x))
    with pytest.raises(ValueError):
        m(x, precision="f16")
    m.close()


@pytest.mark.gpu
def test_load_model_dispatches_on_the_checkpoint(native_lib, tmp_path):
    """A model directory holding the depthwise config + its variables loads through bfcnn.load_model (generic path); the
    16x3x3 family through the same generic kernels equals the tcgen05 / FFMA stacks of `Denoiser`."""
    import bfcnn
    from oracle.generic_oracle import denoise_generic
    from blind_image_denoising_b200.tensorbundle import write_model_variables
    cfg = json.loads(json.dumps(bf.CONFIGS_DICT[DEPTHWISE]))
    cfg["model"]["backbone"]["no_layers"] = 2
    spec = spec_from_config(cfg)
    v = initial_variables(spec, seed=5)
    d = tmp_path / "depthwise_model"
    write_model_variables(str(d / "saved_model" / "variables"), v)
    (d / "pipeline.json").write_text(json.dumps(cfg))
    m = bfcnn.load_model(str(d))
    assert isinstance(m, bf.GenericDenoiser)
    x = np.random.default_rng(0).integers(0, 256, size=(1, 40, 52, 3), dtype=np.uint8)
    _, u8ref = denoise_generic(cfg["model"], v, x)
    assert np.abs(m(x).astype(int) - u8ref.astype(int)).max() <= 1
    fast_cfg = bf.CONFIGS_DICT["resnet_color_1x6_bn_16x3x3_256x256_l1_relu"]
    fv = bf.synthetic_variables(bf.Arch(no_layers=6), 0)
    g = bf.GenericDenoiser(spec_from_config(fast_cfg), fv)
    f = bf.Denoiser(bf.Arch(no_layers=6), fv, precision="fp32")
    assert np.abs(g(x, return_float=True) - f(x, return_float=True)).max() <= 1e-3
    g.close(); f.close(); m.close()
