"""GPU parity of the inference path against the CPU oracle (fp64), through the C ABI.

Gates (BASELINE.json north_star / SURVEY 8d):
  fp32, f16x3 : max-abs <= 0.5 and mean-abs <= 0.05 on the 0..255 scale (pre-round float);
                uint8 outputs differ by <= 1 LSB on < 1 % of values
  f16         : stated looser bound max-abs <= 2.0, mean-abs <= 0.25 (fp16 operands, fp32 accumulation, tcgen05)
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GATES = {"fp32": (0.5, 0.05), "f16x3": (0.5, 0.05), "f16": (2.0, 0.25)}
# what the kernels actually achieve (regression guard, tighter than the gate)
# (measured on B200, round 2: f16 at N = 18 max-abs 0.78-0.88 / mean-abs 0.070; the guard sits at ~2x that)
TIGHT = {"fp32": (0.02, 0.002), "f16x3": (0.02, 0.002), "f16": (1.6, 0.14)}


def _model(n_layers, **kw):
    import blind_image_denoising_b200 as bf
    return bf.synthetic_model(n_layers, seed=0, **kw)


def _oracle(n_layers, x, pad_pow2=True):
    from blind_image_denoising_b200 import Arch, synthetic_variables
    from oracle import bfcnn_oracle as O
    v = synthetic_variables(Arch(no_layers=n_layers), 0)
    return O.denoise(v, x, pad_pow2=pad_pow2)


def _check(y, yref, u8, u8ref, prec):
    d = np.abs(y.astype(np.float64) - yref)
    mx, mean = float(d.max()), float(d.mean())
    print(f"[{prec}] max-abs {mx:.5f} mean-abs {mean:.6f}")
    assert mx <= GATES[prec][0] and mean <= GATES[prec][1], (prec, mx, mean)
    assert mx <= TIGHT[prec][0] and mean <= TIGHT[prec][1], ("regression", prec, mx, mean)
    du = np.abs(u8.astype(np.int32) - u8ref.astype(np.int32))
    f16 = prec == "f16"
    assert du.max() <= (2 if f16 else 1)
    assert (du > 0).mean() < (0.25 if f16 else 0.01)


@pytest.mark.parametrize("prec", ["fp32", "f16x3", "f16"])
@pytest.mark.parametrize("n_layers,shape", [
    (1, (1, 32, 32, 3)),
    (6, (2, 64, 64, 3)),
    (6, (1, 53, 37, 3)),      # odd, non-pow2: exercises the pow2 canvas band
    (12, (1, 96, 80, 3)),
    (18, (1, 128, 128, 3)),
    (18, (1, 100, 150, 3)),   # several tiles + canvas band, deepest model
])
def test_parity_vs_oracle(native_lib, prec, n_layers, shape):
    rng = np.random.default_rng(7)
    x = rng.integers(0, 256, size=shape, dtype=np.uint8)
    m = _model(n_layers, precision=prec)
    yref, u8ref = _oracle(n_layers, x)
    y = m(x, return_float=True)
    u8 = m(x)
    assert u8.dtype == np.uint8 and u8.shape == x.shape
    _check(y, yref, u8, u8ref, prec)


@pytest.mark.parametrize("prec", ["fp32", "f16x3"])
def test_no_pad_pow2(native_lib, prec):
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, size=(1, 45, 70, 3), dtype=np.uint8)
    m = _model(6, precision=prec, pad_pow2=False)
    yref, u8ref = _oracle(6, x, pad_pow2=False)
    _check(m(x, return_float=True), yref, m(x), u8ref, prec)
    # and the pow2 canvas semantics really differ near the bottom/right border
    y2, _ = _oracle(6, x, pad_pow2=True)
    assert np.abs(y2 - yref).max() > 1.0


@pytest.mark.parametrize("prec", ["fp32", "f16x3", "f16"])
def test_edge_shapes(native_lib, prec):
    m = _model(6, precision=prec)
    out = m(np.zeros((0, 16, 16, 3), np.uint8))
    assert out.shape == (0, 16, 16, 3)
    rng = np.random.default_rng(11)
    for shape in [(1, 1, 1, 3), (3, 3, 5, 3), (1, 2, 130, 3), (1, 70, 1, 3)]:
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        yref, u8ref = _oracle(6, x)
        _check(m(x, return_float=True), yref, m(x), u8ref, prec)


def test_golden_fixtures(native_lib):
    """Committed oracle outputs (tests/golden/make_golden.py) reproduced by the CUDA path."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "inference_golden.npz"))
    for n_layers in (6, 12, 18):
        x = g[f"x_{n_layers}"]
        for prec in ("fp32", "f16x3", "f16"):
            m = _model(n_layers, precision=prec)
            _check(m(x, return_float=True), g[f"y_{n_layers}"].astype(np.float64), m(x), g[f"u8_{n_layers}"], prec)


def test_torch_device_tensors(native_lib):
    import torch
    rng = np.random.default_rng(5)
    x = rng.integers(0, 256, size=(2, 64, 48, 3), dtype=np.uint8)
    m = _model(6, precision="f16x3")
    ref = m(x)
    xt = torch.from_numpy(x).cuda()
    out = m(xt)
    assert out.is_cuda and out.dtype == torch.uint8
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), ref)
    outc = m(torch.from_numpy(x))
    assert not outc.is_cuda and np.array_equal(outc.numpy(), ref)


def test_load_model_pretrained_dirs(native_lib):
    """bfcnn.load_model(name) reads the TensorBundle shipped under pretrained/: the oracle on the same variables agrees."""
    import bfcnn
    import blind_image_denoising_b200 as bf
    from oracle import bfcnn_oracle as O
    name = "resnet_color_1x6_bn_16x3x3_256x256_l1_relu"
    assert name in bfcnn.models
    m = bfcnn.load_model(name, allow_synthetic=True)
    rng = np.random.default_rng(1)
    x = rng.integers(0, 256, size=(1, 256, 256, 3), dtype=np.uint8)
    y = m(x)
    _, u8ref = O.denoise(bf.load_variables(bfcnn.models[name]["directory"]), x, pad_pow2=True)
    du = np.abs(y.astype(np.int32) - u8ref.astype(np.int32))
    assert du.max() <= 1 and (du > 0).mean() < 0.01


@pytest.mark.parametrize("prec", ["f16x3", "f16"])
def test_full_size_properties(native_lib, prec):
    """BASELINE config 3 size (2160x3840): size-independent properties.
    (a) determinism; (b) locality: a crop with margin >= R reproduces the interior;
    (c) agreement with the FP32 path on the whole frame."""
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(1, 2160, 3840, 3), dtype=np.uint8)
    m = _model(18, precision=prec, pad_pow2=False)
    y = m(x)
    assert np.array_equal(y, m(x))
    R = 37
    y0, x0, hh, ww = 700, 1900, 200, 300
    crop = np.ascontiguousarray(x[:, y0 - R:y0 + hh + R, x0 - R:x0 + ww + R])
    yc = m(crop)[:, R:R + hh, R:R + ww]
    assert np.array_equal(yc, y[:, y0:y0 + hh, x0:x0 + ww])
    yf = m(x, precision="fp32")
    du = np.abs(yf.astype(np.int32) - y.astype(np.int32))
    assert du.max() <= (1 if prec == "f16x3" else 3)
    assert (du > 0).mean() < (0.01 if prec == "f16x3" else 0.25)


@pytest.mark.parametrize("pad_pow2", [False, True])
def test_row_strips_reassemble_bit_exact(native_lib, pad_pow2):
    """Multi-GPU inference shards a frame into row strips with receptive-field halos and NO collective
    (SURVEY 8e): the strips of ranks 0..world-1, concatenated, equal the whole-frame result bit for bit."""
    from blind_image_denoising_b200.distributed import denoise_rows
    m = _model(6, precision="f16", pad_pow2=pad_pow2)
    x = np.random.default_rng(5).integers(0, 256, size=(2, 150, 100, 3), dtype=np.uint8)
    whole = m(x)
    for world in (2, 3, 8):
        parts = [denoise_rows(m, x, r, world) for r in range(world)]
        assert [p[0] for p in parts][1:] == [p[1] for p in parts][:-1]
        assert np.array_equal(np.concatenate([p[2] for p in parts], axis=1), whole)
    m.close()


def test_host_pipeline_chunks_match_device_path(native_lib):
    """Host buffers are pipelined H2D | conv stack | D2H in chunks of >= ~4 MP (api.cu); the result must be the
    same bits as the single-shot device-buffer path, for uint8 and float outputs."""
    import torch
    m = _model(6, precision="f16", pad_pow2=False)
    x = np.random.default_rng(9).integers(0, 256, size=(5, 1100, 1300, 3), dtype=np.uint8)   # 1.43 MP each -> 3 chunks
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(m(x), m(xd).cpu().numpy())
    assert np.array_equal(m(x, return_float=True), m(xd, return_float=True).cpu().numpy())
    xp = torch.from_numpy(x).pin_memory()
    out = torch.empty_like(xp).pin_memory()
    m(xp, out=out)
    assert np.array_equal(out.numpy(), m(xd).cpu().numpy())
    m.close()


@pytest.mark.parametrize("k0", [1, 5, 7])
def test_base_kernel_sizes(native_lib, k0):
    """k0 is a load-time parameter (every in-tree resnet config uses 7, SURVEY 8 notation): all engines, odd shape."""
    import blind_image_denoising_b200 as bf
    from oracle import bfcnn_oracle as O
    arch = bf.Arch(no_layers=3, base_kernel=k0)
    v = bf.synthetic_variables(arch, 1)
    x = np.random.default_rng(k0).integers(0, 256, size=(2, 45, 70, 3), dtype=np.uint8)
    yref, u8ref = O.denoise(v, x, pad_pow2=True)
    for prec in ("fp32", "f16x3", "f16"):
        m = bf.Denoiser(arch, v, precision=prec)
        _check(m(x, return_float=True), yref, m(x), u8ref, prec)
        m.close()


@pytest.mark.parametrize("n_layers,shape", [
    (6, (96, 24, 40, 3)),      # many short strips: every CTA of the streaming kernel walks several segments
    (3, (40, 17, 300, 3)),     # odd block count (last pass has one block = two conv layers), 3 strips per image
    (6, (3, 700, 130, 3)),     # tall strips cut in the middle by the equal-share partition
])
def test_streaming_segments(native_lib, n_layers, shape):
    """fused_stream.cu: segment hand-over inside a CTA (pipeline drain, TMEM ring restart, X0 ring running on across
    segments), odd block counts and strip cuts -- against the oracle, and bit-identical to the same images run one by one."""
    rng = np.random.default_rng(11)
    x = rng.integers(0, 256, size=shape, dtype=np.uint8)
    m = _model(n_layers, precision="f16")
    yref, u8ref = _oracle(n_layers, x)
    y, u8 = m(x, return_float=True), m(x)
    _check(y, yref, u8, u8ref, "f16")
    for i in (0, shape[0] // 2, shape[0] - 1):
        assert np.array_equal(m(x[i:i + 1]), u8[i:i + 1])
    m.close()


def test_pipelined_denoiser_matches_single_instance(native_lib):
    """PipelinedDenoiser.map: two model instances on one GPU driven from two host threads (the copies of one batch
    overlap the conv stack of the other); results come back in input order and equal the single-instance call."""
    import blind_image_denoising_b200 as bf
    rng = np.random.default_rng(21)
    batches = [rng.integers(0, 256, size=(2, 300, 200 + 16 * i, 3), dtype=np.uint8) for i in range(7)]
    single = _model(6, precision="f16", pad_pow2=False)
    ref = [single(b) for b in batches]
    single.close()
    pipe = bf.PipelinedDenoiser(lambda: _model(6, precision="f16", pad_pow2=False), depth=2)
    got = list(pipe.map(batches))
    assert len(got) == len(ref) and all(np.array_equal(a, b) for a, b in zip(got, ref))
    assert np.array_equal(pipe(batches[0]), ref[0])
    pipe.close()


def test_random_shapes_streaming_vs_fp32(native_lib):
    """Random batch / height / width / depth / canvas settings (tools/fuzz_shapes.py runs more): the row-streaming tcgen05
    stacks against the layer-by-layer FP32 FFMA path -- no hang, uint8 within 1 LSB (f16x3) / 2 LSB (f16), deterministic."""
    rng = np.random.default_rng(2026)
    models = {}
    for _ in range(24):
        nl = int(rng.choice([1, 2, 3, 6]))
        n, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 260)), int(rng.integers(1, 300))
        pad = bool(rng.integers(0, 2))
        if nl not in models:
            models[nl] = _model(nl, precision="f16")
        a = models[nl]
        x = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
        ya, yx, yf = a(x, pad_pow2=pad), a(x, pad_pow2=pad, precision="f16x3"), a(x, pad_pow2=pad, precision="fp32")
        assert np.abs(ya.astype(int) - yf.astype(int)).max() <= 2, (nl, n, h, w, pad)
        assert np.abs(yx.astype(int) - yf.astype(int)).max() <= 1, (nl, n, h, w, pad)
        assert np.array_equal(ya, a(x, pad_pow2=pad)) and np.array_equal(yx, a(x, pad_pow2=pad, precision="f16x3"))
    for a in models.values():
        a.close()


def test_f16x3_streaming_matches_fp32(native_lib):
    """Precision f16x3 runs on the row-streaming kernel (fused_stream_x3.cu) and is FP32-grade: uint8 equal to the FP32
    FFMA path up to 1 LSB on < 1 % of the values, also across segment hand-overs (many small images) and odd shapes; and
    switching precisions on one handle (the stacks share their workspaces) does not disturb either."""
    rng = np.random.default_rng(31)
    for n_layers, shape in [(6, (40, 24, 40, 3)), (3, (2, 333, 217, 3)), (18, (1, 150, 300, 3))]:
        x = rng.integers(0, 256, size=shape, dtype=np.uint8)
        m = _model(n_layers, precision="f16x3")
        a = m(x)
        h16 = m(x, precision="f16")
        f = m(x, precision="fp32")
        d = np.abs(a.astype(np.int32) - f.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.01, (n_layers, shape, int(d.max()), float((d > 0).mean()))
        assert np.array_equal(a, m(x)) and np.array_equal(h16, m(x, precision="f16"))
        m.close()


def test_release_workspaces(native_lib):
    """Workspaces grow to the largest call and can be handed back; the next call allocates again and computes the same."""
    import torch
    m = _model(6, precision="f16")
    x = np.random.default_rng(3).integers(0, 256, size=(2, 700, 900, 3), dtype=np.uint8)
    ref = m(x)
    torch.cuda.synchronize()
    used = torch.cuda.mem_get_info()[0]
    m.release_workspaces()
    assert torch.cuda.mem_get_info()[0] > used + 2 * 700 * 900 * 32          # at least the two fp16 feature maps came back
    assert np.array_equal(m(x), ref)
    m.close()
