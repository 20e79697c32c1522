"""Random-shape sweep of the streaming F16 / F16X3 stacks (against the layer-by-layer FP32 path) and of the tcgen05
training conv (against the FFMA conv): python tools/fuzz_shapes.py [cases] [seed]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
lib = _native.load_library()
models = {}
worst, worst3 = 0, 0
for t in range(cases):
    nl = int(rng.choice([1, 2, 3, 6, 12]))
    n, h, w = int(rng.integers(1, 5)), int(rng.integers(1, 320)), int(rng.integers(1, 420))
    pad = bool(rng.integers(0, 2))
    if nl not in models:
        models[nl] = bf.synthetic_model(nl, precision="f16")
    a = models[nl]
    x = rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    ya, yx, yf = a(x, pad_pow2=pad), a(x, pad_pow2=pad, precision="f16x3"), a(x, pad_pow2=pad, precision="fp32")
    d, d3 = np.abs(ya.astype(int) - yf.astype(int)), np.abs(yx.astype(int) - yf.astype(int))
    worst, worst3 = max(worst, int(d.max())), max(worst3, int(d3.max()))
    assert d.max() <= 3 and d3.max() <= 1, (nl, n, h, w, pad, int(d.max()), int(d3.max()))
    assert np.array_equal(ya, a(x, pad_pow2=pad))
print(f"inference: {cases} random shapes ok (max u8 difference to the FP32 path: f16 {worst}, f16x3 {worst3})", flush=True)
m = models[next(iter(models))]
for t in range(cases):
    n, h, w = int(rng.integers(1, 4)), int(rng.integers(1, 200)), int(rng.integers(1, 300))
    x = torch.tensor(rng.standard_normal((n, h, w, 16)), dtype=torch.float32).cuda()
    wt = torch.tensor(rng.standard_normal((3, 3, 16, 16)) * 0.1, dtype=torch.float32).cuda()
    o0, o2 = torch.full_like(x, float("nan")), torch.full_like(x, float("nan"))
    relu = int(rng.integers(0, 2))
    _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), wt.data_ptr(), o0.data_ptr(), n, h, w, 0, relu, None))
    _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), wt.data_ptr(), o2.data_ptr(), n, h, w, 2, relu, None))
    torch.cuda.synchronize()
    e = float((o0 - o2).abs().max())
    assert e <= 2e-5, (n, h, w, relu, e)
print(f"training conv: {cases} random shapes ok", flush=True)
