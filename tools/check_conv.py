import sys; sys.path.insert(0, ".")
import numpy as np, torch, ctypes
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
import torch.nn.functional as F
m = bf.synthetic_model(1)
lib = _native.load_library()
rng = np.random.default_rng(0)
for scale in (1.0, 1e-3):
    x = torch.tensor(rng.standard_normal((2, 50, 70, 16)) * scale, dtype=torch.float32).cuda()
    w = torch.tensor(rng.standard_normal((3, 3, 16, 16)) * 0.1, dtype=torch.float32).cuda()
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1)
    for eng in (0, 1, 2):
        out = torch.empty_like(x)
        _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), w.data_ptr(), out.data_ptr(), 2, 50, 70, eng, 0, None))
        torch.cuda.synchronize()
        d = (out.double() - ref).abs()
        print(f"scale {scale} engine {eng}: max err {float(d.max()):.3e} mean {float(d.mean()):.3e} (ref max {float(ref.abs().max()):.3f})")
for (n, hh, ww) in [(3, 36, 28), (2, 70, 66), (2, 24, 40), (1, 25, 62), (1, 26, 63), (1, 51, 125), (1, 7, 300), (2, 130, 127), (32, 256, 256)]:
    x = torch.tensor(rng.standard_normal((n, hh, ww, 16)), dtype=torch.float32).cuda()
    w = torch.tensor(rng.standard_normal((3, 3, 16, 16)) * 0.1, dtype=torch.float32).cuda()
    outs = []
    for eng in (0, 1, 2):
        out = torch.full_like(x, float("nan"))
        _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), w.data_ptr(), out.data_ptr(), n, hh, ww, eng, 1, None))
        torch.cuda.synchronize(); outs.append(out)
    for k in (1, 2):
        d = (outs[0] - outs[k]).abs()
        bad = torch.nonzero(~(d < 1e-4))
        print((n, hh, ww), f"engine {k} vs 0: max diff", float(torch.nan_to_num(d, nan=9e9).max()), "bad count", int(bad.shape[0]), bad[:5].tolist())

# timing at the training shape
x = torch.randn(32, 256, 256, 16, device="cuda"); w = (torch.randn(3, 3, 16, 16, device="cuda") * 0.1)
out = torch.empty_like(x)
for eng in (1, 2):
    for _ in range(3):
        _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), w.data_ptr(), out.data_ptr(), 32, 256, 256, eng, 1, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), w.data_ptr(), out.data_ptr(), 32, 256, 256, eng, 1, None))
    e1.record(); torch.cuda.synchronize()
    print(f"engine {eng}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per 32x256x256 conv")
