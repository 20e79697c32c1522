"""Parity of the f16x3 engines (streaming vs region, vs oracle) + 4K timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from oracle import bfcnn_oracle as O
for n_layers, shape in [(1, (1, 33, 70, 3)), (2, (1, 40, 200, 3)), (3, (2, 64, 64, 3)), (6, (2, 96, 80, 3)), (18, (1, 150, 300, 3)), (6, (1, 700, 260, 3))]:
    arch = bf.Arch(no_layers=n_layers)
    v = bf.synthetic_variables(arch, 0)
    x = np.random.default_rng(1).integers(0, 256, size=shape, dtype=np.uint8)
    yref, _ = O.denoise(v, x, pad_pow2=True)
    for reg in ("0", "1"):
        os.environ["BFCNN_X3_REGIONS"] = reg
        m = bf.Denoiser(arch, v, precision="f16x3")
        y = m(x, return_float=True)
        d = np.abs(y - yref)
        print(f"N={n_layers} {shape} regions={reg}: max-abs {d.max():.5f} mean-abs {d.mean():.6f} nan {np.isnan(y).sum()}", flush=True)
        m.close()
frames = 2
x = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
m = bf.synthetic_model(18, precision="f16x3", pad_pow2=False)
for reg in ("0", "1", "0", "1"):
    os.environ["BFCNN_X3_REGIONS"] = reg
    for _ in range(2):
        m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        m(x, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"regions={reg}: {ms:.3f} ms per {frames} frames -> {frames*2160*3840/1e3/ms:.0f} MP/s", flush=True)
