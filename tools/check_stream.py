"""Parity of the F16 engines (streaming vs oracle) on small shapes + 4K timing. BFCNN_UMMA_REGIONS=1 selects the region kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from oracle import bfcnn_oracle as O

for n_layers, shape in [(2, (1, 40, 200, 3)), (1, (1, 33, 70, 3)), (3, (1, 64, 64, 3)), (6, (2, 96, 80, 3)), (18, (1, 150, 300, 3)), (6, (1, 700, 260, 3))]:
    arch = bf.Arch(no_layers=n_layers)
    v = bf.synthetic_variables(arch, 0)
    x = np.random.default_rng(1).integers(0, 256, size=shape, dtype=np.uint8)
    yref, _ = O.denoise(v, x, pad_pow2=True)
    for prec in ("f16", "f16_mma_sync"):
        m = bf.Denoiser(arch, v, precision=prec)
        y = m(x, return_float=True)
        d = np.abs(y - yref)
        print(f"N={n_layers} {shape} {prec}: max-abs {d.max():.4f} mean-abs {d.mean():.5f} nan {np.isnan(y).sum()}", flush=True)
        m.close()
if len(sys.argv) > 1:
    frames = int(sys.argv[1])
    x = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x)
    m = bf.synthetic_model(18, precision="f16", pad_pow2=False)
    for _ in range(2):
        m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{ms:.3f} ms per {frames} frames -> {frames*2160*3840/1e3/ms:.0f} MP/s (stack {m.last_stack_ms():.2f} ms)", flush=True)
