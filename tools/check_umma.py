"""Quick parity + timing of the tcgen05 fused stack against the oracle and the mma.sync path."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import blind_image_denoising_b200 as bf
from oracle import bfcnn_oracle as O
import torch

for n_layers, shape in [(2, (1, 40, 200, 3)), (6, (2, 96, 80, 3)), (18, (1, 150, 300, 3))]:
    arch = bf.Arch(no_layers=n_layers)
    v = bf.synthetic_variables(arch, 0)
    x = np.random.default_rng(1).integers(0, 256, size=shape, dtype=np.uint8)
    yref, _ = O.denoise(v, x, pad_pow2=True)
    for prec in ("f16", "f16_mma_sync", "f16x3", "f16x3_mma_sync"):
        m = bf.Denoiser(arch, v, precision=prec)
        y = m(x, return_float=True)
        d = np.abs(y - yref)
        print(f"N={n_layers} {shape} {prec}: max-abs {d.max():.4f} mean-abs {d.mean():.5f} nan {np.isnan(y).sum()}", flush=True)
        m.close()

arch = bf.Arch(no_layers=18)
v = bf.synthetic_variables(arch, 0)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2, 2160, 3840, 3), dtype=np.uint8)).cuda()
out = torch.empty_like(x)
for prec in ("f16", "f16_mma_sync", "f16x3", "f16x3_mma_sync"):
    m = bf.Denoiser(arch, v, precision=prec, pad_pow2=False)
    for _ in range(2):
        m(x, out=out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        m(x, out=out)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{prec}: {dt*1e3:.2f} ms per 2 frames -> {2*2160*3840/1e6/dt:.0f} MP/s (stack {m.last_stack_ms():.2f} ms)", flush=True)
    res = out.clone() if prec in ("f16", "f16x3") else res
    if prec.endswith("mma_sync"):
        dd = (out.int() - res.int()).abs()
        print("u8 diff umma vs mma.sync: max", int(dd.max()), "frac>0", float((dd > 0).float().mean()))
    m.close()
