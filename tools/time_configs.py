"""Timing of BASELINE configs[0..2] on one GPU (device-resident, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
for name, nl, shape in [("cfg0 1x6 1x256x256", 6, (1, 256, 256, 3)), ("cfg1 1x12 64x256x256", 12, (64, 256, 256, 3)),
                        ("cfg2 1x18 4x2160x3840", 18, (4, 2160, 3840, 3))]:
    x = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x)
    for prec in ("f16", "f16x3", "fp32"):
        m = bf.synthetic_model(nl, precision=prec, pad_pow2=False)
        for _ in range(3):
            m(x, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20 if shape[0] * shape[1] < 100000 else 3
        e0.record()
        for _ in range(reps):
            m(x, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"{name:24s} {prec:6s} {ms:9.3f} ms  {shape[0]*shape[1]*shape[2]/1e6/(ms/1e3):9.1f} MP/s", flush=True)
        m.close()
