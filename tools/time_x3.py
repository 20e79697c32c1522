"""Timing of the fp32-grade tcgen05 stack (f16x3) on 4K frames."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
for prec in ("f16x3", "f16"):
    m = bf.synthetic_model(18, precision=prec, pad_pow2=False)
    for _ in range(2):
        m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        m(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(f"{prec}: {ms:.3f} ms per {frames} frames -> {frames*2160*3840/1e3/ms:.0f} MP/s", flush=True)
    m.close()
