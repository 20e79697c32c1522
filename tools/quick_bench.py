"""Developer timing loop (not the contract bench): device-resident in/out, CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf

def run(n_layers, shape, prec, iters=5):
    m = bf.synthetic_model(n_layers, precision=prec)
    x = torch.randint(0, 256, shape, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(x)
    for _ in range(2): m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): m(x, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    mp = shape[0] * shape[1] * shape[2] / 1e6
    fl = bf.Arch(no_layers=n_layers).flops_per_pixel() * mp * 1e6
    print(f"N={n_layers} {shape} {prec}: {ms:.3f} ms  {mp/ms*1e3:.0f} MP/s  {fl/ms/1e9:.1f} TFLOP/s-alg", flush=True)
    m.close()

if __name__ == "__main__":
    for prec in ("f16", "f16x3", "fp32"):
        run(18, (1, 2160, 3840, 3), prec, iters=3 if prec == "fp32" else 5)
        run(12, (64, 256, 256, 3), prec, iters=3 if prec == "fp32" else 5)
        run(6, (1, 256, 256, 3), prec, iters=10)
