"""Where does a step's time go?  N back-to-back 4-frame steps with per-launch events on (no host sync in between); the
last step's kernel durations and inter-launch gaps are printed next to the loop time per step.
python tools/gap_probe.py [precision] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf

prec = sys.argv[1] if len(sys.argv) > 1 else "f16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
m = bf.synthetic_model(18, precision=prec)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(4, 2160, 3840, 3), dtype=np.uint8)).cuda()
out = torch.empty_like(x)
for timing in (False, True):
    m.set_kernel_timing(timing)
    for _ in range(3):
        m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        m(x, out=out)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"[{prec}] per-launch events {'on' if timing else 'off'}: {e0.elapsed_time(e1) / steps:.3f} ms per step on the device, "
          f"{t_issue / steps * 1e3:.3f} ms per step of host time to issue")
kt = m.kernel_times()
print("last step:", " ".join(f"{'gap' if k < 0 else ('base', 'pass', 'last')[k]}={v * 1e3:.0f}us" for k, v in kt))
print(f"kernels {sum(v for k, v in kt if k >= 0):.3f} ms, gaps {sum(v for k, v in kt if k < 0):.3f} ms")
