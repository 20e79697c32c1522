"""Training step time per conv engine (1x6, batch 32 of 256x256): python tools/train_engines.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
arch = bf.Arch(no_layers=6)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(32, 256, 256, 3), dtype=np.uint8)).cuda()
ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
for eng in ("x3", "t5", "x3", "t5"):
    t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0}, conv_engine=eng)
    losses = []
    def once(s):
        clean, noisy = t.prepare_data(x, ncfg, 0, s * 32)
        total, _, _, g = t.train_step_single_gpu(clean, noisy)
        t.apply_grads(g)
        return total
    for s in range(2):
        once(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(5):
        losses.append(float(once(2 + s)))
    e1.record(); torch.cuda.synchronize()
    print(f"{eng}: {e0.elapsed_time(e1) / 5:.3f} ms per step; losses {['%.6f' % l for l in losses]}", flush=True)
    t.close()
