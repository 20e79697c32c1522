"""Training step time (default engine) on batch 32 of 256x256: python tools/train_time.py [n_layers ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(32, 256, 256, 3), dtype=np.uint8)).cuda()
ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
for nl in [int(a) for a in sys.argv[1:]] or [6, 18]:
    arch = bf.Arch(no_layers=nl)
    t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0})
    losses = []
    def once(s):
        clean, noisy = t.prepare_data(x, ncfg, 0, s * 32)
        total, _, _, g = t.train_step_single_gpu(clean, noisy)
        t.apply_grads(g)
        return total
    for s in range(3):
        once(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(10):
        losses.append(float(once(3 + s)))
    e1.record(); torch.cuda.synchronize()
    print(f"1x{nl}: {e0.elapsed_time(e1) / 10:.3f} ms per step; losses {['%.6f' % l for l in losses[:3]]}", flush=True)
    t.close()
