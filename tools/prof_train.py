"""One short training run for ncu: python tools/prof_train.py [n_layers] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
nl = int(sys.argv[1]) if len(sys.argv) > 1 else 6
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
arch = bf.Arch(no_layers=nl)
t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0})
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(32, 256, 256, 3), dtype=np.uint8)).cuda()
ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
for s in range(steps):
    clean, noisy = t.prepare_data(x, ncfg, 0, s * 32)
    total, _, _, g = t.train_step_single_gpu(clean, noisy)
    t.apply_grads(g)
torch.cuda.synchronize()
print("ok", total)
