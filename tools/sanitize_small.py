"""Small end-to-end run for compute-sanitizer (memcheck): every kernel family once on tiny inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
x = np.random.default_rng(0).integers(0, 256, size=(2, 70, 150, 3), dtype=np.uint8)
for prec in ("f16", "f16x3", "fp32"):
    m = bf.synthetic_model(4, precision=prec)
    y = m(x); yf = m(torch.from_numpy(x).cuda(), return_float=True)
    m.close()
arch = bf.Arch(no_layers=2)
t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0})
xs = torch.from_numpy(x[:, :40, :56].copy()).cuda()
clean, noisy = t.prepare_data(xs, _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 1, 1), 1, 0)
total, _, dl, g = t.train_step_single_gpu(clean, noisy)
t.apply_grads(g)
print("loss", total, t.denoiser_loss(clean, noisy))
t.close()
torch.cuda.synchronize()
print("sanitize run ok")
