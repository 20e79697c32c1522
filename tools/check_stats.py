import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200.training import Trainer
from oracle import bfcnn_oracle as O, corrupt_oracle as C
n_layers, shape = 6, (3, 36, 28, 3)
arch = bf.Arch(no_layers=n_layers); v = bf.synthetic_variables(arch, 0)
x = np.random.default_rng(n_layers).integers(0, 256, size=shape, dtype=np.uint8)
clean, noisy = C.corrupt(x, 11, 0, C.NoiseConfig())
loss = dict(hinge=0.5, cutoff=255.0, mae_multiplier=1.0, mse_multiplier=0.0, regularization=0.01)
ref = O.train_step(v, clean, noisy, **loss)
t = Trainer(arch, v, device=0, loss_config=dict(loss, ssim_multiplier=0.0))
t.train_step_single_gpu(torch.from_numpy(clean).cuda(), torch.from_numpy(noisy).cuda())
new = t.get_weights()
for i in range(n_layers):
    m_ref, v_ref = ref["new_moving"][i]
    bm = (new[1 + 5 * i + 3].astype(np.float64) - 0.995 * v[1 + 5 * i + 3].astype(np.float64)) / 0.005
    bm_ref = (m_ref - 0.995 * v[1 + 5 * i + 3].astype(np.float64)) / 0.005
    bv = (new[1 + 5 * i + 4].astype(np.float64) - 0.995 * v[1 + 5 * i + 4].astype(np.float64)) / 0.005
    bv_ref = (v_ref - 0.995 * v[1 + 5 * i + 4].astype(np.float64)) / 0.005
    print(f"block {i}: batch mean relerr {np.abs(bm - bm_ref).max() / np.abs(bm_ref).max():.2e}  var relerr {np.abs(bv / bv_ref - 1).max():.2e}")
