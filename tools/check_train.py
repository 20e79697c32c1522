import sys, os; sys.path.insert(0, ".")
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200.training import Trainer
from blind_image_denoising_b200.weights import trainable_offsets
from oracle import bfcnn_oracle as O, corrupt_oracle as C
n_layers = 6
shape = (3, int(os.environ.get("HH", 36)), int(os.environ.get("WW", 28)), 3)
arch = bf.Arch(no_layers=n_layers); v = bf.synthetic_variables(arch, 0)
x = np.random.default_rng(n_layers).integers(0, 256, size=shape, dtype=np.uint8)
clean, noisy = C.corrupt(x, 11, 0, C.NoiseConfig())
loss = dict(hinge=0.5, cutoff=255.0, mae_multiplier=1.0, mse_multiplier=0.0, regularization=0.01)
ref = O.train_step(v, clean, noisy, **loss)
t = Trainer(arch, v, device=0, loss_config=dict(loss, ssim_multiplier=0.0))
total, _, _, g = t.train_step_single_gpu(torch.from_numpy(clean).cuda(), torch.from_numpy(noisy).cuda())
g = g.cpu().numpy().astype(np.float64)
refv = np.concatenate([a.reshape(-1) for a in ref["grads"]])
print("mode", os.environ.get("BFCNN_TRAIN_CONV", "x3"), "total", total, ref["total"])
for (o, n, to), r in zip(trainable_offsets(arch), ref["grads"]):
    rr = refv[to:to+n]; e = np.abs(g[to:to+n] - rr).max()
    print(f"  var @{to:6d} n={n:5d} max|g|={np.abs(rr).max():9.4f} relerr={e/np.abs(rr).max():.2e}")
