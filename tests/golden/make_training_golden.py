"""Generate tests/golden/training_golden.npz from the CPU oracles (fp64 autograd train step, numpy
corruption stream).  The reference itself cannot run here (TensorFlow absent); these vectors
pin the CUDA training path to the oracle's restatement on fixed seeded inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blind_image_denoising_b200 import Arch, synthetic_variables  # noqa: E402
from oracle import bfcnn_oracle as O  # noqa: E402
from oracle import corrupt_oracle as C  # noqa: E402

out = {}
rng = np.random.default_rng(2024)
x = rng.integers(0, 256, size=(6, 24, 20, 3), dtype=np.uint8)
clean, noisy = C.corrupt(x, seed=1234567890123, sample_offset=7, cfg=C.NoiseConfig(subsample=True))
out.update(corrupt_x=x, corrupt_seed=np.int64(1234567890123), corrupt_offset=np.int64(7), corrupt_clean=clean,
           corrupt_noisy=noisy)

arch = Arch(no_layers=3)
v = synthetic_variables(arch, 0)
xs = rng.integers(0, 256, size=(3, 20, 28, 3), dtype=np.uint8)
cl, no = C.corrupt(xs, seed=5, sample_offset=0, cfg=C.NoiseConfig())
r = O.train_step(v, cl, no, hinge=0.5, cutoff=255.0, mae_multiplier=1.0, mse_multiplier=0.5, regularization=0.01)
out.update(train_clean=cl, train_noisy=no, train_total=np.float64(r["total"]), train_mae=np.float64(r["mae"]),
           train_reg=np.float64(r["reg"]), train_denoiser_total=np.float64(r["denoiser_total"]),
           train_grads=np.concatenate([g.reshape(-1) for g in r["grads"]]).astype(np.float32))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "training_golden.npz"), **out)
print({k: getattr(v_, "shape", None) for k, v_ in out.items()})
