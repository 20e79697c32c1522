"""Generate tests/golden/inference_golden.npz from the CPU oracle (fp64).
The reference itself cannot run here (TensorFlow absent); these vectors pin the
CUDA path to the oracle's restatement on fixed seeded inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from blind_image_denoising_b200 import Arch, synthetic_variables  # noqa: E402
from oracle import bfcnn_oracle as O  # noqa: E402

out = {}
for n in (6, 12, 18):
    rng = np.random.default_rng(100 + n)
    x = rng.integers(0, 256, size=(1, 72, 88, 3), dtype=np.uint8)
    y, u8 = O.denoise(synthetic_variables(Arch(no_layers=n), 0), x, pad_pow2=True)
    out[f"x_{n}"], out[f"y_{n}"], out[f"u8_{n}"] = x, y.astype(np.float32), u8
np.savez_compressed(os.path.join(os.path.dirname(__file__), "inference_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
