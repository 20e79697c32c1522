"""Generate tests/golden/natural_inputs.npz: the "second, natural input set" of SURVEY 8d.

Source: the four 512x512 stock images the reference evaluates on during training
(/root/reference/bfcnn/images/*_512x512.*, wired at train_loop.py:87-96); a 160x160 window of each (enough for several
strips of the 1x18 receptive field) with the noise recipe of /root/reference/tests/bfcnn/test_pretrained.py:41-56 at
sigma = 20: original + truncated normal, clip [0,255], round, uint8.  tf.random.truncated_normal(seed=0) itself cannot be
reproduced without TensorFlow; the draws come from the oracle's Philox sampler (oracle/corrupt_oracle.py), seed 0.
Runs only where /root/reference exists; the GPU box reads the committed .npz."""
import glob
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import corrupt_oracle as C  # noqa: E402

paths = sorted(glob.glob("/root/reference/bfcnn/images/*_512x512.*"))
assert len(paths) == 4, paths
clean, noisy = [], []
for k, p in enumerate(paths):
    img = np.asarray(Image.open(p).convert("RGB"), np.uint8)
    assert img.shape == (512, 512, 3), (p, img.shape)
    y0, x0 = 176 + 8 * k, 176 - 8 * k
    c = img[y0:y0 + 160, x0:x0 + 160]
    pix = np.arange(160 * 160, dtype=np.uint32).reshape(160, 160)
    z = np.stack([C.truncated_normal_det(pix, 3 + ch, k, 0) for ch in range(3)], -1)      # |z| < 2, as TF's sampler
    n = np.rint(np.clip(c.astype(np.float32) + np.float32(20.0) * z, 0, 255)).astype(np.uint8)
    clean.append(c); noisy.append(n)
out = os.path.join(os.path.dirname(__file__), "natural_inputs.npz")
np.savez_compressed(out, clean=np.stack(clean), noisy=np.stack(noisy), names=np.array([os.path.basename(p) for p in paths]))
print(out, os.path.getsize(out), np.stack(noisy).shape, float(np.abs(np.stack(noisy).astype(int) - np.stack(clean)).mean()))
