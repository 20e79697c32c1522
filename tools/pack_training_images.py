"""Pack the training images for tools/train_pretrained.py into data/train_images.npz (git-ignored; it travels to the GPU box
with the gpurun snapshot).  Sources: the natural images the reference snapshot carries for its own visual tests,
<reference>/images/test/megadepth/files (16 JPEG photographs, ~1600 x 1200, area-averaged 2x here: cleaner targets and a
quarter of the bytes) and <reference>/images/test/kitti/files (the 16 RGB frames, 1241 x 376).  The four 512 x 512 stock
images of <reference>/bfcnn/images are NOT included: they are the held-out evaluation set of tests/golden/natural_inputs.npz.

python tools/pack_training_images.py [/root/reference]"""
import os
import sys

import numpy as np
from PIL import Image

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "train_images.npz")
images = {}
d = os.path.join(ref, "images", "test", "megadepth", "files")
for f in sorted(os.listdir(d)):
    im = Image.open(os.path.join(d, f))
    if im.mode != "RGB":
        continue
    a = np.asarray(im, np.float32)
    h, w = (a.shape[0] // 2) * 2, (a.shape[1] // 2) * 2
    a = a[:h, :w].reshape(h // 2, 2, w // 2, 2, 3).mean(axis=(1, 3))
    images["megadepth_" + os.path.splitext(f)[0]] = np.clip(np.rint(a), 0, 255).astype(np.uint8)
d = os.path.join(ref, "images", "test", "kitti", "files")
for f in sorted(os.listdir(d)):
    im = Image.open(os.path.join(d, f))
    if im.mode != "RGB":
        continue
    images["kitti_" + os.path.splitext(f)[0]] = np.asarray(im, np.uint8)
os.makedirs(os.path.dirname(out), exist_ok=True)
np.savez_compressed(out, **images)
print(f"{len(images)} images, {sum(v.nbytes for v in images.values()) / 1e6:.1f} MB raw -> {out} ({os.path.getsize(out) / 1e6:.1f} MB)")
