"""ncu target: one training-shape 3x3 conv on a chosen engine: python tools/prof_conv.py <engine> [relu]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
eng = int(sys.argv[1])
m = bf.synthetic_model(1)
lib = _native.load_library()
x = torch.randn(32, 256, 256, 16, device="cuda"); w = (torch.randn(3, 3, 16, 16, device="cuda") * 0.1)
out = torch.empty_like(x)
for _ in range(3):
    _native.check(lib.bfcnn_conv3x3(m.handle, x.data_ptr(), w.data_ptr(), out.data_ptr(), 32, 256, 256, eng, 1, None))
torch.cuda.synchronize()
print("ok")
