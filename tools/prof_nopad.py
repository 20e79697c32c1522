"""ncu target: python tools/prof_nopad.py <prec> <n_layers> <frames> [iters] -- 4K frames, pad_pow2 off (the bench workload)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf
prec, nl, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
m = bf.synthetic_model(nl, precision=prec, pad_pow2=False)
x = torch.randint(0, 256, (n, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
for _ in range(iters):
    m(x, out=out)
torch.cuda.synchronize()
print("ok", m.last_stack_ms())
