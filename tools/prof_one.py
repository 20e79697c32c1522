"""One short run for ncu: python tools/prof_one.py <prec> <n_layers> <n> <h> <w> [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf
prec, nl, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 2
m = bf.synthetic_model(nl, precision=prec)
x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
for _ in range(iters):
    m(x, out=out)
torch.cuda.synchronize()
print("ok", m.last_stack_ms())
