"""A/B of an environment switch of the library in ONE process on ONE box (boxes differ by up to 10 %):
python tools/ab_env.py VAR [precision] [rounds] [diff]  -- alternates VAR=0 / VAR=1, 8 steps of 4 4K frames each.
With `diff` the two settings may give different results (another arithmetic): the largest uint8 difference is printed."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf

var = sys.argv[1]
prec = sys.argv[2] if len(sys.argv) > 2 else "f16"
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
allow_diff = len(sys.argv) > 4 and sys.argv[4] == "diff"
maxdiff = 0
m = bf.synthetic_model(18, precision=prec)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(4, 2160, 3840, 3), dtype=np.uint8)).cuda()
out = torch.empty_like(x)
ref = None
res = {"0": [], "1": []}
for r in range(rounds):
    for val in ("0", "1"):
        os.environ[var] = val
        for _ in range(2):
            m(x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            m(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[val].append(e0.elapsed_time(e1) / 8)
        if ref is None:
            ref = out.clone()
        if allow_diff:
            maxdiff = max(maxdiff, int((out.to(torch.int16) - ref.to(torch.int16)).abs().max()))
        else:
            assert torch.equal(out, ref), f"{var}={val} changes the result"
print(f"[{prec}] {var}=0: " + " ".join(f"{v:.3f}" for v in res["0"]) + f"  |  {var}=1: " + " ".join(f"{v:.3f}" for v in res["1"]) + "   (ms per step)" + (f"  max uint8 difference {maxdiff}" if allow_diff else ""))
