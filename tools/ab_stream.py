"""A/B of env-selected variants of the streaming stack in ONE process (same box, same clocks): python tools/ab_stream.py ENV v1,v2,... [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf
env, vals = sys.argv[1], sys.argv[2].split(",")
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
x = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
m = bf.synthetic_model(18, precision="f16", pad_pow2=False)
ref = None
for rnd in range(2):
    for v in vals:
        os.environ[env] = v
        for _ in range(2):
            m(x, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            m(x, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        same = "" if ref is None else f" identical-to-first={bool((out == ref).all())}"
        if ref is None:
            ref = out.clone()
        print(f"{env}={v}: {ms:.3f} ms per {frames} frames -> {frames*2160*3840/1e3/ms:.0f} MP/s{same}", flush=True)
