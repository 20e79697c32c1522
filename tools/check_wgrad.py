"""The two wgrad kernels (BFCNN_WGRAD_V1=1: two 8-warp CTAs per SM; default: one 16-warp CTA with loader warps) in ONE
process: gradients of a training step against each other, run-to-run determinism, and step time.
python tools/check_wgrad.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
arch = bf.Arch(no_layers=6)
ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
for shape in [(3, 36, 28, 3), (2, 100, 130, 3), (32, 256, 256, 3)]:
    x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=shape, dtype=np.uint8)).cuda()
    grads = {}
    for v1 in ("1", "0", "0"):
        os.environ["BFCNN_WGRAD_V1"] = v1
        t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0})
        clean, noisy = t.prepare_data(x, ncfg, 0, 0)
        total, _, _, g = t.train_step_single_gpu(clean, noisy)
        torch.cuda.synchronize()
        g = g.clone()
        if v1 in grads and v1 == "0":
            print(f"{shape}: ws kernel run-to-run max diff {float((g - grads[v1]).abs().max()):.3e}", flush=True)
        grads[v1] = g
        if shape[0] == 32:
            for s in range(2):
                t.apply_grads(t.train_step_single_gpu(clean, noisy)[3])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(5):
                clean, noisy = t.prepare_data(x, ncfg, 0, 32 * s)
                t.apply_grads(t.train_step_single_gpu(clean, noisy)[3])
            e1.record(); torch.cuda.synchronize()
            print(f"V1={v1}: {e0.elapsed_time(e1) / 5:.3f} ms per step", flush=True)
        t.close()
    d = (grads["0"] - grads["1"]).abs().max() / grads["1"].abs().max()
    cos = torch.nn.functional.cosine_similarity(grads["0"].double(), grads["1"].double(), dim=0)
    print(f"{shape}: v2 vs v1 max diff / max |g| = {float(d):.3e}, cosine {float(cos):.9f}, nan {int(torch.isnan(grads['0']).sum())}", flush=True)
