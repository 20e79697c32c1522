"""Run-to-run spread of the step gradient per wgrad kernel (is a spread the wgrad's or upstream of it?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import blind_image_denoising_b200 as bf
from blind_image_denoising_b200 import _native
from blind_image_denoising_b200.training import Trainer
arch = bf.Arch(no_layers=6)
ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
x = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(2, 100, 130, 3), dtype=np.uint8)).cuda()
for v1 in ("1", "0"):
    os.environ["BFCNN_WGRAD_V1"] = v1
    first, out = None, []
    for rep in range(6):
        t = Trainer(arch, bf.synthetic_variables(arch, 0), device=0, optimizer_config={"gradient_clipping_by_norm": 1.0})
        clean, noisy = t.prepare_data(x, ncfg, 0, 0)
        total, _, _, g = t.train_step_single_gpu(clean, noisy)
        torch.cuda.synchronize()
        g = g.clone()
        if first is None:
            first = g
        out.append("%.2e/%.8f" % (float((g - first).abs().max()), float(total)))
        t.close()
    print(f"V1={v1}: max |g - g_first| / loss per repeat: {out}", flush=True)
