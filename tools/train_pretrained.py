"""Train the three model directories `bfcnn.load_model` knows by name on a B200, with this repo's own training path
(`bfcnn.train_loop`: GPU corruption -> forward with batch statistics -> hinged MAE + L1/L2 regularisation -> backward ->
Adam), and write them in the reference's on-disk format (SavedModel variables bundle + pipeline.json).

The reference snapshot ships no resnet weights, so the directories held random ones; these are trained on the
32 natural images of data/train_images.npz (tools/pack_training_images.py) with the recipe of the reference's in-tree
resnet config (bfcnn/configs/resnet_color_1x6_...json: Adam 1e-3, exponential decay 0.9, global clip norm 1.0, additive
noise 5..40, multiplicative 0.05..0.1, flips), the loss of the `_l1_` models (hinged MAE), batch 32 of 256 x 256 crops.
Evaluation: the held-out 512 x 512 stock images at sigma = 20 (tests/golden/natural_inputs.npz, the recipe of the
reference's tests/bfcnn/test_pretrained.py:41-80): MAE / PSNR of the noisy input against the denoised output.

python tools/train_pretrained.py [steps] [layers ...]      # on the GPU box; results under gpurun_out/trained/
TRAIN_SSIM=1 adds the (1 - SSIM) term of the reference's default loss (ssim_multiplier 1.0, loss.py:171): results under
gpurun_out/trained_ssim/ (an end-to-end check of the SSIM forward / backward kernels, not shipped)."""
import json
import logging
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bfcnn  # noqa: E402
from blind_image_denoising_b200.arch import Arch, default_pipeline_config  # noqa: E402
from blind_image_denoising_b200.tensorbundle import write_model_variables  # noqa: E402
from blind_image_denoising_b200.train_loop import Checkpoint  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
layers = [int(a) for a in sys.argv[2:]] or [6, 12, 18]
logging.basicConfig(level=logging.WARNING)
z = np.load(os.path.join(ROOT, "data", "train_images.npz"))
images = [z[k] for k in z.files]
nat = np.load(os.path.join(ROOT, "tests", "golden", "natural_inputs.npz"))
with_ssim = os.environ.get("TRAIN_SSIM", "0") == "1"
out_root = os.path.join(ROOT, "gpurun_out", "trained_ssim" if with_ssim else "trained")


def psnr(a, b):
    return float(10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))


for n in layers:
    name = f"resnet_color_1x{n}_bn_16x3x3_256x256_l1_relu"
    arch = Arch(no_layers=n)
    cfg = default_pipeline_config(arch, name)
    cfg["dataset"].update({"no_crops_per_image": 16, "seed": n})
    if with_ssim:
        cfg["loss"]["ssim_multiplier"] = 1.0
    cfg["train"] = {"epochs": -1, "total_steps": steps, "gpu_batches_per_step": 1, "exact_accumulation": True,
                    "checkpoints_to_keep": 2, "checkpoint_every": steps // 4, "visualization_every": max(steps // 20, 1),
                    "optimizer": {"type": "ADAM", "gradient_clipping_by_norm": 1.0,
                                  "schedule": {"type": "exponential_decay",
                                               "config": {"decay_rate": 0.9, "decay_steps": max(steps // 10, 1), "learning_rate": 0.001}}}}
    ckpt_dir = os.path.join("/tmp", "train_" + name)
    shutil.rmtree(ckpt_dir, ignore_errors=True)
    t0 = time.time()
    bfcnn.train_loop(cfg, ckpt_dir, images=images)
    dt = time.time() - t0
    latest = Checkpoint(None, ckpt_dir).latest_checkpoint
    step, epoch, variables = Checkpoint.read(latest)
    d = os.path.join(out_root, name)
    os.makedirs(os.path.join(d, "saved_model", "variables"), exist_ok=True)
    write_model_variables(os.path.join(d, "saved_model", "variables"), variables)
    shutil.copy(os.path.join(ckpt_dir, "metrics.jsonl"), os.path.join(d, "metrics.jsonl"))
    with open(os.path.join(d, "pipeline.json"), "w") as f:
        json.dump(dict(default_pipeline_config(arch, name), weights="TRAINED (tools/train_pretrained.py)"), f, indent=2)
    # held-out evaluation through the drop-in callable
    model = bfcnn.load_model(d)
    den = model(nat["noisy"])
    ev = {"noisy_mae": float(np.abs(nat["noisy"].astype(int) - nat["clean"]).mean()),
          "denoised_mae": float(np.abs(den.astype(int) - nat["clean"]).mean()),
          "noisy_psnr": psnr(nat["noisy"], nat["clean"]), "denoised_psnr": psnr(den, nat["clean"])}
    model.close()
    pub = default_pipeline_config(arch, name)
    pub["weights"] = (f"TRAINED on one B200 by tools/train_pretrained.py: {step} steps of batch 32 x 256x256 crops of 32 natural images "
                      f"(reference images/test/megadepth + kitti), Adam 1e-3 exp. decay, hinged MAE + L1/L2 reg; NOT the reference's "
                      f"weights (its snapshot ships none, SURVEY F2). Held-out 512x512 stock images at sigma 20: MAE "
                      f"{ev['noisy_mae']:.2f} -> {ev['denoised_mae']:.2f}, PSNR {ev['noisy_psnr']:.2f} -> {ev['denoised_psnr']:.2f} dB")
    pub["training"] = {"steps": step, "seconds": round(dt, 1), "evaluation": ev, "train": cfg["train"], "dataset": cfg["dataset"]}
    with open(os.path.join(d, "pipeline.json"), "w") as f:
        json.dump(pub, f, indent=2)
    last = open(os.path.join(d, "metrics.jsonl")).read().strip().splitlines()[-1]
    print(f"{name}: {step} steps in {dt:.0f} s ({dt / max(step, 1) * 1e3:.2f} ms per step incl. host crops); {ev}; last metrics {last}", flush=True)
