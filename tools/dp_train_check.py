"""Data-parallel `bfcnn.train_loop` on N GPUs of one node (one process per GPU, NCCL): every rank trains on its own shard of
the images, the flat gradient is all-reduced in apply_grads, rank 0 writes the checkpoints.  Checks after a short run: the
trainable variables are IDENTICAL on all ranks (they saw the same averaged gradients), the BN moving statistics differ
(local batch statistics, as the reference's micro-batch accumulation), the loss went down.

python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dp_train_check.py [steps]"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bfcnn  # noqa: E402
from blind_image_denoising_b200.arch import Arch, default_pipeline_config  # noqa: E402
from blind_image_denoising_b200.train_loop import Checkpoint  # noqa: E402
from blind_image_denoising_b200.weights import flatten_variables, gather_trainables  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
arch = Arch(no_layers=6)
cfg = default_pipeline_config(arch, "dp_check")
cfg["dataset"].update({"batch_size": 16, "input_shape": [128, 128, 3], "no_crops_per_image": 8,
                       "inputs": [{"synthetic": {"images": 16 * world, "height": 160, "width": 200, "seed": 3}}]})
# smooth synthetic "images" would be better targets than uniform noise; uniform noise still has a learnable mean
cfg["train"] = {"epochs": -1, "total_steps": steps, "gpu_batches_per_step": 1, "exact_accumulation": True, "checkpoints_to_keep": 1,
                "checkpoint_every": -1, "visualization_every": max(steps // 4, 1),
                "optimizer": {"type": "ADAM", "gradient_clipping_by_norm": 1.0,
                              "schedule": {"type": "exponential_decay", "config": {"decay_rate": 0.9, "decay_steps": 1000, "learning_rate": 0.001}}}}
ckpt_dir = os.path.join(tempfile.gettempdir(), f"dp_check_rank{rank}")   # rank 0's is the real one; the others only hold their metrics
final = {}


def grab(hydra):   # every rank ends with a model of its own: keep its variables
    final["variables"] = [np.array(v) for v in hydra.get_weights()]


bfcnn.train_loop(cfg, ckpt_dir, device=local, on_finish=grab)
flat = flatten_variables(arch, final["variables"])
train = torch.from_numpy(gather_trainables(arch, flat)).cuda()
allv = torch.from_numpy(np.ascontiguousarray(flat)).cuda()
g_train = [torch.empty_like(train) for _ in range(world)]
g_all = [torch.empty_like(allv) for _ in range(world)]
dist.all_gather(g_train, train)
dist.all_gather(g_all, allv)
same_trainables = all(bool(torch.equal(g_train[0], t)) for t in g_train[1:])
stats_differ = any(not bool(torch.equal(g_all[0], t)) for t in g_all[1:])
if rank == 0:
    step, epoch, variables = Checkpoint.read(Checkpoint(None, ckpt_dir).latest_checkpoint)
    import json
    rows = [json.loads(l) for l in open(os.path.join(ckpt_dir, "metrics.jsonl"))]
    print(f"world {world}: {step} steps, total loss {rows[0]['loss/total']:.2f} -> {rows[-1]['loss/total']:.2f}; trainable variables "
          f"identical on all ranks: {same_trainables}; BN moving statistics differ between ranks: {stats_differ}", flush=True)
    assert rows[-1]["loss/total"] < rows[0]["loss/total"] and same_trainables and (stats_differ or world == 1)
dist.destroy_process_group()
