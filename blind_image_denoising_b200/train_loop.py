"""`train_loop(pipeline_config_path, checkpoint_directory, weights_dir=None)` -- the reference's training entry point
(reference bfcnn/train_loop.py:40-601) on the B200 path.

What is kept, line for line in meaning: config handling and `pipeline.json` copy (:63-78), dataset -> corruption
(:81, dataset.py:120-238), loss / optimizer builders (:98-106), model_builder + (step, epoch, model) checkpoints with
`checkpoints_to_keep` / `checkpoint_every`, restore-latest-or-load-`weights_dir` (:146-213), the epoch / total_steps loop
with gradient accumulation over `gpu_batches_per_step` micro-batches INCLUDING the reference's counter behaviour
(:343-348,404-437: the gradients of k+1 micro-batches are summed and scaled by 1/k), the end-of-epoch checkpoint (:598).
What is not: TensorBoard summaries, weight / gradient plots and the evaluation images (:439-559) -- observability, SURVEY 2;
the scalars they would log go to `<checkpoint_directory>/metrics.jsonl` instead, read back from the device only every
`visualization_every` steps so that steps chain on the stream without a host round trip.

All arithmetic runs in libbfcnn_b200.so through `Trainer`; under torch.distributed (one process per GPU) every rank
feeds its own shard of the images, the flat gradient is all-reduced in `apply_grads`, rank 0 writes the checkpoints.
"""
from __future__ import annotations

import json
import logging
import os
import re
import time
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from .model import model_builder
from .optimizer import deep_supervision_schedule_builder, optimizer_builder
from .tensorbundle import read_bundle, write_bundle
from .training import (dataset_builder, loss_function_builder, MAE_LOSS_STR, REGULARIZATION_LOSS_STR, SSIM_LOSS_STR,
                       TOTAL_LOSS_STR)

logger = logging.getLogger("bfcnn_b200")

CONFIG_PATH_STR = "pipeline.json"             # reference bfcnn/constants.py
MODEL_HYDRA_DEFAULT_NAME_STR = "model_hydra"  # reference: "model_hydra.keras"; here a directory (HydraModel.save)
_CKPT_RE = re.compile(r"^ckpt-(\d+)\.index$")


def load_config(config: Union[str, Dict, Path]) -> Dict:
    """utilities.py:59-83."""
    if config is None:
        raise ValueError("config should not be empty")
    if isinstance(config, dict):
        return config
    if isinstance(config, (str, Path)):
        if not os.path.isfile(str(config)):
            raise ValueError("configuration path [{0}] is not valid".format(str(config)))
        with open(str(config), "r") as f:
            return json.load(f)
    raise ValueError("don't know how to handle config [{0}]".format(config))


def save_config(config: Dict, filename: Union[str, Path]) -> None:
    """utilities.py:712-732."""
    with open(str(filename), "w") as f:
        json.dump(config, f, indent=4)


# --------------------------------------------------------------------------------------
# (step, epoch, model) checkpoints: utilities.py:691-706 + tf.train.CheckpointManager of train_loop.py:158-181
# --------------------------------------------------------------------------------------
class Checkpoint:
    """`tf.train.Checkpoint(step, epoch, model)` + `CheckpointManager(checkpoint_name="ckpt", max_to_keep)`.

    On disk: `ckpt-<n>.index` / `ckpt-<n>.data-00000-of-00001` TensorBundles (int64 `step`, `epoch`; float32
    `model/variables/<i>` in `hydra.variables` order, keys with TF's `/.ATTRIBUTES/VARIABLE_VALUE` suffix) and the
    `checkpoint` state file TF's `latest_checkpoint` reads.  `tf.train.load_checkpoint(prefix).get_tensor(key)` reads the
    tensors by name; the object-graph proto an object-based `ckpt.restore` needs is NOT written."""

    SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"

    def __init__(self, model, directory: Union[str, Path], max_to_keep: int = 3, checkpoint_name: str = "ckpt"):
        self.model = model
        self.step = 0
        self.epoch = 0
        self.directory = str(directory)
        self.max_to_keep = int(max_to_keep) if max_to_keep else 0
        self.name = checkpoint_name
        self._saves = [n for n in self._existing()]

    def _existing(self) -> List[int]:
        if not os.path.isdir(self.directory):
            return []
        out = []
        for f in os.listdir(self.directory):
            m = _CKPT_RE.match(f)
            if m:
                out.append(int(m.group(1)))
        return sorted(out)

    @property
    def latest_checkpoint(self) -> Optional[str]:
        state = os.path.join(self.directory, "checkpoint")
        if os.path.isfile(state):
            m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', open(state).read(), re.M)
            if m and os.path.isfile(os.path.join(self.directory, m.group(1) + ".index")):
                return os.path.join(self.directory, m.group(1))
        saves = self._existing()
        return os.path.join(self.directory, f"{self.name}-{saves[-1]}") if saves else None

    def save(self) -> str:
        n = (self._saves[-1] + 1) if self._saves else 1
        prefix = os.path.join(self.directory, f"{self.name}-{n}")
        tensors = {"step" + self.SUFFIX: np.int64(self.step), "epoch" + self.SUFFIX: np.int64(self.epoch)}
        for i, v in enumerate(self.model.get_weights()):
            tensors[f"model/variables/{i}{self.SUFFIX}"] = v
        write_bundle(prefix, tensors)
        self._saves.append(n)
        while self.max_to_keep > 0 and len(self._saves) > self.max_to_keep:
            old = self._saves.pop(0)
            for ext in (".index", ".data-00000-of-00001"):
                try:
                    os.remove(os.path.join(self.directory, f"{self.name}-{old}{ext}"))
                except OSError:
                    pass
        with open(os.path.join(self.directory, "checkpoint"), "w") as f:
            f.write(f'model_checkpoint_path: "{self.name}-{n}"\n')
            for k in self._saves:
                f.write(f'all_model_checkpoint_paths: "{self.name}-{k}"\n')
        return prefix

    @classmethod
    def read(cls, prefix: str) -> Tuple[int, int, List[np.ndarray]]:
        t = read_bundle(prefix)
        step = int(np.asarray(t["step" + cls.SUFFIX]).reshape(-1)[0])
        epoch = int(np.asarray(t["epoch" + cls.SUFFIX]).reshape(-1)[0])
        idx = sorted((int(m.group(1)), v) for k, v in t.items()
                     for m in [re.match(r"^model/variables/(\d+)/", k)] if m)
        return step, epoch, [v for _, v in idx]

    def restore(self, prefix: str) -> "Checkpoint":
        step, epoch, variables = self.read(prefix)
        self.model.set_weights(variables)
        self.step, self.epoch = step, epoch
        return self


def create_checkpoint(model=None, path: Union[str, Path, None] = None, directory: Union[str, Path, None] = None,
                      max_to_keep: int = 3) -> Checkpoint:
    """utilities.py:691-706: a (step, epoch, model) checkpoint; if `path` holds checkpoints the latest one is restored."""
    ckpt = Checkpoint(model, directory if directory is not None else (path or "."), max_to_keep=max_to_keep)
    if path is not None and os.path.isdir(str(path)):
        latest = Checkpoint(model, path).latest_checkpoint
        if latest is not None:
            ckpt.restore(latest)
    return ckpt


def _load_weights_dir(weights_dir: str) -> List[np.ndarray]:
    """train_loop.py:184-211: weights from another run -- its latest checkpoint, or a model directory
    (`saved_model/variables`, `denoiser/variables`, `model_hydra/variables`, `.npz`)."""
    latest = Checkpoint(None, weights_dir).latest_checkpoint
    if latest is not None:
        return Checkpoint.read(latest)[2]
    from . import load_variables
    for sub in ("", MODEL_HYDRA_DEFAULT_NAME_STR):
        try:
            return load_variables(os.path.join(weights_dir, sub) if sub else weights_dir)
        except ValueError:
            continue
    raise ValueError(f"no checkpoint or variables found in [{weights_dir}]")


# --------------------------------------------------------------------------------------
def train_loop(pipeline_config_path: Union[str, Dict, Path],
               checkpoint_directory: Union[str, Path],
               weights_dir: Union[str, Path] = None,
               *, images: Optional[Sequence[np.ndarray]] = None, device: Optional[int] = None, on_finish=None):
    """Trains a blind image denoiser (reference train_loop.py:40-601).

    :param pipeline_config_path: filepath to the configuration (or the dict itself)
    :param checkpoint_directory: directory to save checkpoints into
    :param weights_dir: directory to load weights from
    :param images: (extension) in-memory uint8 images [H,W,3] instead of `dataset.inputs[*].directory`
    :param device: (extension) CUDA device ordinal; default LOCAL_RANK or 0
    :param on_finish: (extension) callable(hydra) run on every rank after the last step, before the model is closed
    :return:
    """
    import torch
    import torch.distributed as dist

    # --- load configuration
    config = load_config(pipeline_config_path)
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))

    # --- create model_dir if not exist
    if not os.path.isdir(str(checkpoint_directory)):
        Path(str(checkpoint_directory)).mkdir(parents=True, exist_ok=True)
        if not os.path.isdir(str(checkpoint_directory)):
            raise ValueError("Model directory [{0}] is not valid".format(checkpoint_directory))

    # --- save configuration into path, makes it easier to compare afterwards
    if rank == 0:
        save_config(config=config, filename=os.path.join(str(checkpoint_directory), CONFIG_PATH_STR))

    # --- build dataset (file list / crops on the host, corruption on the GPU)
    dataset = dataset_builder(config["dataset"], images=images, rank=rank, world=world)
    batch_size = dataset.batch_size
    if dataset.training is None:
        raise ValueError("don't know how to handle non directory datasets")   # dataset.py:253

    # --- train configuration
    train_config = config["train"]
    epochs = int(train_config["epochs"])
    gpu_batches_per_step = int(train_config.get("gpu_batches_per_step", 1))
    if gpu_batches_per_step <= 0:
        raise ValueError("gpu_batches_per_step must be > 0")
    checkpoints_to_keep = train_config.get("checkpoints_to_keep", 3)
    checkpoint_every = int(train_config.get("checkpoint_every", -1))
    visualization_every = int(train_config.get("visualization_every", 1000))
    total_steps = int(train_config.get("total_steps", -1))
    # train_loop.py:421-437 sums k+1 micro-batches before an update and scales by 1/k; "exact_accumulation": true (an
    # extension) applies after exactly k
    exact_accumulation = bool(train_config.get("exact_accumulation", False))

    # --- build optimizer (raises for anything but Adam: optimizer.py:165 defaults to RMSprop, which is not on this path)
    optimizer, lr_schedule = optimizer_builder(config=train_config["optimizer"])

    # --- build the hydra model, its trainer (device copy of the variables) and the loss
    config["model"]["batch_size"] = batch_size
    models = model_builder(config=config["model"], device=device)
    hydra = models.hydra
    ckpt = create_checkpoint(model=hydra, path=None, directory=checkpoint_directory, max_to_keep=checkpoints_to_keep)
    trainer = hydra.build_trainer(loss_config=config["loss"], optimizer_config=dict(train_config["optimizer"],
                                                                                     gpu_batches_per_step=gpu_batches_per_step))
    loss_fn_map = loss_function_builder(config=config["loss"], trainer=trainer)   # sets the trainer's loss configuration
    del loss_fn_map
    trainer.bind_optimizer(optimizer)
    dataset.training.bind(trainer)
    if rank == 0:
        hydra.summary(print_fn=logger.info)
        hydra.save(os.path.join(str(checkpoint_directory), MODEL_HYDRA_DEFAULT_NAME_STR))

    def save_checkpoint_model_fn():
        if rank != 0:
            return
        logger.info("saving checkpoint at step: [{0}]".format(int(ckpt.step)))
        save_path = ckpt.save()
        logger.info(f"saved checkpoint to [{save_path}]")

    latest = ckpt.latest_checkpoint
    if latest:
        logger.info("!!! Found checkpoint to restore !!!")
        ckpt.restore(latest)
        logger.info(f"restored checkpoint at epoch [{int(ckpt.epoch)}] and step [{int(ckpt.step)}]")
        optimizer.iterations = ckpt.step      # restore learning rate (train_loop.py:180-181); Adam slots start afresh,
        trainer.bind_optimizer(optimizer)     # as in the reference (they are not part of its checkpoint)
    else:
        logger.info("!!! Did NOT find checkpoint to restore !!!")
        if weights_dir is not None and len(str(weights_dir)) > 0 and os.path.isdir(str(weights_dir)):
            try:
                logger.info(f"loading weights from [{weights_dir}]")
                hydra.set_weights(_load_weights_dir(str(weights_dir)))
                ckpt.step, ckpt.epoch = 0, 0
                logger.info("successfully loaded weights")
            except Exception as e:   # the reference logs and carries on with the initial weights (:205-211)
                logger.info(f"!!! failed to load weights from [{weights_dir}]] !!!")
                logger.error(f"!!! {e}")
        save_checkpoint_model_fn()

    model_no_outputs = len(hydra.outputs)
    deep_supervision_schedule = deep_supervision_schedule_builder(
        config=train_config.get("deep_supervision", {"type": "linear_low_to_high"}), no_outputs=model_no_outputs)

    metrics_path = os.path.join(str(checkpoint_directory), "metrics.jsonl")

    def log_metrics(extra: Dict):
        total_loss, model_loss, denoiser_loss = trainer.last_losses()     # the only device -> host read of the loop
        rec = {"step": int(ckpt.step), "epoch": int(ckpt.epoch), "loss/total": total_loss,
               "loss/regularization": model_loss[REGULARIZATION_LOSS_STR],
               "loss_denoiser/scale_0/mae": denoiser_loss[MAE_LOSS_STR],
               "loss_denoiser/scale_0/ssim": denoiser_loss[SSIM_LOSS_STR],
               "loss_denoiser/scale_0/total": denoiser_loss[TOTAL_LOSS_STR]}
        rec.update(extra)
        if rank == 0:
            with open(metrics_path, "a") as f:
                f.write(json.dumps(rec) + "\n")
        return rec

    # ---
    finished_training = False
    counter = 0
    start_time_forward_backward = time.time()
    last_record = None
    while not finished_training and (epochs == -1 or ckpt.epoch < epochs):
        logger.info("epoch [{0}], step [{1}]".format(int(ckpt.epoch), int(ckpt.step)))
        start_time_epoch = time.time()
        if epochs > 0:
            percentage_done = float(ckpt.epoch) / float(epochs)
        elif total_steps > 0:
            percentage_done = float(ckpt.step) / float(total_steps)
        else:
            percentage_done = 0.0
        depth_weight = deep_supervision_schedule(percentage_done=percentage_done)
        if abs(float(depth_weight[0]) - 1.0) > 1e-6:
            raise ValueError("deep supervision weights other than [1.0] need a multi-output backbone")
        if 0 < total_steps <= ckpt.step:
            logger.info("total_steps reached [{0}]".format(int(total_steps)))
            finished_training = True

        # --- iterate over the batches of the dataset (clean, noisy: float32 CUDA tensors; the corruption of
        # dataset.py:120-238 ran on the GPU inside the pipeline, per image like the reference's map stage)
        for input_image_batch, noisy_image_batch in dataset.training.epoch(int(ckpt.epoch)):
            if finished_training:
                break
            if counter == 0:
                start_time_forward_backward = time.time()
            _, _, _, grads = trainer.train_step_single_gpu(p_input_image_batch=input_image_batch,
                                                           p_noisy_image_batch=noisy_image_batch, sync=False)
            trainer.accumulate(grads)   # zeroes the accumulator at the first micro-batch of an update (:405-409)
            if counter >= (gpu_batches_per_step - 1 if exact_accumulation else gpu_batches_per_step):
                counter = 0
                # "average" (always 1/gpu_batches_per_step, whatever the number summed) and apply (:421-434)
                trainer.apply_grads(None, divisor=gpu_batches_per_step)
            else:
                counter += 1
                continue

            if visualization_every > 0 and (ckpt.step % visualization_every) == 0:
                dt = time.time() - start_time_forward_backward
                last_record = log_metrics({"training/learning_rate": float(lr_schedule(max(int(optimizer.iterations) - 1, 0))),
                                           "training/steps_per_second": 1.0 / (dt + 0.00001)})
            # --- check if it is time to save a checkpoint
            if checkpoint_every > 0 and ckpt.step > 0 and ckpt.step % checkpoint_every == 0:
                save_checkpoint_model_fn()
            ckpt.step += 1
            if 0 < total_steps <= ckpt.step:
                logger.info("total_steps reached [{0}]".format(int(total_steps)))
                finished_training = True
                break

        epoch_time = time.time() - start_time_epoch
        logger.info("end of epoch [{0}], took [{1}] seconds".format(int(ckpt.epoch), int(round(epoch_time))))
        ckpt.epoch += 1
        save_checkpoint_model_fn()

    if ckpt.step > 0:
        last_record = log_metrics({"training/learning_rate": float(lr_schedule(max(int(optimizer.iterations) - 1, 0))), "final": True})
    torch.cuda.synchronize(device)
    if on_finish is not None:
        on_finish(hydra)
    hydra.close()
    logger.info("finished training")
    return
