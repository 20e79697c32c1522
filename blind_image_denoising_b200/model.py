"""`model_builder(config) -> BuilderResults` with the reference's field names (reference bfcnn/model.py:25-34,58-162)
for the bias-free resnet family.

The reference returns five keras models.  Here the arithmetic lives in libbfcnn_b200.so, so the "models" are thin
host objects: they hold the variables in Keras order (what `hydra.variables` / `get_weights()` return in the reference,
SURVEY 8c), know which of them are trainable, and `hydra(...)` runs the B200 kernels.  Building them needs no GPU
(model_builder is pure graph construction in the reference too, model.py:58-162); the device handle is created on first
use and there is no CPU arithmetic behind any of these objects.
"""
from __future__ import annotations

import json
import os
from collections import namedtuple
from typing import Dict, List, Optional, Sequence

import numpy as np

from .arch import Arch, arch_from_config
from .weights import initial_variables

BACKBONE_STR = "backbone"      # reference bfcnn/constants.py
DENOISER_STR = "denoiser"
BATCH_SIZE_STR = "batch_size"

# reference bfcnn/model.py:25-34
BuilderResults = namedtuple("BuilderResults", ["backbone", "normalizer", "denormalizer", "denoiser", "hydra", "options"])


class ModelPart:
    """One of the sub-models of model.py:58-162 as a named view on the hydra variables."""

    def __init__(self, name: str, hydra: "HydraModel", indices: Sequence[int], note: str = ""):
        self.name, self._hydra, self._indices, self.note = name, hydra, list(indices), note

    @property
    def variables(self) -> List[np.ndarray]:
        v = self._hydra.variables
        return [v[i] for i in self._indices]

    @property
    def trainable_variables(self) -> List[np.ndarray]:
        v, m = self._hydra.variables, self._hydra.arch.trainable_mask()
        return [v[i] for i in self._indices if m[i]]

    def get_weights(self) -> List[np.ndarray]:
        return self.variables

    def count_params(self) -> int:
        return int(sum(v.size for v in self.variables))

    def __call__(self, *args, **kwargs):
        raise RuntimeError(f"[{self.name}] has no stand-alone kernel: {self.note or 'it is fused into the hydra kernels'}; "
                           "call the hydra model")


class HydraModel:
    """`models.hydra` (model.py:143-151): normaliser -> resnet backbone -> denoiser head -> denormaliser.

    variables / trainable_variables / get_weights / set_weights follow `keras.Model`; `__call__(x, training=False)` maps
    a float32 [N,H,W,3] batch (0..255; numpy or CUDA torch) to the float32 prediction (0..255) with the moving BN
    statistics.  Training goes through `trainer` (train_step_single_gpu), which owns the device copy of the variables."""

    name = "hydra"

    def __init__(self, config: Dict, variables: Optional[Sequence[np.ndarray]] = None, *, device: int = 0, seed: int = 0):
        self.config = config
        self.arch: Arch = arch_from_config({"model": config})
        self.device = int(device)
        shapes = self.arch.variable_shapes()
        if variables is None:
            variables = initial_variables(self.arch, seed)   # glorot_normal kernels, BN at its Keras initial state
        variables = [np.asarray(v, np.float32) for v in variables]
        if [tuple(v.shape) for v in variables] != [tuple(s) for s in shapes]:
            raise ValueError("variables do not match the architecture of the config")
        self._variables = variables
        self._trainer = None
        self.inputs = [("input_tensor", (None, None, None, self.arch.in_channels))]
        self.outputs = [("denoiser_head", (None, None, None, self.arch.out_channels))]   # one output: single-scale resnet

    # ---- keras.Model surface the train loop uses -------------------------------------
    @property
    def variables(self) -> List[np.ndarray]:
        if self._trainer is not None:
            self._variables = self._trainer.get_weights()   # the device copy is authoritative once training started
        return self._variables

    @property
    def trainable_variables(self) -> List[np.ndarray]:
        return [v for v, t in zip(self.variables, self.arch.trainable_mask()) if t]

    @property
    def non_trainable_variables(self) -> List[np.ndarray]:
        return [v for v, t in zip(self.variables, self.arch.trainable_mask()) if not t]

    def get_weights(self) -> List[np.ndarray]:
        return [v.copy() for v in self.variables]

    def set_weights(self, weights: Sequence[np.ndarray]):
        weights = [np.asarray(w, np.float32) for w in weights]
        if [w.shape for w in weights] != [v.shape for v in self._variables]:
            raise ValueError("set_weights: shapes do not match hydra.variables")
        self._variables = weights
        if self._trainer is not None:
            self._trainer.set_weights(weights)

    def count_params(self) -> int:
        return self.arch.num_weights()

    def summary(self, print_fn=print):
        a = self.arch
        print_fn(f"Model: \"{self.name}\" (bias-free resnet denoiser, B200 kernels)")
        print_fn(f"  normalizer    clip[0,255]/255-0.5 (fused into the base conv)")
        print_fn(f"  base conv     {a.base_kernel}x{a.base_kernel} {a.in_channels}->{a.filters}, linear, no BN")
        print_fn(f"  {a.no_layers} x block     conv3x3 -> ReLU -> conv3x3 -> BN(no beta) -> add")
        print_fn(f"  denoiser head 1x1 {a.filters}->{a.head_filters} -> 1x1 {a.head_filters}->{a.out_channels} -> tanh(2y)*0.51")
        print_fn(f"  denormalizer  (clip(+-0.5)+0.5)*255")
        print_fn(f"Total params: {a.num_weights()}  Trainable: {a.num_trainable()}  Non-trainable: {a.num_weights() - a.num_trainable()}")

    def save(self, path: str):
        """`ckpt.model.save(<dir>/model_hydra.keras)` (train_loop.py:155-156): here a directory with the variables as a
        TensorBundle (readable by `bfcnn.load_model` and by tf.train.load_checkpoint) and the model config."""
        from .tensorbundle import write_model_variables
        os.makedirs(os.path.join(path, "variables"), exist_ok=True)
        write_model_variables(os.path.join(path, "variables"), self.variables)
        with open(os.path.join(path, "pipeline.json"), "w") as f:
            json.dump({"model": self.config}, f, indent=4)

    # ---- device side -------------------------------------------------------------------
    @property
    def trainer(self):
        if self._trainer is None:
            raise RuntimeError("hydra has no trainer yet: call hydra.build_trainer(...) (train_loop does)")
        return self._trainer

    def build_trainer(self, *, loss_config=None, optimizer_config=None, process_group=None, device: Optional[int] = None):
        from .training import Trainer
        if device is not None:
            self.device = int(device)
        if self._trainer is not None:
            self._trainer.close()
        self._trainer = Trainer(self.arch, self._variables, device=self.device, loss_config=loss_config,
                                optimizer_config=optimizer_config, process_group=process_group)
        return self._trainer

    def __call__(self, inputs, training: bool = False):
        if training:
            raise RuntimeError("hydra(x, training=True) is fused with the loss and the backward pass: "
                               "use hydra.trainer.train_step_single_gpu(clean, noisy)")
        import torch
        if isinstance(inputs, (list, tuple)):   # the reference calls ckpt.model([batch]) (train_loop.py:248-256)
            inputs = inputs[0]
        if self._trainer is None:
            self.build_trainer()
        as_numpy = not torch.is_tensor(inputs)
        x = torch.as_tensor(np.asarray(inputs, np.float32)) if as_numpy else inputs
        if not x.is_cuda:
            x = x.cuda(self.device)
        y = self._trainer.predict(x.float())
        return y.cpu().numpy() if as_numpy else y

    def close(self):
        if self._trainer is not None:
            self._variables = self._trainer.get_weights()
            self._trainer.close()
            self._trainer = None


def model_builder(config: Dict, variables: Optional[Sequence[np.ndarray]] = None, *, device: int = 0,
                  seed: int = 0) -> BuilderResults:
    """model.py:58-162 for `type: "resnet"` backbones of the 16-channel, two-3x3-convs-per-block family.
    `config` is the "model" section of a pipeline config: {"backbone": {...}, "denoiser": {...}[, "batch_size": b]}."""
    if DENOISER_STR not in config or BACKBONE_STR not in config:
        raise KeyError(f"model config needs [{BACKBONE_STR}] and [{DENOISER_STR}] sections")
    hydra = HydraModel(config, variables, device=device, seed=seed)
    n = hydra.arch.num_variables()
    backbone = ModelPart("resnet", hydra, range(0, n - 2), "base conv + residual blocks run inside the fused conv stack")
    denoiser = ModelPart("denoiser_head", hydra, [n - 2, n - 1], "the 1x1 head runs in the epilogue of the last pass")
    normalizer = ModelPart("normalize", hydra, [], "clip[0,255]/255-0.5 is folded into the base conv (utilities.py:449-461)")
    denormalizer = ModelPart("denormalize", hydra, [], "(clip(+-0.5)+0.5)*255 is part of the head epilogue (utilities.py:435-443)")
    return BuilderResults(backbone=backbone, normalizer=normalizer, denormalizer=denormalizer, denoiser=denoiser,
                          hydra=hydra, options={})
