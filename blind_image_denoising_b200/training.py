"""Host-side mirror of the reference's training-step interface (the arithmetic is in
libbfcnn_b200.so; this file moves pointers and reads configs).

reference callable                                   mirror here
---------------------------------------------------  -------------------------------------------
dataset_builder(config).prepare_data_fn              dataset_builder(config) -> DatasetResults with
  (bfcnn/dataset.py:40-238)                            .prepare_data_fn(clean_u8, seed, sample_offset)
loss_function_builder(config)["denoiser"|"model"]    loss_function_builder(config, trainer)
  (bfcnn/loss.py:152-253)
optimizer_builder / schedule_builder                 schedule_builder(config) (exponential / cosine / restarts),
  (bfcnn/optimizer.py:83-224)                          Adam hyper-parameters parsed by Trainer
train_step_single_gpu + apply_grads                  Trainer.train_step_single_gpu / Trainer.apply_grads
  (bfcnn/train_loop.py:263-321,404-434)                (+ gradient all-reduce under torch.distributed)

torch is used for device memory, streams and torch.distributed only.  File IO, tf.data,
TensorBoard and checkpoints are out of scope; inputs are device uint8 tensors.
"""
from __future__ import annotations

import ctypes
import math
from collections import namedtuple
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _native
from .arch import Arch, arch_from_config
from .weights import flatten_variables, unflatten_variables

# reference bfcnn/constants.py:37-50,75-76
TOTAL_LOSS_STR = "total_loss"
MAE_LOSS_STR = "mae_loss"
MSE_LOSS_STR = "mse_loss"
SSIM_LOSS_STR = "ssim_loss"
REGULARIZATION_LOSS_STR = "regularization_loss"
MODEL_LOSS_FN_STR = "model"
DENOISER_LOSS_FN_STR = "denoiser"

# reference bfcnn/dataset.py:27-35 plus the callable our callers need
DatasetResults = namedtuple("DatasetResults",
                            ["config", "batch_size", "input_shape", "training", "testing", "prepare_data_fn",
                             "noise_config"])


def _torch():
    import torch
    return torch


def _stream_ptr(device_index: int):
    torch = _torch()
    return ctypes.c_void_p(torch.cuda.current_stream(device_index).cuda_stream)


# --------------------------------------------------------------------------------------
# configs
# --------------------------------------------------------------------------------------
def noise_cfg_from_config(config: Dict) -> _native.NoiseCfg:
    """Parse the `dataset` section exactly as bfcnn/dataset.py:74-105 does (keys that HEAD parses but
    never applies -- blur, rotate, quantization, jpeg, inpaint -- must be off)."""
    for key, off in (("random_blur", False), ("use_jpeg_noise", False)):
        if config.get(key, off):
            raise ValueError(f"dataset option {key} is parsed but not applied by the reference (dataset.py:74-105)")
    additional = list(config.get("additional_noise", []))
    multiplicative = list(config.get("multiplicative_noise", []))
    a_min, a_max = (min(additional), max(additional)) if additional else (0.0, 0.0)
    m_min, m_max = (min(multiplicative), max(multiplicative)) if multiplicative else (0.0, 0.0)
    return _native.NoiseCfg(float(a_min), float(a_max), float(m_min), float(m_max),
                            int(bool(config.get("random_left_right", False))),
                            int(bool(config.get("random_up_down", False))),
                            int(bool(config.get("subsample", False))),
                            int(bool(config.get("round_values", True))))


def loss_cfg_from_config(config: Dict) -> _native.LossCfg:
    """bfcnn/loss.py:164-181, same keys and defaults (note: the reference's default ssim_multiplier is 1.0)."""
    return _native.LossCfg(float(config.get("hinge", 0.0)), float(config.get("cutoff", 255.0)),
                           float(config.get("mae_multiplier", 1.0)), float(config.get("mse_multiplier", 0.0)),
                           float(config.get("regularization", 1.0)), float(config.get("ssim_multiplier", 1.0)))


def schedule_builder(config: Dict) -> Callable[[int], float]:
    """bfcnn/optimizer.py:83-139: step -> learning rate (keras ExponentialDecay / CosineDecay /
    CosineDecayRestarts closed forms)."""
    if config is None:
        raise ValueError("schedule_type cannot be None")
    schedule_type = config.get("type", None)
    if schedule_type is None:
        raise ValueError("schedule_type cannot be None")
    if not isinstance(schedule_type, str):
        raise ValueError("schedule_type must be a string")
    params = config.get("config", {})
    schedule_type = schedule_type.strip().lower()
    lr0 = float(params["learning_rate"])
    if schedule_type == "exponential_decay":
        rate, steps = float(params["decay_rate"]), float(params["decay_steps"])
        return lambda step: lr0 * rate ** (step / steps)
    if schedule_type == "cosine_decay":
        steps, alpha = float(params["decay_steps"]), float(params.get("alpha", 0.0001))

        def cosine(step):
            p = min(step, steps) / steps
            return lr0 * ((1 - alpha) * 0.5 * (1 + math.cos(math.pi * p)) + alpha)
        return cosine
    if schedule_type == "cosine_decay_restarts":
        first, t_mul = float(params["decay_steps"]), float(params.get("t_mul", 2.0))
        m_mul, alpha = float(params.get("m_mul", 0.9)), float(params.get("alpha", 0.001))

        def restarts(step):
            c = step / first
            if t_mul == 1.0:
                i = math.floor(c)
                frac = c - i
            else:
                i = math.floor(math.log(1 - c * (1 - t_mul)) / math.log(t_mul))
                frac = (c - (1 - t_mul ** i) / (1 - t_mul)) / t_mul ** i
            return lr0 * ((1 - alpha) * (m_mul ** i) * 0.5 * (1 + math.cos(math.pi * frac)) + alpha)
        return restarts
    raise ValueError(f"don't know how to handle learning_rate schedule_type [{schedule_type}]")


# --------------------------------------------------------------------------------------
# the trainer: one model replica on one GPU
# --------------------------------------------------------------------------------------
class Trainer:
    """A trainable replica of the resnet denoiser on one B200 (one handle per device).

    Under torch.distributed (one process per GPU) `apply_grads` all-reduces the flat gradient
    vector (<= 84 272 floats) over NCCL before the fused Adam kernel: data-parallel training
    with local BN batch statistics, the multi-GPU mapping of the reference's micro-batch
    accumulation loop (train_loop.py:404-434)."""

    def __init__(self, arch: Arch, variables: Sequence[np.ndarray], *, device: int = 0,
                 loss_config: Optional[Dict] = None, optimizer_config: Optional[Dict] = None,
                 process_group=None, conv_engine: str = "t5"):
        torch = _torch()
        self._lib = _native.load_library()
        self.arch = arch
        self.device = int(device)
        self.process_group = process_group
        flat = np.ascontiguousarray(flatten_variables(arch, variables), dtype=np.float32)
        carch = arch.to_c()
        h = ctypes.c_void_p()
        _native.check(self._lib.bfcnn_create(ctypes.byref(carch), flat.ctypes.data, flat.size, self.device,
                                             ctypes.byref(h)))
        self._h = h
        engines = {"fp32": 0, "x3": 1, "t5": 2}
        if conv_engine not in engines:
            raise ValueError("conv_engine must be 't5' (tcgen05, fp16 hi/lo split), 'x3' (mma.sync, fp16 hi/lo split) or 'fp32' (FFMA)")
        _native.check(self._lib.bfcnn_set_train_engine(self._h, engines[conv_engine]))
        self.conv_engine = conv_engine
        self.loss_cfg = loss_cfg_from_config(loss_config if loss_config is not None else
                                             {"hinge": 0.5, "cutoff": 255.0, "mae_multiplier": 1.0,
                                              "ssim_multiplier": 0.0, "mse_multiplier": 0.0, "regularization": 0.01})
        oc = dict(optimizer_config or {})
        sched = oc.get("schedule", {"type": "exponential_decay",
                                    "config": {"learning_rate": 1e-3, "decay_rate": 1.0, "decay_steps": 1}})
        self.schedule = schedule_builder(sched)
        self.beta_1 = float(oc.get("beta_1", 0.9))
        self.beta_2 = float(oc.get("beta_2", 0.999))
        self.epsilon = float(oc.get("epsilon", 1e-7))
        clip = oc.get("gradient_clipping_by_norm", None)   # optimizer.py:169 -> global_clipnorm
        self.global_clipnorm = float(clip) if clip else 0.0
        self.gpu_batches_per_step = int(oc.get("gpu_batches_per_step", 1))
        if self.gpu_batches_per_step <= 0:
            raise ValueError("gpu_batches_per_step must be > 0")   # train_loop.py:113-115
        self.step = 0
        self.flat_grads = torch.zeros(arch.num_trainable(), dtype=torch.float32, device=f"cuda:{self.device}")
        self._accum = None
        self._accum_count = 0

    # ------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.bfcnn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def get_weights(self) -> List[np.ndarray]:
        flat = np.empty(self.arch.num_weights(), dtype=np.float32)
        _native.check(self._lib.bfcnn_get_weights(self._h, flat.ctypes.data, flat.size))
        return unflatten_variables(self.arch, flat)

    def set_weights(self, variables: Sequence[np.ndarray]):
        flat = np.ascontiguousarray(flatten_variables(self.arch, variables), dtype=np.float32)
        _native.check(self._lib.bfcnn_set_weights(self._h, flat.ctypes.data, flat.size))

    def launch_count(self) -> int:
        return int(self._lib.bfcnn_launch_count(self._h))

    def _check_dev(self, t, dtype, name):
        torch = _torch()
        if not (torch.is_tensor(t) and t.is_cuda and t.device.index == self.device):
            raise ValueError(f"{name} must be a CUDA tensor on cuda:{self.device}")
        if t.dtype != dtype:
            raise TypeError(f"{name} must be {dtype}")
        if t.dim() != 4 or t.shape[-1] != 3:
            raise ValueError(f"{name} must have shape [N,H,W,3]")
        return t.contiguous()

    # ------------------------------------------------------------------ dataset.py:120-238
    def prepare_data(self, clean_u8, noise_cfg: _native.NoiseCfg, seed: int, sample_offset: int = 0):
        torch = _torch()
        x = self._check_dev(clean_u8, torch.uint8, "clean_u8")
        n, h, w, _ = x.shape
        clean = torch.empty((n, h, w, 3), dtype=torch.float32, device=x.device)
        noisy = torch.empty_like(clean)
        _native.check(self._lib.bfcnn_corrupt(self._h, x.data_ptr(), clean.data_ptr(), noisy.data_ptr(), n, h, w,
                                              ctypes.c_uint64(seed), ctypes.c_uint64(sample_offset),
                                              ctypes.byref(noise_cfg), _stream_ptr(self.device)))
        return clean, noisy

    # ------------------------------------------------------------------ loss.py:190-247
    def denoiser_loss(self, gt_batch, predicted_batch) -> Dict[str, float]:
        torch = _torch()
        gt = self._check_dev(gt_batch, torch.float32, "gt_batch")
        pr = self._check_dev(predicted_batch, torch.float32, "predicted_batch")
        if gt.shape != pr.shape:
            raise ValueError("gt_batch and predicted_batch differ in shape")
        n, h, w, _ = gt.shape
        out = (ctypes.c_float * 5)()
        _native.check(self._lib.bfcnn_loss(self._h, gt.data_ptr(), pr.data_ptr(), n, h, w, ctypes.byref(self.loss_cfg),
                                           out, _stream_ptr(self.device)))
        return {TOTAL_LOSS_STR: out[0], MAE_LOSS_STR: out[1], MSE_LOSS_STR: out[2], SSIM_LOSS_STR: out[4],
                "hinged_mae": out[3]}

    # ------------------------------------------------------------------ train_loop.py:263-312
    def train_step_single_gpu(self, p_input_image_batch, p_noisy_image_batch, update_moving: bool = True):
        """Forward (BN batch statistics) + loss + backward.  Returns (total_loss, model_loss dict,
        denoiser loss dict, flat gradient tensor [num_trainable] in Keras trainable_variables order)."""
        torch = _torch()
        clean = self._check_dev(p_input_image_batch, torch.float32, "p_input_image_batch")
        noisy = self._check_dev(p_noisy_image_batch, torch.float32, "p_noisy_image_batch")
        if clean.shape != noisy.shape:
            raise ValueError("clean and noisy batches differ in shape")
        n, h, w, _ = clean.shape
        out = (ctypes.c_float * 5)()
        _native.check(self._lib.bfcnn_train_step(self._h, clean.data_ptr(), noisy.data_ptr(), n, h, w,
                                                 ctypes.byref(self.loss_cfg), self.flat_grads.data_ptr(), out,
                                                 int(bool(update_moving)), _stream_ptr(self.device)))
        model_loss = {REGULARIZATION_LOSS_STR: out[3], TOTAL_LOSS_STR: out[3] * self.loss_cfg.regularization}
        denoiser = {TOTAL_LOSS_STR: out[1], MAE_LOSS_STR: out[2], SSIM_LOSS_STR: out[4]}
        return out[0], model_loss, denoiser, self.flat_grads

    # ------------------------------------------------------------------ train_loop.py:314-321,418-434
    def accumulate(self, grads) -> bool:
        """`gradients_accumulation[i].assign_add(grad)`; True when `gpu_batches_per_step` micro-batches
        are in (the reference's counter quirk -- k+1 batches divided by k -- is NOT reproduced)."""
        if self._accum is None:
            self._accum = _torch().zeros_like(grads)
        if self._accum_count == 0:
            self._accum.zero_()
        self._accum.add_(grads)
        self._accum_count += 1
        return self._accum_count >= self.gpu_batches_per_step

    def apply_grads(self, grads=None):
        """All-reduce (sum over ranks) the flat gradient, then one fused Adam(+global clipnorm) step with the
        averaging factor 1/(world * accumulated micro-batches)."""
        torch = _torch()
        import torch.distributed as dist
        if grads is None:
            grads, k = self._accum, max(self._accum_count, 1)
            self._accum_count = 0
        else:
            k = 1
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.process_group)
        self.step += 1
        lr = float(self.schedule(self.step - 1))  # keras evaluates the schedule at optimizer.iterations (pre-increment)
        cfg = _native.AdamCfg(lr, self.beta_1, self.beta_2, self.epsilon, self.global_clipnorm)
        _native.check(self._lib.bfcnn_adam_step(self._h, grads.data_ptr(), ctypes.c_float(1.0 / (world * k)),
                                                ctypes.byref(cfg), ctypes.c_int64(self.step), _stream_ptr(self.device)))
        return lr


# --------------------------------------------------------------------------------------
# builders with the reference's names
# --------------------------------------------------------------------------------------
def dataset_builder(config: Dict, trainer: Optional[Trainer] = None) -> DatasetResults:
    """bfcnn/dataset.py:40-305 restricted to the corruption function: `training`/`testing` (the tf.data
    file pipelines) are None; `prepare_data_fn(clean_u8, seed, sample_offset)` needs a trainer (the GPU)."""
    ncfg = noise_cfg_from_config(config)

    def prepare_data_fn(input_batch, seed: int = 0, sample_offset: int = 0, trainer_: Optional[Trainer] = None):
        t = trainer_ or trainer
        if t is None:
            raise RuntimeError("prepare_data_fn needs a Trainer: the corruption runs on the GPU, there is no CPU path")
        return t.prepare_data(input_batch, ncfg, seed, sample_offset)

    return DatasetResults(config=config, batch_size=config.get("batch_size", 32),
                          input_shape=config.get("input_shape", [256, 256, 3]), training=None, testing=None,
                          prepare_data_fn=prepare_data_fn, noise_config=ncfg)


def loss_function_builder(config: Dict, trainer: Trainer) -> Dict[str, Callable]:
    """bfcnn/loss.py:152-253: {"model": fn, "denoiser": fn(gt_batch, predicted_batch)}."""
    trainer.loss_cfg = loss_cfg_from_config(config)

    def denoiser_loss(gt_batch, predicted_batch):
        return trainer.denoiser_loss(gt_batch, predicted_batch)

    def model_loss(model=None):
        raise RuntimeError("the regularisation term is computed inside train_step_single_gpu (fused with the "
                           "weight-gradient epilogue); read it from the step's model_loss result")

    return {MODEL_LOSS_FN_STR: model_loss, DENOISER_LOSS_FN_STR: denoiser_loss}


def trainer_from_config(pipeline_config: Dict, variables: Optional[Sequence[np.ndarray]] = None, *, device: int = 0,
                        seed: int = 0) -> Trainer:
    """model_builder + loss_function_builder + optimizer_builder of train_loop.py:80-148 for this family."""
    from .weights import synthetic_variables
    arch = arch_from_config(pipeline_config)
    if variables is None:
        variables = synthetic_variables(arch, seed)
    train_cfg = dict(pipeline_config.get("train", {}))
    opt = dict(train_cfg.get("optimizer", {}))
    if "gpu_batches_per_step" in train_cfg:
        opt["gpu_batches_per_step"] = train_cfg["gpu_batches_per_step"]
    return Trainer(arch, variables, device=device, loss_config=pipeline_config.get("loss"), optimizer_config=opt)
