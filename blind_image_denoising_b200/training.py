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
import os
from collections import namedtuple
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import _native
from .arch import Arch, arch_from_config
from .optimizer import (AdamOptimizer, deep_supervision_schedule_builder, optimizer_builder, optimizer_config_check,
                        schedule_builder)
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
    # round_values is parsed by the reference (dataset.py:84) but tf.round is applied unconditionally (dataset.py:228):
    # the kernel's round flag is therefore always on here.  draw_group = no_crops_per_image: the crops of one image share
    # one call of prepare_data_fn and with it the flips / noise switches / sigmas (dataset.py:276-297).
    return _native.NoiseCfg(float(a_min), float(a_max), float(m_min), float(m_max),
                            int(bool(config.get("random_left_right", False))),
                            int(bool(config.get("random_up_down", False))),
                            int(bool(config.get("subsample", False))),
                            1, int(config.get("no_crops_per_image", 1)))


def loss_cfg_from_config(config: Dict) -> _native.LossCfg:
    """bfcnn/loss.py:164-181, same keys and defaults (note: the reference's default ssim_multiplier is 1.0)."""
    return _native.LossCfg(float(config.get("hinge", 0.0)), float(config.get("cutoff", 255.0)),
                           float(config.get("mae_multiplier", 1.0)), float(config.get("mse_multiplier", 0.0)),
                           float(config.get("regularization", 1.0)), float(config.get("ssim_multiplier", 1.0)))


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
                 process_group=None, conv_engine: str = "t5", nccl_comm=None):
        torch = _torch()
        self._lib = _native.load_library()
        self.arch = arch
        self.device = int(device)
        self.process_group = process_group
        # a distributed.NcclCommunicator: the gradient exchange then goes through the C ABI (bfcnn_allreduce_grads, a raw
        # ncclAllReduce on the compute stream) instead of torch.distributed
        self.nccl_comm = nccl_comm
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
        # `Trainer` is this package's own entry point: a missing "type" means Adam here (optimizer_builder and
        # train_loop keep the reference's default, RMSprop, and therefore refuse a config without a type); everything the
        # fused Adam kernel does not implement is refused either way (optimizer.py:145-224)
        optimizer_config_check(dict({"type": "Adam"}, **oc))
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
        self._optimizer = None
        self._last_shape = (0, 0, 0)

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
    def train_step_single_gpu(self, p_input_image_batch, p_noisy_image_batch, update_moving: bool = True, sync: bool = True):
        """Forward (BN batch statistics) + loss + backward.  Returns (total_loss, model_loss dict,
        denoiser loss dict, flat gradient tensor [num_trainable] in Keras trainable_variables order).

        sync=False keeps the step asynchronous: nothing is copied back, the first three results are None and
        `last_losses()` fetches them later (the reference reads its loss tensors only when it logs them,
        train_loop.py:439-559), so corrupt -> step -> all-reduce -> Adam chain on the stream without a host round trip."""
        torch = _torch()
        clean = self._check_dev(p_input_image_batch, torch.float32, "p_input_image_batch")
        noisy = self._check_dev(p_noisy_image_batch, torch.float32, "p_noisy_image_batch")
        if clean.shape != noisy.shape:
            raise ValueError("clean and noisy batches differ in shape")
        n, h, w, _ = clean.shape
        self._last_shape = (n, h, w)
        out = (ctypes.c_float * 5)() if sync else None
        _native.check(self._lib.bfcnn_train_step(self._h, clean.data_ptr(), noisy.data_ptr(), n, h, w,
                                                 ctypes.byref(self.loss_cfg), self.flat_grads.data_ptr(), out,
                                                 int(bool(update_moving)), _stream_ptr(self.device)))
        if not sync:
            return None, None, None, self.flat_grads
        return (out[0],) + self._loss_dicts(out) + (self.flat_grads,)

    def _loss_dicts(self, out):
        model_loss = {REGULARIZATION_LOSS_STR: out[3], TOTAL_LOSS_STR: out[3] * self.loss_cfg.regularization}
        denoiser = {TOTAL_LOSS_STR: out[1], MAE_LOSS_STR: out[2], SSIM_LOSS_STR: out[4]}
        return model_loss, denoiser

    def last_losses(self):
        """(total_loss, model_loss dict, denoiser loss dict) of the LAST train step; synchronises the stream."""
        out = (ctypes.c_float * 5)()
        _native.check(self._lib.bfcnn_train_losses(self._h, out, _stream_ptr(self.device)))
        return (out[0],) + self._loss_dicts(out)

    def saved_activation(self, which: str, index: int):
        """Test hook: activation map saved by the last train step, float32 [n,h,w,16].  which: "x" (input of block
        `index`; index N = output of the stack), "t" (ReLU(conv_a)), "u" (conv_b output before BatchNormalization)."""
        torch = _torch()
        n, h, w = self._last_shape
        out = torch.empty((n, h, w, 16), dtype=torch.float32, device=f"cuda:{self.device}")
        _native.check(self._lib.bfcnn_saved_activation(self._h, {"x": 0, "t": 1, "u": 2}[which], int(index), out.data_ptr(),
                                                       _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ utilities.py:625-685, train_loop.py:239-247
    def multiscales(self, input_batch, no_scales: int, clip_values: bool = True, round_values: bool = True):
        """`multiscales_generator_fn(...)(n)`: [n, avg_pool2(n), avg_pool2(avg_pool2(n)), ...] with clip and round per
        level (no_scales pooled levels after the input itself)."""
        torch = _torch()
        x = self._check_dev(input_batch, torch.float32, "input_batch")
        scales = [x]
        for _ in range(int(no_scales)):
            n, h, w, _c = x.shape
            y = torch.empty((n, h // 2, w // 2, 3), dtype=torch.float32, device=x.device)
            _native.check(self._lib.bfcnn_downscale2x(self._h, x.data_ptr(), y.data_ptr(), n, h, w, int(bool(clip_values)),
                                                      int(bool(round_values)), _stream_ptr(self.device)))
            scales.append(y)
            x = y
        return scales

    # ------------------------------------------------------------------ model.py:100-116 (hydra, inference mode)
    def predict(self, image_batch, out=None):
        """`hydra(x, training=False)`: float32 [N,H,W,3] (0..255) -> float32 prediction (0..255), moving BN statistics,
        no pow2 canvas (that belongs to DenoiserModule), on the reference-grade FP32 path."""
        torch = _torch()
        x = self._check_dev(image_batch, torch.float32, "image_batch")
        n, h, w, _ = x.shape
        if out is None:
            out = torch.empty_like(x)
        flags = _native.FLAG_IN_DEVICE | _native.FLAG_OUT_DEVICE | _native.FLAG_NO_PAD_POW2 | _native.FLAG_IN_F32
        _native.check(self._lib.bfcnn_denoise_f32(self._h, x.data_ptr(), out.data_ptr(), n, h, w, _native.PREC_FP32, flags,
                                                  _stream_ptr(self.device)))
        return out

    # ------------------------------------------------------------------ train_loop.py:314-321,418-434
    def accumulate(self, grads) -> bool:
        """`gradients_accumulation[i].assign_add(grad)`; True once `gpu_batches_per_step` micro-batches are in.
        (`train_loop` reproduces the reference's counter quirk -- k+1 micro-batches divided by k -- on top of this.)"""
        if self._accum is None:
            self._accum = _torch().zeros_like(grads)
        if self._accum_count == 0:
            self._accum.zero_()
        self._accum.add_(grads)
        self._accum_count += 1
        return self._accum_count >= self.gpu_batches_per_step

    def bind_optimizer(self, optimizer: AdamOptimizer):
        """Use the hyper-parameters and the iteration counter of an `optimizer_builder` optimizer."""
        self.schedule = optimizer.learning_rate_schedule
        self.beta_1, self.beta_2, self.epsilon = optimizer.beta_1, optimizer.beta_2, optimizer.epsilon
        self.global_clipnorm = optimizer.global_clipnorm
        self.step = int(optimizer.iterations)
        self._optimizer = optimizer

    def apply_grads(self, grads=None, divisor: Optional[float] = None):
        """All-reduce (sum over ranks) the flat gradient, then one fused Adam(+global clipnorm) step with the
        averaging factor 1/(world * divisor); divisor defaults to the number of accumulated micro-batches."""
        torch = _torch()
        import torch.distributed as dist
        if grads is None:
            if self._accum is None or self._accum_count == 0:
                raise ValueError("apply_grads() without gradients: call accumulate(grads) first or pass grads")
            grads, k = self._accum, self._accum_count
            self._accum_count = 0
        else:
            k = 1
        if divisor is not None:
            k = float(divisor)
        world = 1
        if self.nccl_comm is not None:
            world = self.nccl_comm.world
            _native.check(self._lib.bfcnn_allreduce_grads(self._h, grads.data_ptr(), self.nccl_comm.comm, _stream_ptr(self.device)))
        elif dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(self.process_group)
            if world > 1:
                dist.all_reduce(grads, op=dist.ReduceOp.SUM, group=self.process_group)
        self.step += 1
        lr = float(self.schedule(self.step - 1))  # keras evaluates the schedule at optimizer.iterations (pre-increment)
        cfg = _native.AdamCfg(lr, self.beta_1, self.beta_2, self.epsilon, self.global_clipnorm)
        _native.check(self._lib.bfcnn_adam_step(self._h, grads.data_ptr(), ctypes.c_float(1.0 / (world * k)),
                                                ctypes.byref(cfg), ctypes.c_int64(self.step), _stream_ptr(self.device)))
        if getattr(self, "_optimizer", None) is not None:
            self._optimizer.iterations = self.step
        return lr


# --------------------------------------------------------------------------------------
# builders with the reference's names
# --------------------------------------------------------------------------------------
IMAGE_EXTENSIONS = (".png", ".jpg", ".jpeg", ".bmp", ".gif", ".webp")   # file_operations.py:37-96 (image_filenames_generator)


def load_image(path, image_size=None, num_channels: int = 3, expand_dims: bool = True, normalize: bool = False,
               interpolation: str = "bilinear"):
    """file_operations.py:101-159 as far as the training input needs it: decode an image file to a uint8 array
    ([1,H,W,C] with expand_dims), optionally resized to image_size=(h, w).  Decoding is host-side IO (Pillow)."""
    from PIL import Image
    img = Image.open(str(path))
    img = img.convert({1: "L", 3: "RGB", 4: "RGBA"}[int(num_channels)])
    if image_size is not None:
        img = img.resize((int(image_size[1]), int(image_size[0])), Image.BILINEAR if interpolation == "bilinear" else Image.NEAREST)
    a = np.asarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[..., None]
    if normalize:
        a = a.astype(np.float32) / 255.0
    return a[None] if expand_dims else a


class ImagePipeline:
    """The tf.data pipeline of dataset.py:239-297 with the GPU doing the corruption:

        file names -> shuffle(seed 0, reshuffle every epoch) -> load + no_crops_per_image random crops (host, uint8)
        -> prepare_data_fn on the crops of each image (GPU: ONE launch per pool of images; draw_group =
           no_crops_per_image makes the crops of an image share the flips / switches / sigmas, dataset.py:141-187)
        -> unbatch -> shuffle (device permutation of the pool) -> batch(batch_size, drop_remainder=True)

    Sources: image files under `inputs[*].directory`, in-memory uint8 arrays (`images=`), or an `inputs` entry
    {"synthetic": {"images": M, "height": h, "width": w, "seed": s}} (uniform random uint8 images, for benchmarks and
    tests).  Under data-parallel training rank r of `world` takes every world-th image of the epoch's order
    and its Philox sample indices live in their own 2^40-wide range."""

    def __init__(self, config: Dict, noise_cfg: _native.NoiseCfg, images=None, rank: int = 0, world: int = 1):
        self.batch_size = int(config["batch_size"])
        self.input_shape = list(config["input_shape"])
        self.crops_per_image = int(config.get("no_crops_per_image", 1))
        self.seed = int(config.get("seed", 0))
        self.noise_cfg = noise_cfg
        self.rank, self.world = int(rank), int(world)
        self.trainer = None
        self._epoch = 0
        pool = config.get("shuffle_buffer_images", None)
        self.pool_images = int(pool) if pool else max(1, (self.batch_size * 128) // self.crops_per_image)
        self.sources: List = []
        if images is not None:
            self.sources = [np.ascontiguousarray(np.asarray(im, np.uint8)) for im in images]
        else:
            inputs = config.get("inputs", [])
            if isinstance(inputs, dict):
                inputs = [inputs]
            if not isinstance(inputs, list):
                raise ValueError("dont know how to handle anything else than list and dict")
            for entry in inputs:
                if "synthetic" in entry:
                    sp = entry["synthetic"]
                    rng = np.random.default_rng(int(sp.get("seed", 0)))
                    h, w = int(sp.get("height", self.input_shape[0])), int(sp.get("width", self.input_shape[1]))
                    self.sources += [rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8) for _ in range(int(sp["images"]))]
                d = entry.get("directory", None)
                if d:
                    for root, _, files in sorted(os.walk(str(d))):
                        self.sources += [os.path.join(root, f) for f in sorted(files) if f.lower().endswith(IMAGE_EXTENSIONS)]
        if not self.sources:
            raise ValueError("don't know how to handle non directory datasets")   # dataset.py:253

    def bind(self, trainer: "Trainer"):
        self.trainer = trainer
        return self

    def __len__(self):
        per_rank = len(self.sources[self.rank::self.world])
        return (per_rank * self.crops_per_image) // self.batch_size

    def _source(self, i: int, dev: str):
        """Source i as the pipeline crops it: a file name (decoded on the host at every visit, as the reference's map stage
        does), or -- in-memory images -- a uint8 CUDA tensor uploaded once, so that an epoch moves no pixels over PCIe."""
        src = self.sources[i]
        if isinstance(src, str):
            return src
        cache = self.__dict__.setdefault("_dev_sources", {})
        if i not in cache:
            cache[i] = _torch().from_numpy(src).to(dev)
        return cache[i]

    def _crops(self, src, rng: np.random.Generator):
        img = load_image(src, num_channels=3, expand_dims=False) if isinstance(src, str) else src
        ch, cw = int(self.input_shape[0]), int(self.input_shape[1])
        if img.shape[0] < ch or img.shape[1] < cw:
            return []
        out = []
        for _ in range(self.crops_per_image):
            y0 = int(rng.integers(0, img.shape[0] - ch + 1)); x0 = int(rng.integers(0, img.shape[1] - cw + 1))
            out.append(img[y0:y0 + ch, x0:x0 + cw, :3])
        return out

    def __iter__(self):
        e = self._epoch
        self._epoch += 1
        return self.epoch(e)

    def epoch(self, e: int):
        """Yield (clean, noisy) float32 CUDA batches [batch_size, h, w, 3] of epoch e."""
        torch = _torch()
        if self.trainer is None:
            raise RuntimeError("ImagePipeline needs a Trainer (bind): the corruption runs on the GPU, there is no CPU path")
        dev = f"cuda:{self.trainer.device}"
        rng = np.random.default_rng([self.seed, int(e)])
        order = rng.permutation(len(self.sources))[self.rank::self.world]
        gen = torch.Generator(device=dev)
        gen.manual_seed(self.seed * 1000003 + int(e) * 8191 + self.rank)
        c = self.crops_per_image
        shard_len = (len(self.sources) + self.world - 1) // self.world
        carry_clean = carry_noisy = None
        for p0 in range(0, len(order), self.pool_images):
            crops, kept = [], 0
            for i in order[p0:p0 + self.pool_images]:
                cr = self._crops(self._source(int(i), dev), rng)
                if cr:
                    crops += cr
                    kept += 1
            if not crops:
                continue
            if torch.is_tensor(crops[0]):   # in-memory sources live on the GPU: the crops are device-side slices
                u8 = torch.stack(crops)
            else:
                u8 = torch.from_numpy(np.stack(crops)).pin_memory().to(dev, non_blocking=True)
            # global sample index of the pool's first crop: unique per (rank, epoch, position), aligned to c
            sample_offset = ((self.rank << 40) + (int(e) * shard_len + p0)) * c
            clean, noisy = self.trainer.prepare_data(u8, self.noise_cfg, self.seed, sample_offset)
            if carry_clean is not None:
                clean, noisy = torch.cat([carry_clean, clean]), torch.cat([carry_noisy, noisy])
            perm = torch.randperm(clean.shape[0], generator=gen, device=dev)
            nb = clean.shape[0] // self.batch_size
            for b in range(nb):
                idx = perm[b * self.batch_size:(b + 1) * self.batch_size]
                yield clean.index_select(0, idx), noisy.index_select(0, idx)
            rest = perm[nb * self.batch_size:]
            carry_clean, carry_noisy = (clean.index_select(0, rest), noisy.index_select(0, rest)) if rest.numel() else (None, None)


def dataset_builder(config: Dict, trainer: Optional[Trainer] = None, *, images=None, rank: int = 0,
                    world: int = 1) -> DatasetResults:
    """bfcnn/dataset.py:40-305.  `training` is an `ImagePipeline` (iterating it yields (clean, noisy) float32 CUDA
    batches like the reference's tf.data.Dataset; None when the config names no input source), `testing` is None as in
    the reference (:303), `prepare_data_fn(clean_u8, seed, sample_offset)` is the corruption function itself."""
    color_mode = str(config.get("color_mode", "rgb")).strip().lower()
    if color_mode not in ("rgb", "rgba", "grayscale"):
        raise ValueError('`color_mode` must be one of {"rgb", "rgba", "grayscale"}. ' f"Received: color_mode={color_mode}")
    if color_mode != "rgb":
        raise ValueError("only colour (3-channel) models are on the hot path")
    ncfg = noise_cfg_from_config(config)

    def prepare_data_fn(input_batch, seed: int = 0, sample_offset: int = 0, trainer_: Optional[Trainer] = None):
        t = trainer_ or trainer
        if t is None:
            raise RuntimeError("prepare_data_fn needs a Trainer: the corruption runs on the GPU, there is no CPU path")
        return t.prepare_data(input_batch, ncfg, seed, sample_offset)

    training = None
    if images is not None or config.get("inputs"):
        training = ImagePipeline(config, ncfg, images=images, rank=rank, world=world)
        if trainer is not None:
            training.bind(trainer)
    return DatasetResults(config=config, batch_size=config.get("batch_size", 32),
                          input_shape=config.get("input_shape", [256, 256, 3]), training=training, testing=None,
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
                        seed: int = 0, process_group=None) -> Trainer:
    """model_builder + loss_function_builder + optimizer_builder of train_loop.py:80-148 for this family.  The optimizer
    section is held to the reference's rules: a missing `type` means RMSprop there (optimizer.py:165), which this path
    does not implement, so it must say "Adam"."""
    from .weights import synthetic_variables
    arch = arch_from_config(pipeline_config)
    if variables is None:
        variables = synthetic_variables(arch, seed)
    train_cfg = dict(pipeline_config.get("train", {}))
    opt = dict(train_cfg.get("optimizer", {}))
    optimizer_config_check(opt)
    if "gpu_batches_per_step" in train_cfg:
        opt["gpu_batches_per_step"] = train_cfg["gpu_batches_per_step"]
    return Trainer(arch, variables, device=device, loss_config=pipeline_config.get("loss"), optimizer_config=opt,
                   process_group=process_group)
