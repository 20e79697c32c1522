"""Optimizer / schedule builders with the reference's names and argument checks
(reference bfcnn/optimizer.py:21-224).  The arithmetic of an update is the fused Adam kernel of
libbfcnn_b200.so (`bfcnn_adam_step`); this file parses configs and evaluates the closed-form schedules.

  deep_supervision_schedule_builder   optimizer.py:21-78   (weights per model output; one output -> [1.0])
  schedule_builder                    optimizer.py:83-139  (keras ExponentialDecay / CosineDecay / CosineDecayRestarts)
  optimizer_builder                   optimizer.py:145-224 -> (optimizer, lr_schedule)

Only `type: "Adam"` exists on the B200 path.  The reference's DEFAULT type is RMSprop
(optimizer.py:165) and it also offers Adadelta, `amsgrad`, per-variable `clipvalue` / `clipnorm`: a config that asks
for any of those raises here instead of silently training with a different optimizer.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import numpy as np

TYPE_STR = "type"      # reference bfcnn/constants.py
CONFIG_STR = "config"


def deep_supervision_schedule_builder(config: Dict, no_outputs: int) -> Callable[[float], np.ndarray]:
    """optimizer.py:21-78, same names, checks and formulas."""
    if not isinstance(config, dict):
        raise ValueError("config must be a dictionary")
    if no_outputs <= 0:
        raise ValueError("no_outputs must be positive integer")
    schedule_type = config.get(TYPE_STR, None)
    if schedule_type is None:
        raise ValueError("schedule_type cannot be None")
    if not isinstance(schedule_type, str):
        raise ValueError("schedule_type must be a string")
    schedule_type = schedule_type.strip().lower()
    ramp = np.array(list(range(1, no_outputs + 1))).astype(np.float32)
    ramp = ramp / np.sum(ramp)
    if schedule_type == "constant_equal":
        return lambda percentage_done=0.0: np.array([1.0] * no_outputs) / float(no_outputs)
    if schedule_type == "constant_low_to_high":
        return lambda percentage_done=0.0: ramp.copy()
    if schedule_type == "constant_high_to_low":
        return lambda percentage_done=0.0: ramp[::-1].copy()
    if schedule_type == "linear_low_to_high":
        return lambda percentage_done=0.0: ramp * (1.0 - percentage_done) + ramp[::-1] * percentage_done
    if schedule_type == "non_linear_low_to_high":
        def schedule(percentage_done: float = 0.0):
            x = np.clip(np.tanh(2.5 * percentage_done), a_min=0.0, a_max=1.0)
            return ramp * (1.0 - x) + ramp[::-1] * x
        return schedule
    raise ValueError(f"don't know how to handle deep supervision schedule_type [{schedule_type}]")


def schedule_builder(config: Dict) -> Callable[[int], float]:
    """optimizer.py:83-139: step -> learning rate (closed forms of the three keras schedules)."""
    if not isinstance(config, dict):
        raise ValueError("config must be a dictionary")
    schedule_type = config.get(TYPE_STR, None)
    if schedule_type is None:
        raise ValueError("schedule_type cannot be None")
    if not isinstance(schedule_type, str):
        raise ValueError("schedule_type must be a string")
    params = config.get(CONFIG_STR, {})
    schedule_type = schedule_type.strip().lower()
    if schedule_type == "exponential_decay":
        lr0, rate, steps = float(params["learning_rate"]), float(params["decay_rate"]), float(params["decay_steps"])
        return lambda step: lr0 * rate ** (step / steps)
    if schedule_type == "cosine_decay":
        lr0 = float(params["learning_rate"])
        steps, alpha = float(params["decay_steps"]), float(params.get("alpha", 0.0001))

        def cosine(step):
            p = min(step, steps) / steps
            return lr0 * ((1 - alpha) * 0.5 * (1 + math.cos(math.pi * p)) + alpha)
        return cosine
    if schedule_type == "cosine_decay_restarts":
        lr0 = float(params["learning_rate"])
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


class AdamOptimizer:
    """What `tf.keras.optimizers.Adam(...)` is to the reference's train loop: hyper-parameters, `iterations`, and
    `apply_gradients`.  The update itself is `bfcnn_adam_step` on the trainer's handle."""

    name = "Adam"

    def __init__(self, learning_rate: Callable[[int], float], beta_1: float = 0.9, beta_2: float = 0.999,
                 epsilon: float = 1e-7, global_clipnorm: Optional[float] = None):
        self.learning_rate_schedule = learning_rate
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.global_clipnorm = float(global_clipnorm) if global_clipnorm else 0.0
        self.iterations = 0

    @property
    def learning_rate(self) -> float:
        return float(self.learning_rate_schedule(self.iterations))

    def get_config(self) -> Dict:
        return {"name": self.name, "beta_1": self.beta_1, "beta_2": self.beta_2, "epsilon": self.epsilon,
                "global_clipnorm": self.global_clipnorm or None, "amsgrad": False}

    def apply_gradients(self, trainer, grads=None) -> float:
        """`optimizer.apply_gradients(zip(grads, trainable_variables))` (train_loop.py:314-321); returns the rate used."""
        trainer.bind_optimizer(self)
        return trainer.apply_grads(grads)


def optimizer_config_check(config: Dict) -> None:
    """Reject what optimizer.py:145-224 accepts but the B200 path does not implement."""
    if not isinstance(config, dict):
        raise ValueError("config must be a dictionary")
    optimizer_type = str(config.get("type", "RMSprop")).strip().upper()   # the reference's default is RMSprop (:165)
    if optimizer_type not in ("RMSPROP", "ADAM", "ADADELTA"):
        raise ValueError(f"don't know how to handle optimizer_type: [{optimizer_type}]")
    if optimizer_type != "ADAM":
        raise ValueError(f"optimizer_type [{optimizer_type}] is not on the B200 path: only the fused Adam kernel exists "
                         "(set train.optimizer.type = \"Adam\"; the reference's default is RMSprop)")
    if config.get("amsgrad", False):
        raise ValueError("amsgrad=True is not on the B200 path")
    if config.get("gradient_clipping_by_value", None):
        raise ValueError("gradient_clipping_by_value (keras clipvalue) is not on the B200 path; "
                         "use gradient_clipping_by_norm (global_clipnorm)")
    if config.get("gradient_clipping_by_norm_local", None):
        raise ValueError("gradient_clipping_by_norm_local (keras per-variable clipnorm) is not on the B200 path; "
                         "use gradient_clipping_by_norm (global_clipnorm)")


def optimizer_builder(config: Dict) -> Tuple[AdamOptimizer, Callable[[int], float]]:
    """optimizer.py:145-224 -> (optimizer, lr_schedule)."""
    optimizer_config_check(config)
    lr_schedule = schedule_builder(config=config["schedule"])
    optimizer = AdamOptimizer(learning_rate=lr_schedule, beta_1=config.get("beta_1", 0.9), beta_2=config.get("beta_2", 0.999),
                              epsilon=config.get("epsilon", 1e-07),
                              global_clipnorm=config.get("gradient_clipping_by_norm", None))
    return optimizer, lr_schedule
