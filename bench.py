#!/usr/bin/env python
"""bench.py -- megapixels/s denoised on B200 (BASELINE.json metric), one JSON line.

Workload (BASELINE.json configs[2], the one the north-star target is quoted on):
resnet_color_1x18_bn_16x3x3 inference on synthetic 3840x2160x3 uint8 frames,
`--frames` frames per GPU per step (weak scaling: every rank denoises its own frames,
no data-path collective).  A "step" is one pass of the hot path over that batch.

  value     : device-resident uint8 in -> uint8 out, CUDA events, max over ranks
  e2e       : the same through bfcnn.load_model(name)(host array): pinned host buffers,
              H2D + kernels + D2H inside the timed region
  roofline  : dominant kernel = fused_pass_kernel (the fused conv stack); achieved =
              algorithmic FLOPs of the launches / their summed CUDA-event time
  cpu_baseline / --impl reference : the oracle's torch-CPU fp32 restatement of the
              reference path (TensorFlow cannot be installed here), all host threads,
              on a bounded crop of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL_NAME = "resnet_color_1x18_bn_16x3x3_256x256_l1_relu"
N_LAYERS = 18
FRAME_H, FRAME_W = 2160, 3840
METRIC = "megapixels/sec denoised"
UNIT = "MP/s"
WORKLOAD = f"{MODEL_NAME} inference on synthetic 3840x2160x3 uint8 frames (BASELINE configs[2])"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 10 ms; only the samples that fall inside the timed region
    (host timestamps around the CUDA-event bracket) are reported."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        inside = [ln for (ts, ln) in self.lines if t0 <= ts <= t1 + 0.03]
        if not inside:   # region shorter than a sampling period: the closest samples
            inside = [ln for (ts, ln) in sorted(self.lines, key=lambda p: min(abs(p[0] - t0), abs(p[0] - t1)))[:2]]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_mp_s(crop: int, reps: int, frames_seed: int = 0):
    """Oracle torch-CPU fp32 restatement on a crop x crop sample of frame 0 (kind 'port')."""
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from blind_image_denoising_b200 import Arch, synthetic_variables
    from oracle import bfcnn_oracle as O
    v = synthetic_variables(Arch(no_layers=N_LAYERS), 0)
    rng = np.random.default_rng(frames_seed)
    x = rng.integers(0, 256, size=(1, crop, crop, 3), dtype=np.uint8)
    O.denoise_fp32_cpu(v, x[:, :64, :64], pad_pow2=False)  # warm-up
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.denoise_fp32_cpu(v, x, pad_pow2=False)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return crop * crop / 1e6 / med, med, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    crop = args.cpu_crop
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    from blind_image_denoising_b200 import Arch, synthetic_variables
    from oracle import bfcnn_oracle as O
    v = synthetic_variables(Arch(no_layers=N_LAYERS), 0)
    rng = np.random.default_rng(0)
    x = rng.integers(0, 256, size=(1, crop, crop, 3), dtype=np.uint8)
    for _ in range(max(1, args.warmup)):
        O.denoise_fp32_cpu(v, x[:, :128, :128], pad_pow2=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.denoise_fp32_cpu(v, x, pad_pow2=False)
    dt = time.perf_counter() - t0
    mp_s = args.steps * crop * crop / 1e6 / dt
    sample = f"{crop}x{crop} crop of one synthetic 3840x2160 frame per step, fp32 torch-CPU (oneDNN) restatement"
    line = {
        "impl": "reference", "metric": METRIC, "value": mp_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "no_layers": N_LAYERS,
                   "note": "reference TF path is not installable (tensorflow==2.13.1, no wheel for py3.12, no network); "
                           "this is the oracle's CPU restatement of it"},
        "cpu_baseline": {"value": mp_s, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": mp_s, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    _emit(line)
    return 0


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else a library prints to fd 1 (e.g. NCCL's version
    banner) was redirected to stderr at start-up."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4, help="4K frames per GPU per step")
    ap.add_argument("--precision", default="f16", choices=["f16", "f16_mma_sync", "f16x3", "f16x3_mma_sync", "fp32"])
    ap.add_argument("--cpu-crop", type=int, default=768)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the secondary precision modes")
    ap.add_argument("--no-training", action="store_true", help="skip the secondary training-step workload")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import bfcnn  # the drop-in alias; load_model reads the (synthetic) TensorBundle under pretrained/
    from blind_image_denoising_b200 import Arch

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    arch = Arch(no_layers=N_LAYERS)
    peaks = load_peaks()
    F = args.frames
    mp_per_step_rank = F * FRAME_H * FRAME_W / 1e6
    # synthetic frames: frame f of rank r uses seed r*F+f
    host = np.stack([np.random.default_rng(rank * F + f).integers(0, 256, size=(FRAME_H, FRAME_W, 3), dtype=np.uint8)
                     for f in range(F)])
    h_in = torch.from_numpy(host).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    d_in = h_in.cuda()
    d_out = torch.empty_like(d_in)

    def measure(precision: str, steps: int, warmup: int):
        model = bfcnn.load_model(MODEL_NAME, device=local_rank, precision=precision, pad_pow2=False)
        sampler = ClockSampler(local_rank)
        sampler.start()            # nvidia-smi needs ~100 ms to start: it runs through the warm-up
        for _ in range(warmup):
            model(d_in, out=d_out)
        barrier()
        l0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stack_ms = 0.0
        t_host0 = time.time()
        e0.record()
        for _ in range(steps):
            model(d_in, out=d_out)
        e1.record()
        barrier()
        t_host1 = time.time()
        clocks = sampler.stop(t_host0, t_host1)
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = model.launch_count() - l0
        # kernel-only time of the fused stack (events inside the library, same stream)
        for _ in range(3):
            model(d_in, out=d_out)
            stack_ms += model.last_stack_ms()
        stack_ms /= 3
        # e2e: host (pinned) in -> host out through the public API.  Every step's H2D copy, conv stack and D2H copy are
        # inside the timed region; two model instances (PipelinedDenoiser, depth 2) let the copies of one step overlap the
        # conv stack of the other, each call still returns only after its own result is in host memory.
        model.close()
        from blind_image_denoising_b200 import PipelinedDenoiser
        pipe = PipelinedDenoiser(lambda: bfcnn.load_model(MODEL_NAME, device=local_rank, precision=precision, pad_pow2=False), depth=2)
        h_outs = [h_out, torch.empty_like(h_in).pin_memory()]
        e2e_steps = max(2, min(steps, 6))
        for _ in pipe.map([h_in] * 4, outs=[h_outs[i % 2] for i in range(4)]):
            pass
        barrier()
        t0 = time.perf_counter()
        for _ in pipe.map([h_in] * e2e_steps, outs=[h_outs[i % 2] for i in range(e2e_steps)]):
            pass
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        pipe.close()
        return {"ms": ms, "launches": launches, "stack_ms": stack_ms, "e2e_ms": e2e_ms, "e2e_steps": e2e_steps,
                "clocks": clocks}

    r = measure(args.precision, args.steps, args.warmup)
    value = world * mp_per_step_rank * args.steps / (r["ms"] / 1e3)
    e2e_value = world * mp_per_step_rank * r["e2e_steps"] / (r["e2e_ms"] / 1e3)

    # roofline of the dominant kernel (the fused conv-stack pass / FP32 conv layer)
    alg_flops = arch.flops_per_pixel() * mp_per_step_rank * 1e6
    passes = {"f16": (N_LAYERS + 1) // 2, "f16_mma_sync": (N_LAYERS + 1) // 2, "f16x3": N_LAYERS,
              "f16x3_mma_sync": N_LAYERS, "fp32": 2 * N_LAYERS + 2}[args.precision]
    achieved = alg_flops / (r["stack_ms"] / 1e3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    roofline = {
        "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
        "traffic": None, "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
        "kernel": {"f16": "ustream::stream_pass_kernel (tcgen05 row-streaming stack; + one base_conv3_mma_kernel launch inside the timed stack)",
                   "f16_mma_sync": "fused_pass_kernel<1>",
                   "f16x3": "ustream3::stream_pass_kernel (tcgen05 row-streaming stack, fp16 hi/lo operand parts)",
                   "f16x3_mma_sync": "fused_pass_kernel<2>",
                   "fp32": "conv3x3_c16_kernel"}[args.precision],
        "launches_per_step": passes, "avg_launch_ms": r["stack_ms"] / passes,
        "algorithmic_flops_per_launch": alg_flops / passes,
        "note": "achieved = algorithmic FLOPs (167968/px, head un-collapsed, no halo recompute) / CUDA-event time of the stack launches",
    }
    prof = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                per_frame = json.load(f).get(args.precision)   # ncu --set full capture of a ONE-frame launch
            if per_frame is not None:   # the feature map goes once in and once out per pass: linear in the frames of a launch
                roofline["traffic"] = per_frame * F
                roofline["traffic_note"] = (f"dram bytes read + written per launch = {per_frame} B measured by ncu on a 1-frame launch "
                                            f"x {F} frames per launch (algorithmic: {F * 2160 * 3840 * 64} B, the fp16 NHWC16 map in and out)")
        except Exception:
            pass

    modes = {}
    if not args.no_modes:   # every rank takes part (the barriers are collective)
        for prec in ("f16", "f16_mma_sync", "f16x3", "f16x3_mma_sync", "fp32"):
            if prec == args.precision:
                continue
            steps = 2 if prec == "fp32" else max(2, args.steps // 2)
            rr = measure(prec, steps, 3)
            modes[prec] = {"value": world * mp_per_step_rank * steps / (rr["ms"] / 1e3), "unit": UNIT,
                           "tflops_algorithmic": alg_flops / (rr["stack_ms"] / 1e3) / 1e12}

    # secondary workload: BASELINE configs[3]/[4] -- training step (corruption + forward/backward + all-reduce + Adam)
    training = None
    if not args.no_training:
        from blind_image_denoising_b200 import synthetic_variables, _native
        from blind_image_denoising_b200.training import Trainer
        t_layers = 6 if world == 1 else 18
        t_arch = Arch(no_layers=t_layers)
        tr = Trainer(t_arch, synthetic_variables(t_arch, 0), device=local_rank,
                     optimizer_config={"gradient_clipping_by_norm": 1.0})
        clean_u8 = torch.from_numpy(np.random.default_rng(1000 + rank).integers(0, 256, size=(32, 256, 256, 3), dtype=np.uint8)).cuda()
        ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)
        def train_once(step):
            clean, noisy = tr.prepare_data(clean_u8, ncfg, 0, (step * world + rank) * 32)
            _, _, _, g = tr.train_step_single_gpu(clean, noisy)
            tr.apply_grads(g)
        for i in range(2):
            train_once(i)
        barrier()
        l0 = tr.launch_count()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        tsteps = 3
        for i in range(tsteps):
            train_once(2 + i)
        t1e.record()
        barrier()
        tms = max_over_ranks(t0e.elapsed_time(t1e)) / tsteps
        training = {"workload": f"resnet_color_1x{t_layers} training step: corruption + fwd/bwd (BN batch stats, hinged MAE, L1/L2 reg) + "
                                f"{'NCCL all-reduce + ' if world > 1 else ''}Adam, batch 32 x 256x256x3 per GPU (BASELINE configs[{3 if world == 1 else 4}])",
                    "value": world * 32 * 256 * 256 / 1e6 / (tms / 1e3), "unit": UNIT, "ms_per_step": tms, "dtype": "f32",
                    "gpu_launches_per_step": int((tr.launch_count() - l0) / tsteps)}
        tr.close()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mp_s, med, cores = cpu_reference_mp_s(args.cpu_crop, reps=3)
        cpu = {"value": mp_s, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_crop}x{args.cpu_crop} crop of frame 0, median of 3 (fp32 torch-CPU restatement of the reference TF path; TF not installable)"}

    if rank == 0:
        parity = {"f16": "fp16 operands / fp32 accumulate (tcgen05): max-abs <= 2.0, mean-abs <= 0.25 (0-255) vs fp64 oracle (stated bf16-class bound)",
                  "f16_mma_sync": "fp16 operands / fp32 accumulate (mma.sync baseline): max-abs <= 2.0, mean-abs <= 0.25",
                  "f16x3": "fp16 hi/lo split, 3 tcgen05 MMAs per product: max-abs <= 0.5, mean-abs <= 0.05 (the fp32 gate)",
                  "f16x3_mma_sync": "fp16 hi/lo split, 3 mma.sync per product: max-abs <= 0.5, mean-abs <= 0.05",
                  "fp32": "FP32 FFMA: max-abs <= 0.5, mean-abs <= 0.05"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms"] / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_gpu_per_step": F, "no_layers": N_LAYERS, "weights": "synthetic seed 0 (reference ships none, SURVEY F2)",
                       "parity": parity[args.precision], "pad_pow2": False,
                       "l2": f"working set {F * 24.9 * 2 + F * 8.29 * 32 * 2:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "parallelism": f"frames sharded over {world} GPU(s), no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h_in.numel()) * world,
                    "d2h_bytes_per_step": int(h_out.numel()) * world,
                    "api": "bfcnn.load_model(name)(pinned host uint8) on two model instances (PipelinedDenoiser.map, depth 2): every call is "
                           "synchronous, the copies of one step overlap the conv stack of the other"},
            "gpu_launches": int(r["launches"]),
            "clocks": r["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "modes": modes,
            "training": training,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
