#!/usr/bin/env python
"""bench.py -- megapixels/s denoised on B200 (BASELINE.json metric), one JSON line.

Workload (BASELINE.json configs[2], the one the north-star target is quoted on):
resnet_color_1x18_bn_16x3x3 inference on synthetic 3840x2160x3 uint8 frames, `--frames` frames per GPU per step (weak
scaling: every rank denoises its own frames, no data-path collective), through the drop-in callable
`bfcnn.load_model(name)` with its default semantics (pow2 canvas, module_denoiser.py:56-68).  A "step" is one pass of
the hot path over that batch.

  value     : device-resident uint8 in -> uint8 out, CUDA events around the K timed steps, max over ranks
  e2e       : the same K steps through bfcnn.load_model(name)(pinned host uint8): H2D + kernels + D2H inside the timed region
  roofline  : tensor pipe.  achieved = algorithmic FLOPs of a step / ms_per_step of THE TIMED LOOP (so frac x peak x
              ms_per_step reproduces the algorithmic FLOPs); `dominant_kernel` = the fused pass kernel alone, its launches
              timed with CUDA events inside this run (bfcnn_set_kernel_timing, a few extra steps after the timed loop)
  arms      : the fp32-grade arm (precision f16x3, the default of bfcnn.load_model) as a full record of its own:
              value, ms_per_step, e2e, roofline
  training  : BASELINE configs[3] / [4]: 1x18 data-parallel step on every N (and 1x6 at N = 1), with a dp_check at N > 1
  small_images : BASELINE configs[0] / [1] (1 and 64 images of 256 x 256), device-resident, both tensor-core precisions
  strong    : ONE 4K frame split into row strips over the N ranks (BASELINE configs[2] as written), per-frame latency
  cpu_baseline / --impl reference : the oracle's torch-CPU fp32 restatement of the reference path (TensorFlow cannot be
              installed here), all host threads, on whole 3840x2160 frames with the reference's own pow2 canvas.
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


def WEIGHTS_NOTE() -> str:
    """Where the weights of the benchmarked model directory come from (its pipeline.json says: the reference snapshot ships
    no resnet weights; throughput does not depend on the values)."""
    try:
        with open(os.path.join(ROOT, "blind_image_denoising_b200", "pretrained", MODEL_NAME, "pipeline.json")) as f:
            return str(json.load(f).get("weights", "unknown"))[:160]
    except Exception:
        return "unknown"


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
    """SM clock and throttle reasons DURING the timed region: `nvidia-smi -lms 10` (line-buffered through stdbuf: into a pipe
    it otherwise flushes in 4 KB blocks, and the lines of a 0.2 s region arrive after it or never) and, beside it, the same
    NVML counters polled from a thread every 5 ms.  Only samples whose host timestamp falls inside the region are reported;
    nvidia-smi's are preferred, NVML's fill in when it delivered none (`source` says which)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.proc, self.lines, self.nvml, self._stop = index, None, [], [], False

    def start(self):
        try:
            cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10", "-i", str(self.index)]
            if os.path.exists("/usr/bin/stdbuf"):
                cmd = ["/usr/bin/stdbuf", "-oL"] + cmd
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # the process may see a subset of the GPUs (CUDA_VISIBLE_DEVICES): NVML indexes the physical ones
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = self.index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.index < len(ids) and ids[self.index].isdigit():
                    phys = int(ids[self.index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nv = pynvml
            self.t2 = threading.Thread(target=self._poll, daemon=True)
            self.t2.start()
        except Exception:
            self._nv = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def _poll(self):
        nv = self._nv
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = None
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                self.nvml.append((time.time(), sm, mx, [nm for nm, b in zip(self.NAMES, bits) if r & b]))
            except Exception:
                break
            time.sleep(0.005)

    def stop(self, t0: float, t1: float):
        self._stop = True
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons, source = [], None, set(), "nvidia-smi -lms 10"
        for ln in [ln for (ts, ln) in self.lines if t0 <= ts <= t1 + 0.03]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            source = "NVML (nvmlDeviceGetClockInfo / CurrentClocksThrottleReasons every 5 ms; nvidia-smi delivered no sample inside the region)"
            for (ts, v, m, rs) in self.nvml:
                if t0 <= ts <= t1 + 0.005:
                    sm.append(v); mx = m; reasons.update(rs)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock sample inside the timed region"], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "source": source}


# --------------------------------------------------------------------------------------
# CPU arm: the oracle's fp32 restatement of DenoiserModule.__call__ on whole frames (pow2 canvas as the reference pads it)
# --------------------------------------------------------------------------------------
def _cpu_setup():
    import numpy as np
    import torch
    torch.set_num_threads(os.cpu_count() or 1)   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    from blind_image_denoising_b200 import Arch, synthetic_variables
    from oracle import bfcnn_oracle as O
    v = synthetic_variables(Arch(no_layers=N_LAYERS), 0)
    frame = np.random.default_rng(0).integers(0, 256, size=(1, FRAME_H, FRAME_W, 3), dtype=np.uint8)
    O.denoise_fp32_cpu(v, frame[:, :256, :256], pad_pow2=True)   # library warm-up (thread pool, primitive cache)
    return O, v, frame, torch.get_num_threads()


CPU_SAMPLE = ("one whole synthetic 3840x2160 frame per step, computed as the reference does: zero-padded to the 4096x4096 "
              "canvas (module_denoiser.py:56), fp32 torch-CPU (oneDNN) restatement of the TF path, cropped back")


def cpu_reference_mp_s(reps: int = 1):
    O, v, frame, threads = _cpu_setup()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        O.denoise_fp32_cpu(v, frame, pad_pow2=True)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return FRAME_H * FRAME_W / 1e6 / med, med, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    O, v, frame, threads = _cpu_setup()
    for _ in range(1 if args.warmup > 0 else 0):   # one whole-frame warm-up step (each costs ~10 s of all host cores)
        O.denoise_fp32_cpu(v, frame, pad_pow2=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.denoise_fp32_cpu(v, frame, pad_pow2=True)
    dt = time.perf_counter() - t0
    mp_s = args.steps * FRAME_H * FRAME_W / 1e6 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": mp_s, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "no_layers": N_LAYERS, "frames_per_step": 1, "pad_pow2": True,
                   "weights": WEIGHTS_NOTE(),
                   "note": "reference TF path is not installable (tensorflow==2.13.1, no wheel for py3.12, no network); "
                           "this is the oracle's CPU restatement of it; a step is ONE frame (the GPU arm's step is "
                           "frames_per_gpu_per_step frames of the same size: the metric is per pixel)"},
        "cpu_baseline": {"value": mp_s, "unit": UNIT, "cores": threads, "kind": "port", "sample": CPU_SAMPLE},
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

PARITY = {"f16": "fp16 operands / fp32 accumulate (tcgen05): max-abs <= 2.0, mean-abs <= 0.25 (0-255) vs fp64 oracle (stated bf16-class bound); on the shipped TRAINED 1x18 weights and natural images: max-abs 0.35, mean-abs 0.031 (tests/test_pretrained_gpu.py; not a bound, DESIGN.md 4.4)",
          "f16x3": "fp16 hi/lo split, 3 tcgen05 MMAs per product: max-abs <= 0.5, mean-abs <= 0.05 (the fp32 gate); default of bfcnn.load_model",
          "fp32": "FP32 FFMA: max-abs <= 0.5, mean-abs <= 0.05"}
KERNEL = {"f16": "ustream::stream_pass_kernel (tcgen05 row-streaming stack, 2 residual blocks = 4 convs per launch)",
          "f16x3": "ustream3::stream_pass_kernel (tcgen05 row-streaming stack, fp16 hi/lo operand parts, 1 residual block per launch)",
          "fp32": "conv3x3_c16_kernel (FP32 FFMA, one conv per launch)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=4, help="4K frames per GPU per step")
    ap.add_argument("--precision", default="f16", choices=["f16", "f16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-arms", action="store_true", help="skip the fp32-grade arm")
    ap.add_argument("--no-training", action="store_true", help="skip the training-step workloads")
    ap.add_argument("--no-strong", action="store_true", help="skip the strip-sharded single-frame leg")
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
    from blind_image_denoising_b200 import Arch, PipelinedDenoiser

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
    alg_flops_step = arch.flops_per_pixel() * mp_per_step_rank * 1e6          # per rank
    # synthetic frames: frame f of rank r uses seed r*F+f
    host = np.stack([np.random.default_rng(rank * F + f).integers(0, 256, size=(FRAME_H, FRAME_W, 3), dtype=np.uint8)
                     for f in range(F)])
    h_in = torch.from_numpy(host).pin_memory()
    h_outs = [torch.empty_like(h_in).pin_memory(), torch.empty_like(h_in).pin_memory()]
    d_in = h_in.cuda()
    d_out = torch.empty_like(d_in)

    def load(precision):
        return bfcnn.load_model(MODEL_NAME, device=local_rank, precision=precision, allow_synthetic=True)   # pad_pow2=True

    def measure(precision: str, steps: int, warmup: int):
        model = load(precision)
        sampler = ClockSampler(local_rank)
        sampler.start()            # nvidia-smi needs ~100 ms to start: it runs through the warm-up
        for _ in range(warmup):
            model(d_in, out=d_out)
        barrier()
        l0 = model.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
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
        # the kernels of a step one by one (CUDA events around every launch, inside this run, same stream, right after
        # the timed loop while the GPU is in the same thermal / power state): 3 more steps
        model.set_kernel_timing(True)
        kt = []
        for _ in range(3):
            model(d_in, out=d_out)
            kt.append(model.kernel_times())
        model.set_kernel_timing(False)
        gaps_ms = float(np.mean([sum(m for k, m in kt[r] if k < 0) for r in range(3)]))
        kt = [[(k, m) for k, m in one if k >= 0] for one in kt]
        kinds = [k for k, _ in kt[0]]
        per_launch = [float(np.mean([kt[r][i][1] for r in range(3)])) for i in range(len(kinds))]
        # e2e: host (pinned) in -> host out through the public API, the SAME number of steps.  Every step's H2D copy, conv
        # stack and D2H copy are inside the timed region; two model instances (PipelinedDenoiser, depth 2) let the copies
        # of one step overlap the conv stack of the other, each call still returns only after its own result is in host memory.
        model.close()
        pipe = PipelinedDenoiser(lambda: load(precision), depth=2)
        for _ in pipe.map([h_in] * 4, outs=[h_outs[i % 2] for i in range(4)]):
            pass
        barrier()
        t0 = time.perf_counter()
        for _ in pipe.map([h_in] * steps, outs=[h_outs[i % 2] for i in range(steps)]):
            pass
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        pipe.close()
        return {"ms": ms, "steps": steps, "launches": launches, "e2e_ms": e2e_ms, "clocks": clocks, "kinds": kinds,
                "per_launch_ms": per_launch, "gaps_ms": gaps_ms}

    def record(precision: str, r: dict):
        steps = r["steps"]
        ms_step = r["ms"] / steps
        value = world * mp_per_step_rank * steps / (r["ms"] / 1e3)
        achieved = alg_flops_step / (ms_step / 1e3) / 1e12
        peak = peaks["bf16_tflops_sustained"]
        pass_ms = [m for k, m in zip(r["kinds"], r["per_launch_ms"]) if k in (1, 2)]
        base_ms = sum(m for k, m in zip(r["kinds"], r["per_launch_ms"]) if k == 0)
        mid_ms = [m for k, m in zip(r["kinds"], r["per_launch_ms"]) if k == 1]
        convs_per_pass = 4 if precision == "f16" else 2
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a long step)",
            "basis": "algorithmic FLOPs of a step (167968/px x pixels, head un-collapsed, no halo recompute, no canvas band) / "
                     "ms_per_step of the timed loop",
            "kernel": KERNEL[precision],
            "launches_per_step": r["launches"] / steps,
        }
        if mid_ms:
            flops_launch = 2.0 * 9 * 16 * 16 * convs_per_pass * mp_per_step_rank * 1e6      # the convs of one middle pass
            avg = float(np.mean(mid_ms))
            roofline["dominant_kernel"] = {
                "launches_per_step": len(pass_ms), "avg_launch_ms": avg, "last_pass_ms": pass_ms[-1], "base_conv_ms": base_ms,
                "kernel_ms_per_step": sum(pass_ms) + base_ms, "idle_between_launches_ms_per_step": r["gaps_ms"],
                "algorithmic_flops_per_launch": flops_launch, "achieved": flops_launch / (avg / 1e3) / 1e12,
                "frac": flops_launch / (avg / 1e3) / 1e12 / peak,
                "how": "CUDA events around every launch of 3 steps run right after the timed loop (bfcnn_set_kernel_timing)"}
        if precision == "f16x3":
            roofline["issued_mma_factor"] = 3
            roofline["frac_of_fp32_grade_tensor_peak"] = achieved / (peak / 3.0)
            roofline["note"] = ("fp32-grade results from three fp16 MMAs per product (hi*hi + hi*lo + lo*hi): the tensor pipe "
                                "issues 3x the algorithmic FLOPs, so peak / 3 bounds this arithmetic")
        for name in ("r02_traffic.json", "r01_traffic.json"):
            prof = os.path.join(ROOT, "profiles", name)
            if os.path.exists(prof):
                try:
                    with open(prof) as f:
                        per_frame = json.load(f).get(precision)   # ncu --set full capture of a ONE-frame launch
                    if per_frame is not None:   # the feature map goes once in and once out per pass: linear in the frames of a launch
                        roofline["traffic"] = per_frame * F
                        roofline["traffic_note"] = (f"dram bytes read + written per launch of the dominant kernel = {per_frame} B "
                                                    f"(ncu, 1-frame launch, profiles/{name}) x {F} frames per launch")
                        break
                except Exception:
                    pass
        e2e_value = world * mp_per_step_rank * steps / (r["e2e_ms"] / 1e3)
        return {"value": value, "unit": UNIT, "dtype": precision, "ms_per_step": ms_step, "steps": steps,
                "gpu_launches": int(r["launches"]), "parity": PARITY[precision], "clocks": r["clocks"],
                "e2e": {"value": e2e_value, "unit": UNIT, "steps": steps, "h2d_bytes_per_step": int(h_in.numel()) * world,
                        "d2h_bytes_per_step": int(h_in.numel()) * world,
                        "api": "bfcnn.load_model(name)(pinned host uint8) on two model instances (PipelinedDenoiser.map, depth 2): "
                               "every call is synchronous, the copies of one step overlap the conv stack of the other"},
                "roofline": roofline}

    primary = record(args.precision, measure(args.precision, args.steps, args.warmup))

    arms = {}
    if not args.no_arms:   # every rank takes part (the barriers are collective)
        for prec in ("f16x3",) if args.precision != "f16x3" else ("f16",):
            arms[prec] = record(prec, measure(prec, max(3, args.steps // 2), 3))

    # ---- BASELINE configs[2] as written: ONE frame, row strips over the ranks (strong scaling, no collective)
    strong = None
    if not args.no_strong:
        from blind_image_denoising_b200.distributed import denoise_rows, strip_for_rank
        model = load(args.precision)
        frame = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(1, FRAME_H, FRAME_W, 3), dtype=np.uint8)).cuda()
        for _ in range(3):
            denoise_rows(model, frame, rank, world)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ssteps = max(5, args.steps)
        e0.record()
        for _ in range(ssteps):
            denoise_rows(model, frame, rank, world)
        e1.record()
        barrier()
        sms = max_over_ranks(e0.elapsed_time(e1)) / ssteps
        in_lo, in_hi, out_lo, out_hi = strip_for_rank(FRAME_H, rank, world, arch.receptive_radius)
        strong = {"workload": "one 3840x2160 frame split into row strips over the ranks (halo rows recomputed, never exchanged)",
                  "scaling": "strong", "frame_latency_ms": sms, "value": FRAME_H * FRAME_W / 1e6 / (sms / 1e3), "unit": UNIT,
                  "dtype": args.precision, "rows_owned_rank0": out_hi - out_lo, "rows_read_rank0": in_hi - in_lo,
                  "halo_overhead": (in_hi - in_lo) / max(out_hi - out_lo, 1) - 1.0}
        model.close()

    # ---- BASELINE configs[3] / [4]: training step (corruption + forward/backward + all-reduce + Adam), async steps
    training = None
    if not args.no_training:
        from blind_image_denoising_b200 import synthetic_variables, _native
        from blind_image_denoising_b200.training import Trainer

        def train_leg(t_layers: int, tsteps: int = 10):
            t_arch = Arch(no_layers=t_layers)
            comm = None
            if world > 1:   # the gradient exchange goes through the C ABI: bfcnn_allreduce_grads on a raw ncclComm_t
                from blind_image_denoising_b200.distributed import NcclCommunicator
                comm = NcclCommunicator.from_torch_group(local_rank)
            tr = Trainer(t_arch, synthetic_variables(t_arch, 0), device=local_rank,
                         optimizer_config={"gradient_clipping_by_norm": 1.0}, nccl_comm=comm)
            clean_u8 = torch.from_numpy(np.random.default_rng(1000 + rank).integers(0, 256, size=(32, 256, 256, 3), dtype=np.uint8)).cuda()
            ncfg = _native.NoiseCfg(5.0, 40.0, 0.05, 0.1, 1, 1, 0, 1)

            def train_once(step):
                clean, noisy = tr.prepare_data(clean_u8, ncfg, 0, (step * world + rank) * 32)
                _, _, _, g = tr.train_step_single_gpu(clean, noisy, sync=False)    # nothing read back: steps chain on the stream
                tr.apply_grads(g)
            for i in range(3):
                train_once(i)
            dp = None
            if world > 1:
                # dp_check: the all-reduced gradient equals the mean of the per-rank gradients gathered separately, and the
                # trainable variables stay identical on all ranks after the update
                clean, noisy = tr.prepare_data(clean_u8, ncfg, 0, (2 * world + rank) * 32)
                _, _, _, g = tr.train_step_single_gpu(clean, noisy, update_moving=False, sync=False)
                mine = g.clone()
                gathered = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(gathered, mine)
                mean = torch.stack(gathered).double().mean(0)
                red = mine.clone()
                import ctypes
                _native.check(tr._lib.bfcnn_allreduce_grads(tr.handle, red.data_ptr(), comm.comm,
                                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
                err = float(((red.double() / world) - mean).abs().max() / mean.abs().max())
                differ = float((gathered[0] - gathered[-1]).abs().max() / mean.abs().max())
                tr.apply_grads(g)
                from blind_image_denoising_b200.weights import flatten_variables, gather_trainables
                w = torch.from_numpy(gather_trainables(t_arch, flatten_variables(t_arch, tr.get_weights()))).cuda()
                ws = [torch.empty_like(w) for _ in range(world)]
                dist.all_gather(ws, w)
                same = all(bool(torch.equal(ws[0], x)) for x in ws[1:])
                dp = {"exchange": "bfcnn_allreduce_grads (C ABI, raw ncclAllReduce on the compute stream)",
                      "allreduce_vs_gathered_mean_rel_err": err, "rank_gradients_differ_rel": differ,
                      "trainables_identical_on_all_ranks": same, "ok": bool(err <= 1e-6 and differ > 1e-3 and same)}
                assert dp["ok"], dp
            barrier()
            l0 = tr.launch_count()
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record()
            for i in range(tsteps):
                train_once(3 + i)
            t1e.record()
            total_loss, _, dl = tr.last_losses()          # ONE device-to-host read after the timed steps
            barrier()
            tms = max_over_ranks(t0e.elapsed_time(t1e)) / tsteps
            px = 32 * 256 * 256
            # HBM floor of SURVEY 8d: the minimum save-for-backward set written once and read once
            bytes_px = 2 * (3 * t_layers + 1) * 16 * 4 + 3
            hbm_ms = px * bytes_px / (peaks["hbm_gbs"] * 1e9) * 1e3
            out = {"workload": f"resnet_color_1x{t_layers} training step: corruption + fwd/bwd (BN batch stats, hinged MAE, L1/L2 reg) + "
                               f"{'NCCL all-reduce + ' if world > 1 else ''}Adam, batch 32 x 256x256x3 per GPU "
                               f"(BASELINE configs[{3 if t_layers == 6 else 4}])",
                   "value": world * px / 1e6 / (tms / 1e3), "unit": UNIT, "ms_per_step": tms, "steps": tsteps, "dtype": "f32 (fp16 hi/lo tensor-core convs)",
                   "gpu_launches_per_step": int((tr.launch_count() - l0) / tsteps), "host_syncs_per_step": 0,
                   "final_total_loss": total_loss, "final_mae": dl["mae_loss"],
                   "roofline": {"bound": "hbm", "floor_ms": hbm_ms, "frac": hbm_ms / tms, "bytes_per_px": bytes_px,
                                "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "basis": "SURVEY 8d: (3N+1) fp32 16-channel maps written once + read once"}}
            if dp is not None:
                out["dp_check"] = dp
            tr.close()
            if comm is not None:
                comm.close()
            return out

        training = {"1x18": train_leg(18)}
        if world == 1:
            training["1x6"] = train_leg(6)

    # ---- BASELINE configs[0] / [1]: the 256 x 256 workloads of the metric (device-resident, the drop-in semantics), rank 0's GPU
    small = None
    if rank == 0 and not args.no_arms:
        small = {}
        for key, name, shape in (("configs[0]", "resnet_color_1x6_bn_16x3x3_256x256_l1_relu", (1, 256, 256, 3)),
                                 ("configs[1]", "resnet_color_1x12_bn_16x3x3_256x256_l1_relu", (64, 256, 256, 3))):
            xs = torch.from_numpy(np.random.default_rng(7).integers(0, 256, size=shape, dtype=np.uint8)).cuda()
            os_ = torch.empty_like(xs)
            rec = {"workload": f"{name} inference on {shape[0]}x256x256x3 uint8, device-resident"}
            for prec in ("f16", "f16x3"):
                ms_model = bfcnn.load_model(name, device=local_rank, precision=prec, allow_synthetic=True)
                for _ in range(5):
                    ms_model(xs, out=os_)
                torch.cuda.synchronize()
                reps = 50
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    ms_model(xs, out=os_)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                rec[prec] = {"ms_per_call": ms, "value": shape[0] * 256 * 256 / 1e6 / (ms / 1e3), "unit": UNIT}
                ms_model.close()
            small[key] = rec
    barrier()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mp_s, med, cores = cpu_reference_mp_s(reps=1)
        cpu = {"value": mp_s, "unit": UNIT, "cores": cores, "kind": "port", "sample": CPU_SAMPLE + f" ({med:.1f} s)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": primary["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": primary["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "frames_per_gpu_per_step": F, "no_layers": N_LAYERS, "weights": WEIGHTS_NOTE(),
                       "parity": PARITY[args.precision], "pad_pow2": True,
                       "api": "bfcnn.load_model(name)(uint8 [N,H,W,3]) with its defaults except precision",
                       "l2": f"working set {F * 24.9 * 2 + F * 8.52 * 32 * 2:.0f} MB per step > 126 MB L2 (no flush needed)",
                       "parallelism": f"frames sharded over {world} GPU(s), no collective"},
            "e2e": primary["e2e"],
            "gpu_launches": primary["gpu_launches"],
            "clocks": primary["clocks"],
            "roofline": primary["roofline"],
            "cpu_baseline": cpu,
            "arms": arms,
            "strong_scaling": strong,
            "training": training,
            "small_images": small,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
