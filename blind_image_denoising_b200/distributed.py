"""Multi-GPU partitioning of the hot path: one process per GPU.

Inference shards independent units -- whole images of a batch, or row strips of one large
frame, each strip carrying R = receptive-field-radius extra input rows that are RECOMPUTED,
never exchanged -- so there is no data-path collective.  Training is data parallel with one
all-reduce of the flat gradient vector (Trainer.apply_grads).  The reference has no
multi-device code at all (last row); these helpers are the new surface.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced shard [lo, hi) of n_items for `rank` of `world` (first n % world ranks get one extra)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def strip_for_rank(height: int, rank: int, world: int, radius: int) -> Tuple[int, int, int, int]:
    """Row strip of a frame for `rank`: (in_lo, in_hi, out_lo, out_hi).  Output rows [out_lo, out_hi) depend on
    input rows [out_lo - R, out_hi + R) only; the strip's own zero-padded edges pollute at most R rows, which
    are cropped.  Strips never need a neighbour's data."""
    out_lo, out_hi = shard_range(height, rank, world)
    return max(0, out_lo - radius), min(height, out_hi + radius), out_lo, out_hi


def denoise_rows(model, frame_u8, rank: int, world: int, *, pad_pow2: bool = None):
    """Denoise this rank's row strip of `frame_u8` [N,H,W,3] (numpy or torch) with `model` (a Denoiser).

    Returns (out_lo, out_hi, strip_u8[N, out_hi-out_lo, W, 3]).  Concatenating the strips of all ranks in rank
    order reproduces `model(frame_u8)` bit for bit.  With pad_pow2 (the reference's pad_to_power_of_2,
    utilities.py:736-751) the raw-zero canvas rows/columns that can reach the crop are materialised here so
    that every strip sees what the whole-frame call sees."""
    pad = model.pad_pow2 if pad_pow2 is None else bool(pad_pow2)
    is_torch = type(frame_u8).__module__.split(".")[0] == "torch"
    n, h, w, c = frame_u8.shape
    R = model.arch.receptive_radius
    in_lo, in_hi, out_lo, out_hi = strip_for_rank(h, rank, world, R)
    if out_hi <= out_lo:
        empty = frame_u8[:, 0:0]
        return out_lo, out_hi, empty
    strip = frame_u8[:, in_lo:in_hi]
    extra_rows = extra_cols = 0
    if pad:
        hc = 1 << max(0, int(np.ceil(np.log2(h)))) if h > 1 else 1
        wc = 1 << max(0, int(np.ceil(np.log2(w)))) if w > 1 else 1
        extra_cols = min(R, wc - w)
        if in_hi == h:
            extra_rows = min(R, hc - h)
        if extra_rows or extra_cols:
            if is_torch:
                import torch
                strip = torch.nn.functional.pad(strip, (0, 0, 0, extra_cols, 0, extra_rows))
            else:
                strip = np.pad(strip, ((0, 0), (0, extra_rows), (0, extra_cols), (0, 0)))
    if is_torch:
        strip = strip.contiguous()
    else:
        strip = np.ascontiguousarray(strip)
    out = model(strip, pad_pow2=False)
    return out_lo, out_hi, out[:, out_lo - in_lo:out_hi - in_lo, :w]


def allreduce_mean_(tensor, group=None):
    """In-place mean over ranks of a flat gradient tensor (sum all-reduce, then 1/world).  The only
    collective of the path; NCCL on GPUs, gloo in the CPU tests."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
            tensor.div_(world)
    return tensor


class NcclCommunicator:
    """A raw `ncclComm_t` for the C-ABI gradient exchange (`bfcnn_allreduce_grads`), created with ctypes on the
    libnccl.so.2 already loaded in the process (PyTorch's).  Non-PyTorch hosts pass their own ncclComm_t to the C ABI;
    this class exists so that the Python `Trainer` can drive the same entry point instead of `torch.distributed`.

    `NcclCommunicator.from_torch_group()` makes rank 0 draw the unique id and shares it through the (already
    initialised) torch.distributed group; `NcclCommunicator(rank, world, unique_id)` takes the 128 id bytes directly."""

    def __init__(self, rank: int, world: int, unique_id: bytes, device: int = 0):
        import ctypes
        import torch
        self._lib = self._load()

        class _Id(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]
        if len(unique_id) != 128:
            raise ValueError("an ncclUniqueId is 128 bytes")
        uid = _Id()
        ctypes.memmove(ctypes.byref(uid), unique_id, 128)
        self._lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _Id, ctypes.c_int]
        self._lib.ncclCommInitRank.restype = ctypes.c_int
        self.comm = ctypes.c_void_p()
        torch.cuda.set_device(device)
        r = self._lib.ncclCommInitRank(ctypes.byref(self.comm), int(world), uid, int(rank))
        if r != 0:
            raise RuntimeError(f"ncclCommInitRank failed with ncclResult {r}")
        self.rank, self.world = int(rank), int(world)

    @staticmethod
    def _load():
        import ctypes
        import torch  # noqa: F401  (loads the bundled libnccl.so.2 into the process)
        for name in ("libnccl.so.2", "libnccl.so"):
            try:
                return ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                continue
        raise ImportError("libnccl.so.2 not found")

    @classmethod
    def unique_id(cls) -> bytes:
        import ctypes
        lib = cls._load()
        buf = ctypes.create_string_buffer(128)
        lib.ncclGetUniqueId.argtypes = [ctypes.c_void_p]
        lib.ncclGetUniqueId.restype = ctypes.c_int
        r = lib.ncclGetUniqueId(buf)
        if r != 0:
            raise RuntimeError(f"ncclGetUniqueId failed with ncclResult {r}")
        return buf.raw

    @classmethod
    def from_torch_group(cls, device: int, group=None) -> "NcclCommunicator":
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(rank, world, box[0], device=device)

    def close(self):
        import ctypes
        if getattr(self, "comm", None):
            self._lib.ncclCommDestroy.argtypes = [ctypes.c_void_p]
            self._lib.ncclCommDestroy(self.comm)
            self.comm = None
