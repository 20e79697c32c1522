"""Names the limiter of the e2e leg at N ranks: concurrent pinned H2D + D2H copies of the bench's step size on every rank,
no kernels at all.  torchrun --nproc-per-node N tools/pcie_probe.py"""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nbytes = 4 * 2160 * 3840 * 3
h_in, h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory(), torch.empty(nbytes, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(nbytes, dtype=torch.uint8, device="cuda"), torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for mode in ("h2d", "d2h", "both"):
    for it in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    gbs = torch.tensor([10 * nbytes / dt / 1e9], device="cuda")
    if world > 1:
        dist.all_reduce(gbs)
    res[mode] = float(gbs.item())
if rank == 0:
    print(f"{world} ranks, {nbytes / 1e6:.0f} MB per copy: aggregate GB/s per direction  h2d alone {res['h2d']:.1f}, d2h alone {res['d2h']:.1f}, "
          f"both at once {res['both']:.1f} each ({res['both'] / world:.1f} per GPU; the f16 arm needs {nbytes / 8.6e-3 / 1e9:.1f} per GPU per direction)")
if world > 1:
    dist.destroy_process_group()
