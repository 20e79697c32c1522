"""Host-side multi-GPU logic on CPU: shard arithmetic, and the gradient all-reduce under world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest

from blind_image_denoising_b200.distributed import shard_range, strip_for_rank


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 64, 2160):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_strips_carry_the_receptive_field():
    H, R = 2160, 37       # 4K frame, 1x18 model (270 rows + 74 halo rows at 8 GPUs)
    for world in (1, 2, 4, 8):
        rows = []
        for r in range(world):
            in_lo, in_hi, out_lo, out_hi = strip_for_rank(H, r, world, R)
            assert in_lo == max(0, out_lo - R) and in_hi == min(H, out_hi + R)
            rows += list(range(out_lo, out_hi))
        assert rows == list(range(H))
    in_lo, in_hi, out_lo, out_hi = strip_for_rank(H, 3, 8, R)
    assert (out_hi - out_lo, in_hi - in_lo) == (270, 344)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from blind_image_denoising_b200.distributed import allreduce_mean_, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # data-parallel gradient exchange: each rank holds the flat gradient of its own micro-batch
    g = torch.full((84272,), float(rank + 1))        # 1x18 trainable count
    allreduce_mean_(g)
    # frames sharded with no collective: every rank owns a disjoint range, all ranges together cover the batch
    lo, hi = shard_range(64, rank, world)
    owned = torch.zeros(64)
    owned[lo:hi] = 1
    dist.all_reduce(owned)
    q.put((rank, float(g[0]), float(g[-1]), float(owned.min()), float(owned.max())))
    dist.destroy_process_group()


def test_gloo_world2_gradient_mean_and_sharding():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, g0, g1, omin, omax in res:
        assert g0 == 1.5 and g1 == 1.5
        assert omin == 1.0 and omax == 1.0
