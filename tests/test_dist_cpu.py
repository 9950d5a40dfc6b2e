"""World-size-2 gloo test (CPU) of the multi-GPU host logic: one weight broadcast, contiguous pair shards,
flow gather in global order.  The GPU run uses the same functions over NCCL."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flownet2_b200 import parallel as P


def test_shard_bounds_cover_everything():
    for total in (1, 4, 5, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert P.shard_bounds(32, 8, 3) == (12, 16)          # config 4: 32 pairs over 8 GPUs, 4 each


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # weights: only rank 0 has them before the broadcast
        arena = torch.arange(1000, dtype=torch.float32) if rank == 0 else torch.zeros(1000)
        P.broadcast_arena(arena, src=0)
        ok_w = bool(torch.equal(arena, torch.arange(1000, dtype=torch.float32)))
        lo, hi = P.shard_bounds(total, world, rank)
        # "forward": flow of global pair i is filled with i (+ channel offset) so the order is checkable
        local = torch.stack([torch.full((2, 3, 4), float(i)) + torch.tensor([0.0, 0.5]).view(2, 1, 1) for i in range(lo, hi)])
        allf = P.gather_flows(local, total)
        ok_g = allf.shape == (total, 2, 3, 4) and all(float(allf[i, 0, 0, 0]) == i and float(allf[i, 1, 0, 0]) == i + 0.5 for i in range(total))
        if total % world == 0:                      # equal shards: root-only gather (what bench.py uses at N > 1)
            root = P.gather_flows_to_root(local)
            ok_g = ok_g and ((root is None) if rank != 0 else bool(torch.equal(root, allf)))
        # data-parallel training: one all-reduce of the gradient arena, averaged
        g = torch.full((257,), float(rank + 1))
        P.allreduce_gradients(g)
        ok_g = ok_g and bool(torch.allclose(g, torch.full((257,), sum(range(1, world + 1)) / world)))
        q.put((rank, ok_w, bool(ok_g)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [4, 5])
def test_broadcast_and_gather_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True), (1, True, True)]
