"""CPU tests (gloo, world_size 2) of the multi-process host logic: chain sharding, ragged gather, gradient averaging."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from damc_b200 import parallel


def test_shard_range_partitions():
    for n in (0, 1, 7, 128, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        z = torch.randn(n, 16)
        x = torch.randn(n, 4)
        # (1) sharded "sampling" with a stand-in per-chain map: shards + gather == unsharded, in global order
        f = lambda zz, xx, c0: zz * 2.0 + xx.sum(1, keepdim=True) + torch.arange(c0, c0 + len(zz)).float()[:, None]
        zl, start = parallel.shard(z, rank, world)
        xl, _ = parallel.shard(x, rank, world)
        got = parallel.gather_chains(f(zl, xl, start), n)
        assert torch.equal(got, f(z, x, 0))
        # (2) gradient averaging of a batch-mean loss over equal shards == single-process gradient
        lin = torch.nn.Linear(16, 3)
        with torch.no_grad():
            for p in lin.parameters():
                p.copy_(torch.linspace(-1, 1, p.numel()).view_as(p))
        ref = torch.nn.Linear(16, 3)
        ref.load_state_dict(lin.state_dict())
        m = (n // world) * world
        ref(z[:m]).pow(2).sum(1).mean().backward()
        zl2, _ = parallel.shard(z[:m], rank, world)
        lin(zl2).pow(2).sum(1).mean().backward()
        parallel.allreduce_mean_grads(list(lin.parameters()), bucket_bytes=64)  # tiny buckets: exercise the flush path
        for a, b in zip(lin.parameters(), ref.parameters()):
            assert torch.allclose(a.grad, b.grad, atol=1e-6), (a.grad - b.grad).abs().max()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 7])
def test_two_rank_sharding_and_grad_average(n):
    mp.spawn(_worker, args=(2, _free_port(), n), nprocs=2, join=True)
