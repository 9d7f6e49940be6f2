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


class _TwoHead(torch.nn.Module):
    """fc1 -> (head_a | head_b): a loss through head_a alone leaves head_b without a gradient (like Q.prior_emb in
    calculate_loss with an image batch)."""

    def __init__(self):
        super().__init__()
        self.fc1, self.head_a, self.head_b = torch.nn.Linear(16, 32), torch.nn.Linear(32, 3), torch.nn.Linear(32, 5)
        with torch.no_grad():
            for i, p in enumerate(self.parameters()):
                p.copy_(torch.linspace(-0.5, 0.5 + 0.1 * i, p.numel()).view_as(p))

    def forward(self, z):
        return self.head_a(torch.tanh(self.fc1(z)))


def _reducer_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        z = torch.randn(12, 16)
        ref, net = _TwoHead(), _TwoHead()
        red = parallel.FlatGradReducer(net.parameters(), bucket_bytes=256)   # several small buckets
        assert len(red.buckets) > 2 and red.world == world
        for step in range(3):
            ref.zero_grad()
            ref(z).pow(2).sum(1).mean().backward()
            if step == 1:   # an outside zero_grad(set_to_none=True) detaches the views: the hooks must fold the fresh grads back
                for p in net.parameters():
                    p.grad = None
                red._pending = [c for _, _, c in red.buckets]
                red._launched = [False] * len(red.buckets)
                red.fired = [False] * len(red.params)
                red.flat_grad.zero_()
            else:
                red.zero_grad()
            zl, _ = parallel.shard(z, rank, world)
            net(zl).pow(2).sum(1).mean().backward()
            launched_in_backward = sum(red._launched)
            red.finish()
            assert launched_in_backward >= 1          # complete buckets went out from the hooks, during backward
            assert all(red._launched)
            for (name, a), b in zip(net.named_parameters(), ref.parameters()):
                fired = red.fired[[id(q) for q in red.params].index(id(a))]
                if name.startswith("head_b"):
                    assert not fired and (a.grad is None or float(a.grad.abs().max()) == 0.0)
                    continue
                assert fired and a.grad.data_ptr() >= red.flat_grad.data_ptr()   # still a view of the flat buffer
                assert torch.allclose(a.grad * red.grad_scale(), b.grad, atol=1e-6), (name, (a.grad * red.grad_scale() - b.grad).abs().max())
    finally:
        dist.destroy_process_group()


def test_flat_grad_reducer_overlapped_buckets_two_ranks():
    mp.spawn(_reducer_worker, args=(2, _free_port()), nprocs=2, join=True)
