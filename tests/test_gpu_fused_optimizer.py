"""The optimiser half of the training step around the samplers: FlatAdam (damc_fused_clip_adam over
flat buffers, gradients all-reduced bucket by bucket from backward hooks) against torch's clip_grad_norm_ + Adam / AdamW,
the sequence of reference train_gen_recon.py:216-219 / :227-230 / :237-240 with the optimisers of :152-154."""
import copy
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc1, self.fc2 = torch.nn.Linear(24, 64), torch.nn.Linear(64, 8)
        self.unused = torch.nn.Linear(64, 4)   # never part of the loss: torch leaves it untouched, so must FlatAdam

    def forward(self, z):
        return self.fc2(torch.nn.functional.leaky_relu(self.fc1(z), 0.2))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


@pytest.mark.parametrize("decoupled,wd,max_norm", [(False, 0.0, 100.0), (True, 1e-2, 100.0), (False, 0.0, 0.05), (True, 1e-4, 0.05),
                                                   (False, 1e-3, None)])
def test_flat_adam_matches_torch_clip_and_adam(decoupled, wd, max_norm, dev):
    from damc_b200 import parallel
    torch.manual_seed(3)
    ref = _Net().to(dev)
    net = copy.deepcopy(ref)
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    ropt = cls(ref.parameters(), lr=2e-3, betas=(0.5, 0.999), weight_decay=wd)
    fopt = parallel.FlatAdam(net.parameters(), lr=2e-3, betas=(0.5, 0.999), weight_decay=wd, decoupled=decoupled,
                             max_norm=max_norm, bucket_bytes=4096)
    unused0 = net.unused.weight.detach().clone()
    for step in range(6):
        z = torch.randn(32, 24, device=dev)
        ropt.zero_grad()
        ref(z).pow(2).sum(1).mean().backward()
        rnorm = torch.nn.utils.clip_grad_norm_(ref.parameters(), max_norm) if max_norm else None
        ropt.step()
        fopt.zero_grad()
        net(z).pow(2).sum(1).mean().backward()
        fnorm = fopt.step()
        if rnorm is not None:
            assert abs(float(fnorm) - float(rnorm)) <= 1e-5 * float(rnorm), (step, float(fnorm), float(rnorm))
        for (name, a), b in zip(net.named_parameters(), ref.parameters()):
            err = float((a - b).abs().max() / (b.abs().max() + 1e-12))
            assert err < 2e-6, (step, name, err)
    assert torch.equal(net.unused.weight, unused0)                       # no gradient -> untouched (no decay, no moment update)
    assert net.fc1.weight.data_ptr() >= fopt.flat_param.data_ptr()       # parameters are views of the flat buffer
    # a changed learning rate (the reference decays lr by hand, train_gen_recon.py:247-256) is picked up on the next step
    fopt.lr = 0.0
    before = net.fc1.weight.detach().clone()
    fopt.zero_grad()
    net(torch.randn(4, 24, device=dev)).sum().backward()
    fopt.step()
    if not (decoupled and wd):
        assert torch.equal(net.fc1.weight, before)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _nccl_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from damc_b200 import parallel
        dev = torch.device("cuda", rank)
        torch.manual_seed(5)
        ref = _Net().to(dev)
        net = copy.deepcopy(ref)
        ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3, betas=(0.5, 0.999), weight_decay=1e-4)
        fopt = parallel.FlatAdam(net.parameters(), lr=1e-3, betas=(0.5, 0.999), weight_decay=1e-4, decoupled=True, max_norm=0.5,
                                 bucket_bytes=1024)
        assert fopt.reducer.world == world and len(fopt.reducer.buckets) > 1
        worst = 0.0
        for step in range(4):
            z = torch.randn(16 * world, 24, device=dev)      # same seed on every rank: the full batch
            ropt.zero_grad()
            ref(z).pow(2).sum(1).mean().backward()           # single-process gradient of the full batch
            torch.nn.utils.clip_grad_norm_(ref.parameters(), 0.5)
            ropt.step()
            fopt.zero_grad()
            zl, _ = parallel.shard(z, rank, world)
            net(zl).pow(2).sum(1).mean().backward()          # this rank's shard; buckets all-reduced from the hooks
            fopt.step()
            for a, b in zip(net.parameters(), ref.parameters()):
                worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-12)))
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with gpurun --gpus 2)")
def test_flat_adam_two_gpu_nccl_equals_single_process(dev):
    """Sharded batch + NCCL bucketed all-reduce from backward hooks + fused clip/AdamW == one process on the full batch."""
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_nccl_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert len(ret) == 2 and max(ret.values()) < 5e-6, dict(ret)
