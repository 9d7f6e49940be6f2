"""Multi-GPU plumbing for the sampling hot path: one process per GPU, chains sharded by contiguous ranges.

Every term of the Langevin target is a sum over the batch (reference MCMC.py:32-34, 56-60), so chains are independent:
sampling needs NO collective.  A rank that passes ``chain0 = first global chain index of its shard`` to the samplers
draws the same Philox noise the single-GPU run would (tests/test_gpu_parity.py::test_philox_noise_is_shard_invariant...).
The only exchange step is in the training iteration that CALLS the samplers: averaging parameter gradients after each
backward (reference train_gen_recon.py:217,228,238) -- ``allreduce_mean_grads`` below, NCCL on GPUs, gloo in CPU tests.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced partition of n chains: returns (start, count) for ``rank``; the first n % world ranks get
    one extra chain."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def shard(t, rank, world, dim=0):
    start, count = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, start, count), start


def gather_chains(local, n_total, group=None):
    """Reassemble per-rank shards (possibly ragged) in global chain order on every rank.  Outside the timed path."""
    world = dist.get_world_size(group)
    counts = [shard_range(n_total, r, world)[1] for r in range(world)]
    pad = max(counts)
    buf = local.new_zeros((pad,) + tuple(local.shape[1:]))
    buf[: local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], 0)


def allreduce_mean_grads(params, group=None, bucket_bytes=64 << 20):
    """Average .grad over ranks in flat buckets (one collective per ~64 MB: NVSwitch collectives are latency- not
    link-bound).  Call after backward() and BEFORE clip_grad_norm_ (the reference clips the batch-mean gradient)."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        bucket.append(g)
        size += g.numel() * g.element_size()
        if size >= bucket_bytes:
            flush()
    flush()
