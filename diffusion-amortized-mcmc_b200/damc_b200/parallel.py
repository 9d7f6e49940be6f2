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


# ----------------------------------------------------------------------------------------------------------------------
# gradient exchange overlapped with backward + fused clip / Adam over flat buffers (the training step around the samplers)
# ----------------------------------------------------------------------------------------------------------------------
class FlatGradReducer:
    """All gradients of a parameter list live in ONE persistent flat buffer (``p.grad`` are views of it), and are summed
    over ranks bucket by bucket WHILE backward is still running: a post-accumulate hook on every parameter counts its
    bucket down and launches ``all_reduce(bucket slice, async_op=True)`` when the bucket is complete -- in place, no
    ``torch.cat`` and no copy back.  Buckets are contiguous slices in reverse parameter order (the order autograd produces
    gradients in).  ``finish()`` launches what the hooks did not (parameters without a gradient this step) and waits.

    The sum is NOT divided here: the consumer folds 1 / world_size into its own pass (FlatAdam: ``grad_scale``).
    Replaces the per-backward gradient exchange a data-parallel run of reference train_gen_recon.py:217,228,238 needs."""

    def __init__(self, params, group=None, bucket_bytes=32 << 20):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise RuntimeError("all parameters must share one device and dtype")
        self.group = group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        # flat layout in REVERSE parameter order: the first gradients autograd finishes sit at the front, so a bucket is a prefix slice
        order = list(reversed(range(len(self.params))))
        self.offset = [0] * len(self.params)
        off = 0
        for i in order:
            self.offset[i] = off
            off += self.params[i].numel()
        self.numel = off
        self.flat_grad = torch.zeros(off, dtype=dt, device=dev)
        self._attach()
        # buckets: consecutive parameters (in flat order) up to bucket_bytes
        self.bucket_of = [0] * len(self.params)
        self.buckets = []   # (begin, end, n_params)
        b0, nb, cnt = 0, 0, 0
        for i in order:
            self.bucket_of[i] = len(self.buckets)
            nb += self.params[i].numel() * self.params[i].element_size()
            cnt += 1
            if nb >= bucket_bytes:
                end = self.offset[i] + self.params[i].numel()
                self.buckets.append((b0, end, cnt))
                b0, nb, cnt = end, 0, 0
        if cnt:
            self.buckets.append((b0, off, cnt))
        self._pending = [c for _, _, c in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self.fired = [False] * len(self.params)
        for i, p in enumerate(self.params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _attach(self):
        self._views = [self.flat_grad[self.offset[i]:self.offset[i] + p.numel()].view_as(p) for i, p in enumerate(self.params)]
        for p, v in zip(self.params, self._views):
            p.grad = v

    def _make_hook(self, i):
        def hook(p):
            g = p.grad
            if g is not self._views[i]:
                # someone replaced .grad (zero_grad(set_to_none=True), a fresh tensor from autograd): fold it back into the flat buffer
                if g is not None:
                    self._views[i].copy_(g)
                p.grad = self._views[i]
            self.fired[i] = True
            b = self.bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch(b)
        return hook

    def _launch(self, b):
        if self._launched[b]:
            return
        self._launched[b] = True
        if self.world > 1:
            beg, end, _ = self.buckets[b]
            self._handles.append(dist.all_reduce(self.flat_grad[beg:end], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def zero_grad(self):
        self.flat_grad.zero_()   # the .grad views stay attached (a detached one is folded back by its hook)
        self._pending = [c for _, _, c in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        self.fired = [False] * len(self.params)

    def finish(self):
        """Launch the buckets backward did not complete (zeros for parameters without a gradient) and wait for all of them.
        Every rank launches every bucket exactly once per step, in hook order for the complete ones: ranks run the same
        graph, so the collective order matches across ranks."""
        for b in range(len(self.buckets)):
            self._launch(b)
        for h in self._handles:
            h.wait()
        self._handles = []

    def grad_scale(self):
        return 1.0 / self.world


class FlatAdam:
    """Adam / AdamW + clip_grad_norm_ on flat buffers through ``damc_fused_clip_adam`` (two reduction launches + one update
    launch per step), fed by a FlatGradReducer.  Matches torch.optim.Adam / AdamW (no amsgrad) preceded by
    ``clip_grad_norm_(params, max_norm)`` on the rank-averaged gradient -- the sequence of reference
    train_gen_recon.py:216-219 -- including torch's rule that a parameter without a gradient is left untouched.
    ``step()`` returns the total gradient norm as a 0-d device tensor (what clip_grad_norm_ returns)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, max_norm=None,
                 group=None, bucket_bytes=32 << 20):
        self.reducer = FlatGradReducer(params, group, bucket_bytes)
        r = self.reducer
        if r.flat_grad.dtype != torch.float32 or not r.flat_grad.is_cuda:
            raise RuntimeError("FlatAdam needs CUDA float32 parameters (damc_b200 has no CPU fallback)")
        self.lr, self.betas, self.eps, self.weight_decay, self.decoupled = lr, betas, eps, weight_decay, decoupled
        self.max_norm = max_norm
        self.flat_param = torch.empty_like(r.flat_grad)
        for i, p in enumerate(r.params):   # parameters become views of the flat buffer (values preserved)
            view = self.flat_param[r.offset[i]:r.offset[i] + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros_like(r.flat_grad)
        self.exp_avg_sq = torch.zeros_like(r.flat_grad)
        self.steps = [0] * len(r.params)
        self.scratch = torch.zeros(1032, dtype=torch.float32, device=r.flat_grad.device)

    def zero_grad(self, set_to_none=False):
        self.reducer.zero_grad()

    def step(self):
        import ctypes as C
        from ._lib import lib, check
        r = self.reducer
        r.finish()
        # element ranges of the parameters that received a gradient, merged while adjacent in the flat buffer and at the same step
        act = sorted((r.offset[i], r.params[i].numel(), i) for i in range(len(r.params)) if r.fired[i])
        begins, counts, steps = [], [], []
        for off, n, i in act:
            self.steps[i] += 1
            if begins and begins[-1] + counts[-1] == off and steps[-1] == self.steps[i]:
                counts[-1] += n
            else:
                begins.append(off); counts.append(n); steps.append(self.steps[i])
        nr = len(begins)
        U64, I32 = C.c_ulonglong * max(nr, 1), C.c_int * max(nr, 1)
        dev = r.flat_grad.device
        with torch.cuda.device(dev):
            check(lib().damc_fused_clip_adam(
                C.c_void_p(self.flat_param.data_ptr()), C.c_void_p(r.flat_grad.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
                C.c_void_p(self.exp_avg_sq.data_ptr()), r.numel, nr, U64(*begins), U64(*counts), I32(*steps), float(self.lr),
                float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay), int(bool(self.decoupled)),
                float(self.max_norm) if self.max_norm else 0.0, r.grad_scale(), C.c_void_p(self.scratch.data_ptr()),
                C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "damc_fused_clip_adam")
        return self.scratch[1024]
