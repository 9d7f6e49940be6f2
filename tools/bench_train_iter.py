"""One full training iteration around the samplers (reference train_gen_recon.py:196-261 at CIFAR-10 shape, B chains per GPU):
amortised init Q_dummy(x) -> posterior Langevin (K = 30) -> prior Langevin (K = 60) -> 6 denoiser updates, 1 generator update,
1 EBM update -- with the gradient exchange of a data-parallel run.  Launch under torchrun for N > 1.

Reports per-iteration time for (a) torch optimisers + the blocking flatten / all-reduce / copy-back exchange and (b) FlatAdam
(bucketed all-reduce from backward hooks + fused clip/Adam), and for (b) the time with the collectives disabled (world-1
timing of the same kernels) so the exposed share of the all-reduce is visible.  One JSON line from rank 0."""
import copy
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn, parallel, train  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = int(os.environ.get("B", "128"))
    iters, warm = int(os.environ.get("ITERS", "5")), 2
    torch.manual_seed(1)
    G0, E0 = dn._netG_cifar10(128, 128, 3).to(dev), dn._netE(128).to(dev)
    Q0 = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=100, logsnr_min=-5.1,
                    logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev)
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    cfg = train.TrainConfig(precision=os.environ.get("PREC", "bf16"), q_loss_engine=os.environ.get("Q_ENGINE", "torch"))
    MCMC.set_default_denoiser_precision("fp16")
    out = {"world": world, "chains_per_gpu": B, "precision": cfg.precision, "q_loss_engine": cfg.q_loss_engine}

    def run(kind):
        G, E, Q = copy.deepcopy(G0), copy.deepcopy(E0), copy.deepcopy(Q0)
        Qd = copy.deepcopy(Q)
        if kind == "torch":
            opts = (torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999)),
                    torch.optim.Adam(E.parameters(), lr=1e-4, betas=(0.5, 0.999)),
                    torch.optim.AdamW(Q.parameters(), lr=2e-4, weight_decay=1e-4, betas=(0.5, 0.999)))
        else:
            opts = train.make_fused_optimizers(G, E, Q, cfg)
            if kind == "fused_nocomm":
                for o in opts:
                    o.reducer.world = 1
        ts = []
        for i in range(warm + iters):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            train.training_iteration(x, G, E, Q, Qd, *opts, cfg=cfg)
            train.ema_update(Q, Qd)
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor(sorted(ts[warm:])[len(ts[warm:]) // 2], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out["ms_torch_optim_blocking_allreduce"] = run("torch")
    out["ms_flat_adam_overlapped"] = run("fused")
    out["ms_flat_adam_no_collectives"] = run("fused_nocomm") if world > 1 else out["ms_flat_adam_overlapped"]
    out["exposed_allreduce_ms"] = out["ms_flat_adam_overlapped"] - out["ms_flat_adam_no_collectives"]
    out["grad_bytes_per_iter"] = 4 * (6 * sum(p.numel() for p in Q0.parameters()) + sum(p.numel() for p in G0.parameters()) +
                                      sum(p.numel() for p in E0.parameters()))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
