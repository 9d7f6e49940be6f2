"""One training iteration around the CUDA samplers -- the caller on either side of the hot path.

Mirrors reference workspace/train_gen_recon.py:187-261: amortized init z0 = Q_dummy(x) -> posterior Langevin ->
prior Langevin on [z0, randn] -> 6 denoiser updates, 1 generator update, 1 EBM update.  Differences, all deliberate:
  * the reference's extra ``zp = Q(x=None, ...)`` (:198) is dropped: its result is never used (100 wasted reverse steps);
  * with torch.distributed initialised, parameter gradients are averaged over ranks after every backward and BEFORE
    clip_grad_norm_ (the losses are batch means, so equal shards + averaging reproduce the single-process gradient);
  * sampling itself needs no collective: every rank samples its own shard of the batch.
The losses and optimisers are ordinary PyTorch; only the three sampler calls run in libdamc_b200.
"""
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import MCMC, parallel


@dataclass
class TrainConfig:  # defaults = train_gen_recon.py:383-402
    g_l_steps: int = 30
    g_l_step_size: float = 0.1
    g_l_with_noise: bool = True
    g_llhd_sigma: float = 0.1
    e_l_steps: int = 60
    e_l_step_size: float = 0.4
    e_l_with_noise: bool = True
    p_mask: float = 0.2
    q_updates: int = 6
    max_norm: float = 100.0
    precision: str = "tf32"   # what the reference's cuDNN convolutions compute in by default on a GPU
    q_loss_engine: str = "torch"   # "library" | "library_graphed": the Linear layers of Q.p in Q.calculate_loss on the tcgen05 TF32 GEMMs


def _step(loss, params, opt, cfg, group):
    if isinstance(opt, parallel.FlatAdam):
        # flat gradient buffer: buckets are all-reduced from backward hooks while backward runs; clip + Adam(W) in one fused pass
        opt.zero_grad()
        loss.backward()
        opt.step()
        return
    opt.zero_grad()
    loss.backward()
    if dist.is_available() and dist.is_initialized():
        parallel.allreduce_mean_grads(params, group)
    torch.nn.utils.clip_grad_norm_(params, max_norm=cfg.max_norm)
    opt.step()


def make_fused_optimizers(G, E, Q, cfg, g_lr=2e-4, e_lr=1e-4, q_lr=2e-4, group=None, bucket_bytes=32 << 20):
    """The reference's three optimisers (train_gen_recon.py:152-154: Adam / Adam / AdamW(weight_decay=1e-4), betas (0.5, 0.999))
    as FlatAdam instances with the clip threshold folded in.  Returns (G_opt, E_opt, Q_opt) in training_iteration's order."""
    mk = lambda net, lr, wd, dec: parallel.FlatAdam(net.parameters(), lr=lr, betas=(0.5, 0.999), weight_decay=wd, decoupled=dec,
                                                    max_norm=cfg.max_norm, group=group, bucket_bytes=bucket_bytes)
    return mk(G, g_lr, 0.0, False), mk(E, e_lr, 0.0, False), mk(Q, q_lr, 1e-4, True)


def training_iteration(x, G, E, Q, Q_dummy, G_opt, E_opt, Q_opt, cfg=TrainConfig(), group=None, chain0=None):
    """x: this rank's image shard [B,nc,H,W] on its GPU.  Returns dict of detached scalar losses.
    chain0: global index of this shard's first chain (Philox key).  Default: rank * B under torch.distributed -- the
    reference seeds every process alike (torch.manual_seed(args.seed), train_gen_recon.py:353), so without the offset all
    ranks would draw the same seed AND the same chain indices, i.e. identical Langevin noise on every shard."""
    B = x.size(0)
    if chain0 is None:
        chain0 = dist.get_rank(group) * B if (dist.is_available() and dist.is_initialized()) else 0
    z_mask = (torch.rand(B, device=x.device) >= cfg.p_mask).float().unsqueeze(-1)
    Q.eval(); G.eval(); E.eval()
    with torch.no_grad():
        z0 = Q_dummy(x)
    zk_pos = z0.detach().clone().requires_grad_(True)
    zk_pos = MCMC.sample_langevin_post_z_with_prior(zk_pos, x, G, E, cfg.g_l_steps, cfg.g_llhd_sigma, cfg.g_l_with_noise,
                                                    cfg.g_l_step_size, chain0=chain0, precision=cfg.precision)
    z_neg0 = torch.cat([z0.detach().clone(), torch.randn_like(z0)], dim=0).requires_grad_(True)
    zk_neg = MCMC.sample_langevin_prior_z(z_neg0, E, cfg.e_l_steps, cfg.e_l_step_size, cfg.e_l_with_noise,
                                          chain0=2 * chain0)
    q_params, g_params, e_params = list(Q.parameters()), list(G.parameters()), list(E.parameters())
    Q.train()
    for _ in range(cfg.q_updates):
        q_loss = Q.calculate_loss(x=x, z=zk_pos, mask=z_mask, engine=cfg.q_loss_engine).mean()
        _step(q_loss, q_params, Q_opt, cfg, group)
    G.train()
    g_loss = torch.sum((G(zk_pos) - x) ** 2, dim=[1, 2, 3]).mean()
    _step(g_loss, g_params, G_opt, cfg, group)
    E.train()
    e_loss = E(zk_pos).mean() - E(zk_neg).mean()
    _step(e_loss, e_params, E_opt, cfg, group)
    Q.eval(); G.eval(); E.eval()
    return {"q_loss": q_loss.detach(), "g_loss": g_loss.detach(), "e_loss": e_loss.detach(), "zk_pos": zk_pos,
            "zk_neg": zk_neg}


@torch.no_grad()
def ema_update(Q, Q_dummy, rho=0.005):
    """Q_dummy <- rho Q + (1-rho) Q_dummy (train_gen_recon.py:258-261).  In-place .data updates: the packed-weight
    cache re-reads values on every sampler call, so the new weights are picked up."""
    for p, t in zip(Q.parameters(), Q_dummy.parameters()):
        t.data.copy_(rho * p.data + (1 - rho) * t.data)
