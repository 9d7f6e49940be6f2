"""Host-side mirror of the reference's network interfaces for the sampling hot path.

These classes keep the attribute layout and state-dict keys the samplers (and the
reference's checkpoints) rely on -- ``netG.gen`` / ``netG.nz``, ``netE.ebm``,
``Q.p`` / ``Q.encoder`` / ``Q.prior_emb`` / ``Q.n_interval`` ... -- so a user of
the reference can construct them with the same arguments and load the same
``*_state_dict`` entries (reference: workspace/src/diffusion_net.py:20-203 for the
generators, :207-223 for the EBM, :227-413 encoders, :417-533 denoiser, :537-622
amortizer).  They are built from small shape tables instead of one class body
per dataset.

Only ``_netQ_U.forward`` (the DAMC ancestral sampler, reference :585-622) is
routed to the CUDA library; module ``forward`` of G / E / p stay ordinary
PyTorch because the *training* losses that call them are outside the hot path
(``calculate_loss`` is out of scope).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# (channel multiplier of ngf | "nc", kernel, stride, padding) per ConvTranspose2d.
# Reference: diffusion_net.py:26-45 (cifar10), :59-78 (svhn), :92-116 (celeba64),
# :130-164 (celebaHQ), :178-197 (mnist).
GENERATOR_TABLE = {
    "cifar10": dict(nz=128, ngf=128, nc=3, layers=[(8, 8, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 3, 1, 1)]),
    "svhn": dict(nz=100, ngf=64, nc=3, layers=[(8, 4, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 4, 2, 1)]),
    "celeba64": dict(nz=100, ngf=128, nc=3,
                     layers=[(8, 4, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), (1, 4, 2, 1), ("nc", 4, 2, 1)]),
    "celebaHQ": dict(nz=128, ngf=128, nc=3,
                     layers=[(16, 4, 1, 0), (8, 4, 2, 1), (4, 4, 2, 1), (4, 4, 2, 1), (2, 4, 2, 1), (1, 4, 2, 1),
                             ("nc", 4, 2, 1)]),
    "mnist": dict(nz=100, ngf=128, nc=1, layers=[(8, 7, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 3, 1, 1)]),
}


class _netG(nn.Module):
    """Transposed-conv decoder z[B,nz] -> x[B,nc,H,W]; ``gen`` is an nn.Sequential with
    ConvTranspose2d at even indices, LeakyReLU(0.2) between and Tanh last."""

    def __init__(self, dataset, nz=None, ngf=None, nc=None):
        super().__init__()
        spec = GENERATOR_TABLE[dataset]
        self.dataset = dataset
        self.nz = spec["nz"] if nz is None else nz
        ngf = spec["ngf"] if ngf is None else ngf
        nc = spec["nc"] if nc is None else nc
        mods, cin = [], self.nz
        for i, (mult, k, s, p) in enumerate(spec["layers"]):
            cout = nc if mult == "nc" else ngf * mult
            mods.append(nn.ConvTranspose2d(cin, cout, k, s, p, bias=True))
            mods.append(nn.LeakyReLU(0.2) if i + 1 < len(spec["layers"]) else nn.Tanh())
            cin = cout
        self.gen = nn.Sequential(*mods)

    def forward(self, z):
        return self.gen(z.reshape((len(z), self.nz, 1, 1)))


def _netG_cifar10(nz=128, ngf=128, nc=3):
    return _netG("cifar10", nz, ngf, nc)


def _netG_svhn(nz=100, ngf=64, nc=3):
    return _netG("svhn", nz, ngf, nc)


def _netG_celeba64(nz=100, ngf=128, nc=3):
    return _netG("celeba64", nz, ngf, nc)


def _netG_celebaHQ(nz=128, ngf=128, nc=3):
    return _netG("celebaHQ", nz, ngf, nc)


def _netG_mnist(nz=100, ngf=128, nc=1):
    return _netG("mnist", nz, ngf, nc)


class _netE(nn.Module):
    """Latent EBM prior: Linear(nz,ndf)-LReLU(.2)-Linear(ndf,ndf)-LReLU(.2)-Linear(ndf,1), squeezed
    (reference diffusion_net.py:207-223)."""

    def __init__(self, nz=128, ndf=200, nez=1):
        super().__init__()
        self.ebm = nn.Sequential(nn.Linear(nz, ndf), nn.LeakyReLU(0.2), nn.Linear(ndf, ndf), nn.LeakyReLU(0.2),
                                 nn.Linear(ndf, nez))

    def forward(self, z):
        return self.ebm(z).squeeze()


# Conv2d stacks of the amortizer's image encoder: (out multiple of nif | "emb", k, s, p).
# Reference: diffusion_net.py:233-263 / :274-310 / :321-369 / :380-409.
ENCODER_TABLE = {
    "cifar10": [(1, 3, 1, 1), (2, 4, 2, 1), (4, 4, 2, 1), (8, 4, 2, 1), ("emb", 4, 1, 0)],
    "mnist": [(1, 3, 1, 1), (2, 4, 2, 1), (4, 4, 2, 1), (8, 4, 2, 1), ("emb", 3, 1, 0)],
    "celeba64": [(1, 3, 1, 1), (2, 4, 2, 1), (4, 4, 2, 1), (8, 4, 2, 1), (8, 4, 2, 1), ("emb", 4, 1, 0)],
    "celebaHQ": [(1, 3, 1, 1), (2, 4, 2, 1), (4, 4, 2, 1), (4, 4, 2, 1), (8, 4, 2, 1), (8, 4, 2, 1), (8, 4, 2, 1),
                 ("emb", 4, 1, 0)],
}


class Encoder(nn.Module):
    """Conv2d + InstanceNorm2d(affine) + LeakyReLU(0.2) stack -> [B, nemb].  Runs once per DAMC call;
    stays PyTorch."""

    def __init__(self, dataset, nc=3, nemb=128, nif=64):
        super().__init__()
        self.nemb = nemb
        mods, cin = [], nc
        table = ENCODER_TABLE[dataset]
        for i, (mult, k, s, p) in enumerate(table):
            cout = nemb if mult == "emb" else nif * mult
            mods.append(nn.Conv2d(cin, cout, k, s, p, bias=True))
            if i + 1 < len(table):
                mods.append(nn.InstanceNorm2d(cout, affine=True))
                mods.append(nn.LeakyReLU(0.2, inplace=True))
            cin = cout
        self.net = nn.Sequential(*mods)

    def forward(self, x):
        return self.net(x).reshape((len(x), self.nemb))


class ConcatSquashLinearSkipCtx(nn.Module):
    """out = Linear(x) * sigmoid(Wg c + bg) + Wb c + Skip(x),  c = SiLU(Linear(SiLU(ctx)))
    (reference diffusion_net.py:417-445)."""

    def __init__(self, dim_in, dim_out, nxemb, ntemb):
        super().__init__()
        self._layer = nn.Sequential(nn.Linear(dim_in, dim_out))
        self._layer_ctx = nn.Sequential(nn.SiLU(), nn.Linear(ntemb + nxemb, dim_out), nn.SiLU())
        self._hyper_bias = nn.Linear(dim_out, dim_out, bias=False)
        self._hyper_gate = nn.Linear(dim_out, dim_out)
        self._skip = nn.Linear(dim_in, dim_out)

    def forward(self, ctx, x):
        c = self._layer_ctx(ctx)
        return self._layer(x) * torch.sigmoid(self._hyper_gate(c)) + self._hyper_bias(c) + self._skip(x)


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim, max_time=1000.0):
        super().__init__()
        self.dim, self.max_time = dim, max_time

    def forward(self, x):
        x = x * (1000.0 / self.max_time)  # the reference scales in place (:454); value-identical
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
        arg = x[:, None] * freq[None, :]
        return torch.cat((arg.sin(), arg.cos()), dim=-1)


class Diffusion_UnetA(nn.Module):
    """epsilon-network of the latent diffusion amortizer (reference diffusion_net.py:463-533)."""

    def __init__(self, nz=128, nxemb=128, ntemb=128, residual=False, nf=4):
        super().__init__()
        self.nz, self.nxemb, self.ntemb, self.residual, self.nf = nz, nxemb, ntemb, residual, nf
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(ntemb, max_time=1.0), nn.Linear(ntemb, ntemb), nn.SiLU(),
                                      nn.Linear(ntemb, ntemb))
        self.B = nn.Parameter(torch.randn(nz, nz // 2), requires_grad=True)
        L = ConcatSquashLinearSkipCtx
        self.in_layers = nn.ModuleList([L(nz * 2, 32 * nf, nxemb, ntemb), L(32 * nf, 64 * nf, nxemb, ntemb),
                                        L(64 * nf, 64 * nf, nxemb, ntemb)])
        self.mid_layers = nn.ModuleList([L(64 * nf, 64 * nf, nxemb, ntemb)])
        self.out_layers = nn.ModuleList([L(128 * nf, 64 * nf, nxemb, ntemb), L(128 * nf, 32 * nf, nxemb, ntemb),
                                         L(64 * nf, nz, nxemb, ntemb)])

    def input_emb(self, x):
        proj = 2 * np.pi * torch.matmul(x, self.B)
        return torch.cat([torch.sin(proj), torch.cos(proj), x], dim=1)

    def forward(self, z, logsnr, xemb):
        u = torch.arctan(torch.exp(-0.5 * torch.clamp(logsnr, min=-20.0, max=20.0))) / (0.5 * np.pi)
        ctx = torch.cat([self.time_mlp(u), xemb], dim=1)
        skips, out = [], self.input_emb(z)
        for layer in self.in_layers:
            out = layer(ctx=ctx, x=out)
            skips.append(out)
            out = F.leaky_relu(out, negative_slope=0.01)
        out = self.mid_layers[0](ctx=ctx, x=out)
        for layer in self.out_layers:
            out = F.leaky_relu(torch.cat([out, skips.pop()], dim=1), negative_slope=0.01)
            out = layer(ctx=ctx, x=out)
        return z + out if self.residual else out


class _netQ_U(nn.Module):
    """Diffusion-based amortizer.  ``forward`` is the DAMC ancestral sampler (reference
    diffusion_net.py:585-622) and runs on the CUDA library (``damc_denoise``)."""

    def __init__(self, nc=3, nz=128, nxemb=128, ntemb=128, nf=4, nif=64, diffusion_residual=False, n_interval=20,
                 logsnr_min=-20.0, logsnr_max=20.0, var_type="small", with_noise=False, cond_w=0, net_arch="A",
                 dataset="cifar10"):
        super().__init__()
        self.n_interval, self.logsnr_min, self.logsnr_max = n_interval, logsnr_min, logsnr_max
        self.var_type, self.nz, self.nxemb, self.with_noise, self.cond_w = var_type, nz, nxemb, with_noise, cond_w
        enc = {"cifar10": "cifar10", "svhn": "cifar10", "mnist": "mnist", "celeba64": "celeba64"}.get(dataset,
                                                                                                      "celebaHQ")
        self.encoder = Encoder(enc, nc=1 if dataset == "mnist" else nc, nemb=nxemb, nif=nif)
        self.p = Diffusion_UnetA(nz=nz, nxemb=nxemb, ntemb=ntemb, residual=diffusion_residual, nf=nf)
        self.xemb = nn.Parameter(torch.randn(1, nxemb), requires_grad=True)
        self.prior_emb = nn.Sequential(nn.Linear(nz, 128), nn.LeakyReLU(), nn.Linear(128, nxemb))

    def forward(self, x=None, b=None, device=None, cond_w=-1, noise=None, precision=None):
        from . import MCMC  # late import: MCMC needs the CUDA library
        return MCMC.damc_sample(self, x=x, b=b, device=device, cond_w=cond_w, noise=noise, precision=precision)

    def calculate_loss(self, x=None, z=None, mask=None, engine="torch"):
        """Denoising loss 0.5*|eps - eps_hat|^2 per sample at a random noise level (reference diffusion_net.py:624-646).
        engine="torch": ordinary PyTorch autograd, as the reference.  engine="library": the 35 Linear layers of the seven
        ConcatSquashLinearSkipCtx blocks of Q.p run forward AND backward as tcgen05 TF32 GEMMs of libdamc_b200
        (damc_b200/denoiser_train.py); encoder, prior_emb, time_mlp and the loss stay autograd.  Same random draws either way.
        engine="library_graphed": the same, with the core's forward and backward replayed from CUDA graphs."""
        assert z is not None
        n = len(z)
        if x is not None:
            xemb = self.encoder(x)
            if mask is not None:
                xemb = xemb * mask + self.prior_emb(torch.randn(len(x), self.nz, device=x.device)) * (1 - mask)
        else:
            assert mask is None
            xemb = self.prior_emb(torch.randn(n, self.nz, device=z.device))
        u = torch.rand(n).to(z.device)
        logsnr = logsnr_schedule_fn(u, logsnr_min=self.logsnr_min, logsnr_max=self.logsnr_max)
        lam = logsnr.reshape(n, 1)
        eps = torch.randn_like(z)
        zt = z * torch.sqrt(torch.sigmoid(lam)) + torch.sqrt(torch.sigmoid(-lam)) * eps
        if engine in ("library", "library_graphed"):
            from . import denoiser_train
            eps_pred = denoiser_train.eps_network(self.p, zt, logsnr, xemb, graphed=engine == "library_graphed")
        elif engine == "torch":
            eps_pred = self.p(z=zt, logsnr=logsnr, xemb=xemb)
        else:
            raise ValueError("engine must be 'torch', 'library' or 'library_graphed'")
        return 0.5 * torch.sum((eps - eps_pred) ** 2, dim=1)


def logsnr_schedule_fn(t, logsnr_min=-20.0, logsnr_max=20.0):
    """lambda(t) = -2 log tan(a t + b), b = atan(exp(-lambda_max/2)), a = atan(exp(-lambda_min/2)) - b
    (reference diffusion_helper_func.py:41-50)."""
    b = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_max)))
    a = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_min))) - b
    return -2.0 * torch.log(torch.tan(a * t + b))
