"""CPU ORACLE for the sampling hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
this file; nothing under ``diffusion-amortized-mcmc_b200/`` does (the product path fails loudly without its CUDA
library and has no CPU fallback).

It restates, on the CPU with torch tensors (fp32 or fp64), the algorithm of the reference
``/root/reference/workspace`` for the path in BASELINE.json:north_star.  The reference's arithmetic lives in PyTorch
(ATen), pinned by the reference at PyTorch 1.10.0 (README.md:15) and present here as torch 2.11 -- so the restatement
calls the same ATen operators in the same order, and a second, *analytic* restatement (explicit backward formulas, no
autograd: exactly what the CUDA kernels implement) is checked against it.

PARITY PINNING: the reference ships no tests, golden vectors or fixtures for this path, so the oracle
is pinned against outputs of the reference itself, run in the build container by ``oracle/make_golden.py`` (imports
/root/reference read-only) and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them.

All functions take plain tensors / lists of tensors (no nn.Module), so the oracle shares no code with the product.
"""
import math

import torch
import torch.nn.functional as F

LRELU_G = 0.2  # generator / EBM slope, reference diffusion_net.py:24,211


# ----------------------------------------------------------------------------------------------------------------------
# networks (functional)
# ----------------------------------------------------------------------------------------------------------------------
def gen_forward(gen, z):
    """x_hat = G(z).  ``gen`` = list of (W[Cin,Cout,k,k], b[Cout], stride, pad).
    Reference: diffusion_net.py:26-51 (ConvTranspose2d + LeakyReLU(0.2) ... + Tanh), forward :49-51."""
    h = z.reshape(len(z), -1, 1, 1)
    for i, (W, b, s, p) in enumerate(gen):
        h = F.conv_transpose2d(h, W, b, stride=s, padding=p)
        h = F.leaky_relu(h, LRELU_G) if i + 1 < len(gen) else torch.tanh(h)
    return h


def ebm_forward(ebm, z):
    """E(z) -> [B].  ``ebm`` = [(W1,b1),(W2,b2),(W3,b3)] in nn.Linear layout [out,in].
    Reference: diffusion_net.py:212-223."""
    (W1, b1), (W2, b2), (W3, b3) = ebm
    a1 = F.leaky_relu(F.linear(z, W1, b1), LRELU_G)
    a2 = F.leaky_relu(F.linear(a1, W2, b2), LRELU_G)
    return F.linear(a2, W3, b3).reshape(len(z))


# ----------------------------------------------------------------------------------------------------------------------
# Langevin samplers -- autograd restatement (same op order as the reference)
# ----------------------------------------------------------------------------------------------------------------------
def langevin_prior(z, ebm, steps, step_size, with_noise, noise=None, trace=None):
    """K-step ULA on E(z) + |z|^2/2.  Reference: MCMC.py:27-46.
    noise: optional [K,B,nz] injected standard normals (replaces torch.randn_like, :38).
    trace: optional list; receives (i, en, z_norm) at i%5==0 or last (:40-41)."""
    z = z.detach().clone().requires_grad_(True)
    for i in range(steps):
        en = ebm_forward(ebm, z).sum()
        z_norm = 0.5 * torch.sum(z ** 2)
        g = torch.autograd.grad(en + z_norm, z)[0]
        z.data = z.data - 0.5 * step_size * step_size * g
        if with_noise:
            z.data += step_size * (noise[i] if noise is not None else torch.randn_like(z))
        if trace is not None and (i % 5 == 0 or i == steps - 1):
            trace.append((i, en.item(), z_norm.item()))
    return z.detach()


def langevin_posterior(z, x, gen, ebm, steps, sigma, with_noise, step_size, noise=None, trace=None):
    """K-step ULA on |G(z)-x|^2/(2 sigma^2) + E(z) + |z|^2/2.  Reference: MCMC.py:48-74.
    trace receives (i, en, g_log_lkhd, z_n, mean(grad)) every step (:65-67)."""
    z = z.detach().clone().requires_grad_(True)
    for i in range(steps):
        x_hat = gen_forward(gen, z)
        llhd = 1.0 / (2.0 * sigma * sigma) * torch.sum((x_hat - x) ** 2)
        z_n = 0.5 * torch.sum(z ** 2)
        en = ebm_forward(ebm, z).sum() if ebm is not None else z.new_zeros(())
        g = torch.autograd.grad(llhd + en + z_n, z)[0]
        z.data = z.data - 0.5 * step_size * step_size * g
        if with_noise:
            z.data += step_size * (noise[i] if noise is not None else torch.randn_like(z))
        if trace is not None:
            trace.append((i, float(en), llhd.item(), z_n.item(), g.mean().item()))
    return z.detach()


# ----------------------------------------------------------------------------------------------------------------------
# Langevin samplers -- analytic restatement (what the CUDA kernels compute)
# ----------------------------------------------------------------------------------------------------------------------
def ebm_grad(ebm, z):
    """(E[B], dE/dz[B,nz]) without autograd: gE = W1^T(n1 * W2^T(n2 * w3)), n = 1 | 0.2."""
    (W1, b1), (W2, b2), (W3, b3) = ebm
    h1 = z @ W1.t() + b1
    a1 = torch.where(h1 > 0, h1, LRELU_G * h1)
    h2 = a1 @ W2.t() + b2
    a2 = torch.where(h2 > 0, h2, LRELU_G * h2)
    e = (a2 @ W3.t() + b3).reshape(len(z))
    d2 = torch.where(h2 > 0, 1.0, LRELU_G).to(z.dtype) * W3.reshape(1, -1)
    d1 = torch.where(h1 > 0, 1.0, LRELU_G).to(z.dtype) * (d2 @ W2)
    return e, d1 @ W1


def gen_grad(gen, z, x, sigma):
    """(x_hat, d/dz |G(z)-x|^2/(2 sigma^2)) without autograd.
    The input-gradient of ConvTranspose2d is the ordinary strided conv with the same weight tensor."""
    acts, h = [], z.reshape(len(z), -1, 1, 1)
    for i, (W, b, s, p) in enumerate(gen):
        h = F.conv_transpose2d(h, W, b, stride=s, padding=p)
        if i + 1 < len(gen):
            acts.append(h)
            h = torch.where(h > 0, h, LRELU_G * h)
    x_hat = torch.tanh(h)
    g = (x_hat - x) / (sigma * sigma) * (1.0 - x_hat * x_hat)
    for i in range(len(gen) - 1, -1, -1):
        W, b, s, p = gen[i]
        g = F.conv2d(g, W, None, stride=s, padding=p)
        if i > 0:
            g = g * torch.where(acts[i - 1] > 0, 1.0, LRELU_G).to(z.dtype)
    return x_hat, g.reshape(len(z), -1)


def langevin_posterior_analytic(z, x, gen, ebm, steps, sigma, with_noise, step_size, noise=None):
    z = z.detach().clone()
    for i in range(steps):
        _, gG = gen_grad(gen, z, x, sigma)
        gE = ebm_grad(ebm, z)[1] if ebm is not None else 0.0
        z = z - 0.5 * step_size * step_size * (gG + gE + z)
        if with_noise:
            z = z + step_size * noise[i]
    return z


def langevin_prior_analytic(z, ebm, steps, step_size, with_noise, noise=None):
    z = z.detach().clone()
    for i in range(steps):
        z = z - 0.5 * step_size * step_size * (ebm_grad(ebm, z)[1] + z)
        if with_noise:
            z = z + step_size * noise[i]
    return z


# ----------------------------------------------------------------------------------------------------------------------
# toy example (reference toy_example/toy_example.py:22-47 G, :110-131 sampler closure)
# ----------------------------------------------------------------------------------------------------------------------
def toy_gen_forward(mlp, z):
    """ReLU MLP 2-128-128-128-2; ``mlp`` = [(W,b)]*4.  Reference: toy_example.py:26-34."""
    h = z
    for i, (W, b) in enumerate(mlp):
        h = F.linear(h, W, b)
        if i + 1 < len(mlp):
            h = F.relu(h)
    return h


def toy_langevin_posterior(z, x, mlp, steps, with_noise, step_size, noise=None, sigma=0.25):
    """Reference: toy_example.py:110-131 (sigma hard-coded .25 at :117, no EBM term: en = |z|^2/2 at :118)."""
    z = z.detach().clone().requires_grad_(True)
    for i in range(steps):
        x_hat = toy_gen_forward(mlp, z)
        llhd = 1.0 / (2.0 * sigma ** 2) * torch.sum((x_hat - x) ** 2)
        en = 0.5 * torch.sum(z ** 2)
        g = torch.autograd.grad(llhd + en, z)[0]
        z.data = z.data - 0.5 * step_size * step_size * g
        if with_noise:
            z.data += step_size * (noise[i] if noise is not None else torch.randn_like(z))
    return z.detach()


# ----------------------------------------------------------------------------------------------------------------------
# DAMC sampler
# ----------------------------------------------------------------------------------------------------------------------
def logsnr_schedule(t, logsnr_min, logsnr_max):
    """lambda(t) = -2 log tan(a t + b).  Reference: diffusion_helper_func.py:41-50."""
    b = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_max)))
    a = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_min))) - b
    return -2.0 * torch.log(torch.tan(a * t + b))


def _css_layer(L, ctx, x):
    """ConcatSquashLinearSkipCtx.  ``L`` = dict(W,b,Wc,bc,Wg,bg,Wb,Ws,bs).  Reference: diffusion_net.py:439-445."""
    c = F.silu(F.linear(F.silu(ctx), L["Wc"], L["bc"]))
    gate = torch.sigmoid(F.linear(c, L["Wg"], L["bg"]))
    return F.linear(x, L["W"], L["b"]) * gate + F.linear(c, L["Wb"]) + F.linear(x, L["Ws"], L["bs"])


def denoiser_eps(P, z, logsnr, xemb):
    """eps-prediction of Diffusion_UnetA.  ``P`` = dict(Wt1,bt1,Wt2,bt2,B,layers=[7 dicts],residual,ntemb).
    Reference: diffusion_net.py:497-533 (+ SinusoidalPosEmb :453-460)."""
    u = torch.arctan(torch.exp(-0.5 * torch.clamp(logsnr, min=-20.0, max=20.0))) / (0.5 * math.pi)
    half = P["ntemb"] // 2
    freq = torch.exp(torch.arange(half, device=z.device) * -(math.log(10000) / (half - 1))).to(z.dtype)
    arg = (u * 1000.0)[:, None] * freq[None, :]
    pe = torch.cat((arg.sin(), arg.cos()), dim=-1)
    temb = F.linear(F.silu(F.linear(pe, P["Wt1"], P["bt1"])), P["Wt2"], P["bt2"])
    ctx = torch.cat([temb, xemb], dim=1)
    proj = 2 * math.pi * (z @ P["B"])
    out = torch.cat([torch.sin(proj), torch.cos(proj), z], dim=1)
    skips = []
    for L in P["layers"][:3]:
        out = _css_layer(L, ctx, out)
        skips.append(out)
        out = F.leaky_relu(out, 0.01)
    out = _css_layer(P["layers"][3], ctx, out)
    for L in P["layers"][4:]:
        out = F.leaky_relu(torch.cat([out, skips.pop()], dim=1), 0.01)
        out = _css_layer(L, ctx, out)
    return z + out if P["residual"] else out


def damc_reverse_coeffs(i, T, logsnr_min, logsnr_max, var_type, dtype=torch.float64):
    """Batch-constant scalars of reverse step i: (lambda_t, c_pred, c_eps, c_zt, c_x, std).
    pred_z = c_pred*(z - eps*c_eps) (helper :36-39); z_s = c_zt*z + c_x*pred_z + std*noise (helper :52-70)."""
    it = torch.tensor([float(i)], dtype=dtype)
    lt = logsnr_schedule(it / (T - 1.0), logsnr_min, logsnr_max)
    ls = logsnr_schedule(torch.clamp(it - 1.0, min=0.0) / (T - 1.0), logsnr_min, logsnr_max)
    c_pred = torch.sqrt(1.0 + torch.exp(-lt))
    c_eps = torch.rsqrt(1.0 + torch.exp(lt))
    alpha_st = torch.sqrt((1.0 + torch.exp(-lt)) / (1.0 + torch.exp(-ls)))
    alpha_s = torch.sqrt(torch.sigmoid(ls))
    r = torch.exp(lt - ls)
    omr = -torch.expm1(lt - ls)
    if var_type == "large":
        var = omr * torch.sigmoid(-lt)
    elif var_type == "small":
        a_t, a_s = torch.sigmoid(lt), torch.sigmoid(ls)
        var = (1.0 - a_s) / (1.0 - a_t) * (1 - a_t / a_s)
    else:
        raise NotImplementedError(var_type)
    return tuple(float(v) for v in (lt, c_pred, c_eps, r * alpha_st, omr * alpha_s, torch.sqrt(var)))


def damc_sample(P, xemb, z_T, T, logsnr_min, logsnr_max, var_type, with_noise, noise=None, return_all=False):
    """DAMC ancestral sampler from a given xemb and initial z_T.  Reference: diffusion_net.py:595-622.
    noise: [T-1,B,nz]; noise[k] is consumed at the k-th executed reverse step (i = T-1-k), as the reference draws."""
    zt, hist = z_T, []
    b = len(zt)
    for k, i in enumerate(reversed(range(T))):
        it = torch.ones(b, dtype=zt.dtype) * float(i)
        lt = logsnr_schedule(it / (T - 1.0), logsnr_min, logsnr_max)
        ls = logsnr_schedule(torch.clamp(it - 1.0, min=0.0) / (T - 1.0), logsnr_min, logsnr_max)
        eps = denoiser_eps(P, zt, lt, xemb)
        lt, ls = lt.reshape(b, 1), ls.reshape(b, 1)
        pred = torch.sqrt(1.0 + torch.exp(-lt)) * (zt - eps * torch.rsqrt(1.0 + torch.exp(lt)))
        if i == 0:
            zt = pred
        else:
            alpha_st = torch.sqrt((1.0 + torch.exp(-lt)) / (1.0 + torch.exp(-ls)))
            alpha_s = torch.sqrt(torch.sigmoid(ls))
            r = torch.exp(lt - ls)
            omr = -torch.expm1(lt - ls)
            mean = r * alpha_st * zt + omr * alpha_s * pred
            if var_type == "large":
                var = omr * torch.sigmoid(-lt)
            else:
                a_t, a_s = torch.sigmoid(lt), torch.sigmoid(ls)
                var = (1.0 - a_s) / (1.0 - a_t) * (1 - a_t / a_s)
            e = noise[k] if noise is not None else torch.randn_like(zt)
            zt = mean + torch.sqrt(var) * e if with_noise else mean
        if return_all:
            hist.append(zt.clone())
    return (zt, hist) if return_all else zt
