"""Deterministic synthetic weights / inputs / noise for the parity tests -- TEST INFRASTRUCTURE (see damc_oracle.py).

Values come from numpy's frozen legacy ``RandomState`` stream keyed by (seed, tensor name), so the fixture generator
(run once in the build container against the real reference) and the tests (run anywhere, including the GPU box where
/root/reference does not exist) regenerate bit-identical fp32 tensors without committing megabytes of weights.
"""
import zlib

import numpy as np
import torch


def det_normal(name, shape, seed=0, scale=1.0):
    rs = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
    return torch.from_numpy((rs.standard_normal(tuple(shape)) * scale).astype(np.float32))


def det_uniform(name, shape, seed=0, bound=1.0):
    rs = np.random.RandomState((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
    return torch.from_numpy(rs.uniform(-bound, bound, tuple(shape)).astype(np.float32))


def generator_state(layers, seed=0, gain=0.85, bias_scale=0.05):
    """layers: list of (cin, cout, k, stride, pad).  Returns state dict with the reference's key names
    ``gen.{2i}.weight`` [Cin,Cout,k,k] / ``gen.{2i}.bias``.

    gain > 0 ("trained-like"): N(0, (gain/sqrt(effective fan-in))^2) so that pre-activations stay O(1) and LeakyReLU
      kinks / tanh curvature are exercised.  At full width this makes the reference's default step
      (s=0.1, sigma=0.1) expand perturbations ~50x per step, so K-step parity there is only meaningful for small K.
    gain == 0 ("default-init profile"): U(-b, b) with b = 1/sqrt(Cout k^2), the scale torch's default ConvTranspose2d
      init produces and the benchmark uses; well conditioned through K = 30."""
    sd = {}
    if gain == 0:
        for i, (cin, cout, k, s, p) in enumerate(layers):
            b = 1.0 / np.sqrt(cout * k * k)
            sd[f"gen.{2 * i}.weight"] = det_uniform(f"gen.{2 * i}.weight", (cin, cout, k, k), seed, b)
            sd[f"gen.{2 * i}.bias"] = det_uniform(f"gen.{2 * i}.bias", (cout,), seed, b)
        return sd
    for i, (cin, cout, k, s, p) in enumerate(layers):
        fan = cin * (k / s) ** 2 if i > 0 else cin
        sd[f"gen.{2 * i}.weight"] = det_normal(f"gen.{2 * i}.weight", (cin, cout, k, k), seed, gain / np.sqrt(fan))
        sd[f"gen.{2 * i}.bias"] = det_normal(f"gen.{2 * i}.bias", (cout,), seed, bias_scale)
    return sd


def ebm_state(nz, ndf=200, seed=0, gain=2.0):
    dims = [(ndf, nz), (ndf, ndf), (1, ndf)]
    sd = {}
    for i, (o, n) in enumerate(dims):
        sd[f"ebm.{2 * i}.weight"] = det_normal(f"ebm.{2 * i}.weight", (o, n), seed, gain / np.sqrt(n))
        sd[f"ebm.{2 * i}.bias"] = det_normal(f"ebm.{2 * i}.bias", (o,), seed, 0.1)
    return sd


def module_state_like(module, seed=0, prefix="", gain=1.0):
    """Deterministic values for every parameter of an arbitrary module (used for the amortizer Q): weights
    ~ gain/sqrt(fan_in), 1-D tensors small; InstanceNorm affine weights near 1."""
    sd = {}
    for name, t in module.state_dict().items():
        key = prefix + name
        if t.dim() >= 2:
            fan = t[0].numel()
            sd[name] = det_normal(key, t.shape, seed, gain / np.sqrt(fan))
        elif "net." in name and name.endswith("weight"):  # InstanceNorm2d affine scale
            sd[name] = 1.0 + det_normal(key, t.shape, seed, 0.1)
        else:
            sd[name] = det_normal(key, t.shape, seed, 0.1)
    return sd


def gen_layers(dataset, nz, ngf, nc):
    """(cin,cout,k,s,p) list for a reference generator family at a given width."""
    table = {
        "cifar10": [(8, 8, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 3, 1, 1)],
        "svhn": [(8, 4, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 4, 2, 1)],
        "celeba64": [(8, 4, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), (1, 4, 2, 1), ("nc", 4, 2, 1)],
        "celebaHQ": [(16, 4, 1, 0), (8, 4, 2, 1), (4, 4, 2, 1), (4, 4, 2, 1), (2, 4, 2, 1), (1, 4, 2, 1),
                     ("nc", 4, 2, 1)],
        "mnist": [(8, 7, 1, 0), (4, 4, 2, 1), (2, 4, 2, 1), ("nc", 3, 1, 1)],
    }[dataset]
    out, cin = [], nz
    for mult, k, s, p in table:
        cout = nc if mult == "nc" else ngf * mult
        out.append((cin, cout, k, s, p))
        cin = cout
    return out


def gen_list_from_state(sd, layers, dtype=torch.float32):
    return [(sd[f"gen.{2 * i}.weight"].to(dtype), sd[f"gen.{2 * i}.bias"].to(dtype), s, p)
            for i, (_, _, _, s, p) in enumerate(layers)]


def ebm_list_from_state(sd, dtype=torch.float32):
    return [(sd[f"ebm.{2 * i}.weight"].to(dtype), sd[f"ebm.{2 * i}.bias"].to(dtype)) for i in range(3)]


def denoiser_params_from_state(sd, residual, ntemb, dtype=torch.float32, prefix="p."):
    """Flatten a _netQ_U state dict's ``p.*`` entries into the dict damc_oracle.denoiser_eps expects."""
    g = lambda k: sd[prefix + k].to(dtype)
    layers = []
    for grp, n in (("in_layers", 3), ("mid_layers", 1), ("out_layers", 3)):
        for j in range(n):
            b = f"{grp}.{j}."
            layers.append(dict(W=g(b + "_layer.0.weight"), b=g(b + "_layer.0.bias"), Wc=g(b + "_layer_ctx.1.weight"),
                               bc=g(b + "_layer_ctx.1.bias"), Wg=g(b + "_hyper_gate.weight"),
                               bg=g(b + "_hyper_gate.bias"), Wb=g(b + "_hyper_bias.weight"), Ws=g(b + "_skip.weight"),
                               bs=g(b + "_skip.bias")))
    return dict(Wt1=g("time_mlp.1.weight"), bt1=g("time_mlp.1.bias"), Wt2=g("time_mlp.3.weight"),
                bt2=g("time_mlp.3.bias"), B=g("B"), layers=layers, residual=residual, ntemb=ntemb)


def synth_problem(layers, nz, B, K, sigma, seed=0, gain=0.85):
    """Weights + (z0, x, noise) for a posterior-Langevin case.  x = clamp(G(z*) + sigma*n, -1, 1)."""
    from . import damc_oracle as O
    gsd = generator_state(layers, seed, gain)
    esd = ebm_state(nz, seed=seed)
    gen = gen_list_from_state(gsd, layers)
    zstar = det_normal("zstar", (B, nz), seed)
    with torch.no_grad():
        x = O.gen_forward(gen, zstar)
        x = torch.clamp(x + sigma * det_normal("xnoise", x.shape, seed), -1.0, 1.0)
    z0 = det_normal("z0", (B, nz), seed)
    noise = det_normal("noise", (K, B, nz), seed)
    return gsd, esd, z0, x, noise
