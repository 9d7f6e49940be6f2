"""GPU debugging aid: K-step posterior error vs the fp64 oracle for combinations of (EBM on/off, noise on/off).
   python tools/debug_steps.py cifar10 128 128 3 4 5 [fp32|bf16]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import __graft_entry__ as ge  # noqa: E402

ge.build()
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402
from oracle import damc_oracle as O, synth  # noqa: E402

dataset, nz, ngf, nc, B, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6])
prec = sys.argv[7] if len(sys.argv) > 7 else "fp32"
sigma = 0.1
dev = torch.device("cuda:0")
layers = synth.gen_layers(dataset, nz, ngf, nc)
gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, K, sigma)
G, E = dn._netG(dataset, nz, ngf, nc), dn._netE(nz)
G.load_state_dict(gsd)
E.load_state_dict(esd)
G, E = G.to(dev), E.to(dev)
gen = synth.gen_list_from_state(gsd, layers, torch.float64)
ebm = synth.ebm_list_from_state(esd, torch.float64)
for use_e in (False, True):
    for use_n in (False, True):
        for k in sorted({1, 2, K}):
            ref = O.langevin_posterior_analytic(z0.double(), x.double(), gen, ebm if use_e else None, k, sigma, use_n,
                                                0.1, noise.double()[:k])
            z = z0.to(dev).clone().requires_grad_(True)
            out = MCMC.sample_langevin_post_z_with_prior(z, x.to(dev), G, E if use_e else None, k, sigma, use_n, 0.1,
                                                         noise=noise[:k].to(dev).contiguous() if use_n else None,
                                                         precision=prec)
            err = float((out.cpu().double() - ref).abs().max() / ref.abs().max())
            print(f"ebm={use_e} noise={use_n} K={k}: rel err {err:.3e}")
