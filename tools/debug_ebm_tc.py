"""Debug: dE/dz recovered from ONE noise-free posterior step with the likelihood switched off (sigma = 1e3), tensor-core EBM tail
vs CUDA-core tail vs the fp64 oracle."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402
from oracle import damc_oracle as O, synth  # noqa: E402

dev = torch.device("cuda:0")
for dataset, nz, ngf, nc, B in (("mnist", 8, 128, 1, 200), ("svhn", 100, 64, 3, 333), ("cifar10", 128, 64, 3, 200)):
    layers = synth.gen_layers(dataset, nz, ngf, nc)
    gsd, esd, z0, x, _ = synth.synth_problem(layers, nz, B, 1, 0.1, seed=3, gain=0.0)
    for amp in (1.0, 4.0):
        esd2 = {k: (v * amp if k.endswith("weight") else v) for k, v in esd.items()}
        G, E = dn._netG(dataset, nz, ngf, nc), dn._netE(nz)
        G.load_state_dict(gsd)
        E.load_state_dict(esd2)
        G, E = G.to(dev), E.to(dev)
        ebm64 = synth.ebm_list_from_state(esd2, torch.float64)
        gref = O.ebm_grad(ebm64, z0.double())[1]
        s = 0.1
        for prec in ("bf16", "fp16"):
            res = {}
            for tc in ("1", "0"):
                os.environ["DAMC_EBM_TC"] = tc
                z = z0.to(dev).clone().requires_grad_(True)
                out = MCMC.sample_langevin_post_z_with_prior(z, x.to(dev), G, E, 1, 1e3, False, s, precision=prec).cpu().double()
                res[tc] = (z0.double() - out) / (0.5 * s * s) - z0.double()
            os.environ.pop("DAMC_EBM_TC")
            per1 = ((res["1"] - gref).abs().amax(1) / gref.abs().amax(1))
            per0 = ((res["0"] - gref).abs().amax(1) / gref.abs().amax(1))
            print(f"{dataset} nz={nz} amp={amp} [{prec}]: |gE|max {float(gref.abs().max()):.3f}  tc per-chain err median {float(per1.median()):.3e} "
                  f"max {float(per1.max()):.3e} frac>1e-2 {float((per1 > 1e-2).float().mean()):.2f}   cuda-core median {float(per0.median()):.3e} max {float(per0.max()):.3e}")
