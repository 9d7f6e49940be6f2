"""Secondary timings on one GPU (CUDA events): prior Langevin, DAMC sampler, toy posterior, small-batch posterior."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import bench  # noqa: E402
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402

dev = torch.device("cuda:0")
out = {}


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


G, E = bench.make_nets(dev)
# prior Langevin: 2B = 256 chains x 60 steps (train_gen_recon.py:206-209) and 16384 chains
for B in (256, 16384):
    z0 = torch.randn(B, 128, device=dev)
    ms = timed(lambda: MCMC.sample_langevin_prior_z(z0.clone().requires_grad_(True), E, 60, 0.4, True, seed=1))
    out[f"prior_B{B}_K60"] = {"ms": ms, "chain_steps_per_s": B * 60 / ms * 1e3}
# DAMC sampler: Q(x), T = 100
torch.manual_seed(1)
Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=100, logsnr_min=-5.1,
               logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
for B in (128, 4096):
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    with torch.no_grad():
        ms_enc = timed(lambda: Q.encoder(x), reps=3)
        out[f"damc_B{B}_T100_torch_encoder"] = {"ms": ms_enc}
        for prec in ("fp32", "fp16", "bf16"):  # fp32: persistent CUDA-core kernel + torch encoder; fp16/bf16: tcgen05 + library encoder
            ms = timed(lambda: Q(x, precision=prec), reps=3)
            out[f"damc_B{B}_T100_{prec}"] = {"ms": ms, "reverse_steps_per_s": B * 100 / ms * 1e3}
        ms_p = timed(lambda: Q(x=None, b=B, device=dev, precision="fp16"), reps=3)
        out[f"damc_prior_B{B}_T100_fp16"] = {"ms": ms_p, "reverse_steps_per_s": B * 100 / ms_p * 1e3}
# posterior at the reference's training batch (B = 128), fp32 and bf16
_, x = bench.make_inputs(G, 128, dev, 3)
z0 = torch.randn(128, 128, device=dev)
for prec in ("bf16", "fp32"):
    ms = timed(lambda: MCMC.sample_langevin_post_z_with_prior(z0.clone().requires_grad_(True), x, G, E, 30, 0.1, True, 0.1,
                                                              seed=1, precision=prec), reps=3)
    out[f"posterior_B128_K30_{prec}"] = {"ms": ms, "chain_steps_per_s": 128 * 30 / ms * 1e3}
# toy: 500 chains x 1000 steps
toyG = torch.nn.Module()
toyG.net = torch.nn.Sequential(torch.nn.Linear(2, 128), torch.nn.ReLU(), torch.nn.Linear(128, 128), torch.nn.ReLU(),
                               torch.nn.Linear(128, 128), torch.nn.ReLU(), torch.nn.Linear(128, 2))
toyG = toyG.to(dev)
zt, xt = torch.randn(500, 2, device=dev), torch.randn(500, 2, device=dev)
ms = timed(lambda: MCMC.sample_langevin_post_z(zt.clone().requires_grad_(True), xt, toyG, 1000, True, 0.1, seed=1), reps=3)
out["toy_B500_K1000"] = {"ms": ms, "chain_steps_per_s": 500 * 1000 / ms * 1e3}
print(json.dumps(out, indent=1))
