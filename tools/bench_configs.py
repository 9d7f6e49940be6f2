"""Posterior-Langevin throughput on every BASELINE.json config shape (1 GPU, CUDA events, synthetic inputs, default-init
weights seed 1): SVHN (configs[1]), CIFAR-10 (configs[2]), CelebA-HQ 256x256 (configs[3]), MNIST anomaly sweep (configs[4])."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402

dev = torch.device("cuda:0")
CASES = {  # name: (ctor, nz, nc, H, K, sigma, noise, MMAC per chain (fwd), batch sizes)
    "svhn": (lambda: dn._netG_svhn(100, 64, 3), 100, 3, 32, 30, 0.1, True, 69.50, (1024, 16384)),
    "cifar10": (lambda: dn._netG_cifar10(128, 128, 3), 128, 3, 32, 30, 0.1, True, 1089.2, (128, 1024, 4096)),
    "celebaHQ": (lambda: dn._netG_celebaHQ(128, 128, 3), 128, 3, 256, 30, 1.0, True, 6547.0, (8, 128)),
    "mnist": (lambda: dn._netG_mnist(8, 128, 1), 8, 1, 28, 5, 1.0, False, 824.3, (500, 4096, 65536)),
}
only = sys.argv[1:] or list(CASES)
prec = os.environ.get("PREC", "bf16")
out = {}
for name in only:
    ctor, nz, nc, H, K, sigma, noise, mmac, sizes = CASES[name]
    torch.manual_seed(1)
    G, E = ctor().to(dev).eval(), dn._netE(nz).to(dev).eval()
    for B in sizes:
        x = torch.rand(B, nc, H, H, device=dev) * 2 - 1
        z0 = torch.randn(B, nz, device=dev)
        fn = lambda: MCMC.sample_langevin_post_z_with_prior(z0.clone().requires_grad_(True), x, G, E, K, sigma, noise,
                                                            0.1, seed=1, precision=prec)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            o = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        cs = B * K / ms * 1e3
        tf = cs * 4 * mmac * 1e6 / 1e12
        assert torch.isfinite(o).all()
        out[f"{name}_B{B}"] = {"ms": round(ms, 3), "chain_steps_per_s": round(cs), "algorithmic_tflops": round(tf, 1)}
        print(f"{name} B={B} K={K} [{prec}]: {ms:.2f} ms  {cs:,.0f} chain-steps/s  {tf:.0f} TFLOP/s", flush=True)
print(json.dumps(out))
