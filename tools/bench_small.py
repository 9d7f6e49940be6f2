"""Posterior Langevin at small chain counts (CIFAR-10 shape, bf16): wave-quantisation check."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import bench
from damc_b200 import MCMC
dev = torch.device("cuda:0")
G, E = bench.make_nets(dev)
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 192, 256, 384, 512, 1024]:
    _, x = bench.make_inputs(G, B, dev, 3)
    z0 = torch.randn(B, 128, device=dev)
    fn = lambda: MCMC.sample_langevin_post_z_with_prior(z0.clone().requires_grad_(True), x, G, E, 30, 0.1, True, 0.1, seed=1, precision="bf16")
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"B={B}: {ms:.2f} ms  {B * 30 / ms * 1e3:,.0f} chain-steps/s", flush=True)
