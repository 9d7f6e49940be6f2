"""Would chain chunks small enough for a layer's output to stay in L2 pay?  Posterior Langevin (CIFAR-10 shape, bf16, K = 30) at
batch sizes around the L2 fit (a3 + g3 = 0.52 MB per chain: <= 100 MB up to 192 chains)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import bench
from damc_b200 import MCMC
dev = torch.device("cuda:0")
G, E = bench.make_nets(dev)
out = {}
for B in (96, 128, 192, 256, 384, 512, 1024):
    z0, x = bench.make_inputs(G, B, dev, 3)
    z0 = z0.to(dev)
    f = lambda: MCMC.sample_langevin_post_z_with_prior(z0.clone().requires_grad_(True), x, G, E, 30, 0.1, True, 0.1, seed=1, precision="bf16")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"B{B}"] = {"ms": round(ms, 3), "chain_steps_per_s": round(B * 30 / ms * 1e3)}
print(json.dumps(out))
