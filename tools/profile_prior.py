"""Small driver for ncu: prior Langevin (persistent ebm_langevin_kernel), B chains x K steps in one launch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import bench
from damc_b200 import MCMC
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
_, E = bench.make_nets(dev)
z0 = torch.randn(B, 128, device=dev)
for rep in range(2):
    MCMC.sample_langevin_prior_z(z0.clone().requires_grad_(True), E, K, 0.4, True, seed=rep)
torch.cuda.synchronize()
print("ok")
