"""One DAMC sampler call on the hoisted-context schedule (for ncu launch lists): python tools/profile_denseq.py B [T]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn
dev = torch.device("cuda:0")
torch.manual_seed(1)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 100
Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
               logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
xemb = torch.randn(B, 1024, device=dev) * 0.5
zT = torch.randn(B, 128, device=dev)
for i in range(2):
    MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp16")
    torch.cuda.synchronize()
