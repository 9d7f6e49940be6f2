import os, sys, torch
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn
dev = torch.device("cuda:0")
def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
import time
for T in (2, 10, 50, 100):
    torch.manual_seed(1)
    Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
    for B in (128,):
        xemb = torch.randn(B, 1024, device=dev) * 0.5
        zT = torch.randn(B, 128, device=dev)
        for seq in ("1", "0"):
            os.environ["DAMC_DEN_SEQ"] = seq
            ms = timed(lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp16"))
            t0 = time.perf_counter()
            for _ in range(10): MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp16")
            host = (time.perf_counter() - t0) / 10 * 1e3
            torch.cuda.synchronize()
            print(f"T={T} B={B} seq={seq}: {ms:.3f} ms per call (host enqueue {host:.3f} ms)", flush=True)
