"""DAMC denoiser loop timings (CUDA events, encoder excluded): fp32 persistent kernel vs tcgen05 per-layer GEMMs."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(1)
T = 100
Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
               logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()


def timed(fn, reps=3, warm=2):
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


out = {}
sizes = [int(a) for a in sys.argv[1:]] or [128, 1024, 4096, 16384, 65536]
for B in sizes:
    xemb = torch.randn(B, 1024, device=dev) * 0.5
    zT = torch.randn(B, 128, device=dev)
    for prec in os.environ.get("PRECS", "fp32,bf16,fp16").split(","):
        if prec == "fp32" and B > 16384:
            continue
        import time
        MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=3, precision=prec)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=3, precision=prec)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"   host enqueue {1e3 * (t1 - t0):.2f} ms, drain {1e3 * (t2 - t1):.2f} ms")
        ms = timed(lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=3, precision=prec))
        out[f"B{B}_{prec}"] = {"ms": round(ms, 3), "reverse_steps_per_s": round(B * T / ms * 1e3)}
        print(f"B={B} {prec}: {ms:.3f} ms  {B * T / ms * 1e3 / 1e6:.2f} M reverse steps/s", flush=True)
print(json.dumps(out))

# whole Q(x) call (encoder + T reverse steps), torch encoder vs library encoder
if os.environ.get("WITH_ENCODER", "1") == "1":
    for B in sizes:
        if B > 16384:
            continue
        x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
        with torch.no_grad():
            ms_t = timed(lambda: Q.encoder(x))
        ms_l = timed(lambda: MCMC.encoder_forward(Q.encoder, x, precision="fp16"))
        ms_q = timed(lambda: MCMC.damc_sample(Q, x=x, seed=3, precision="fp16"))
        print(f"B={B}: encoder torch {ms_t:.3f} ms, library fp16 {ms_l:.3f} ms; Q(x) fp16 end to end {ms_q:.3f} ms", flush=True)
