"""DAMC sampler loop only (xemb given, T = 100, Philox noise): hoisted-context schedule vs the other tensor-core schedules
(DAMC_DEN_SEQ=0: 8-CTA cluster kernel up to 1 024 chains, per-layer launches + CUDA graph above) and the fp32 kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402

dev = torch.device("cuda:0")
FLOP_PER_STEP = 2.949e6   # 1.4746 MMAC per chain and reverse step
PEAK = 1410.6e12          # measured sustained bf16 (MEASURED_PEAKS.json)


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


torch.manual_seed(1)
T = 100
Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
               logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
out = {}
for B in (128, 1024, 4096, 16384):
    xemb = torch.randn(B, 1024, device=dev) * 0.5
    zT = torch.randn(B, 128, device=dev)   # (a host z_T would put a pageable 8 MB copy per call into the timing at 16 384 chains)
    for prec in ("fp16", "bf16"):
        for seq, name in (("1", "hoisted"), ("0", "other")):
            os.environ["DAMC_DEN_SEQ"] = seq
            ms = timed(lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision=prec))
            fl = B * T * FLOP_PER_STEP / (ms * 1e-3)
            out[f"B{B}_{prec}_{name}"] = {"ms": round(ms, 3), "reverse_steps_per_s": round(B * T / ms * 1e3), "tflops": round(fl / 1e12, 1),
                                          "frac_of_sustained_bf16": round(fl / PEAK, 3)}
    os.environ.pop("DAMC_DEN_SEQ", None)
    if B <= 4096:
        ms = timed(lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp32"), reps=2, warm=1)
        out[f"B{B}_fp32_kernel"] = {"ms": round(ms, 3), "reverse_steps_per_s": round(B * T / ms * 1e3)}
print(json.dumps(out, indent=1))
