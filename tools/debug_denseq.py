"""Hoisted-context DAMC schedule (denoiser_seq.cu) against the other schedules and the fp32 kernel; timings."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200"), os.path.join(ROOT, "tests")]
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402
from oracle import synth  # noqa: E402

dev = torch.device("cuda:0")


def relmax(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max())


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


what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "parity"):
    T, nz = 12, 128
    Q = dn._netQ_U(nc=3, nz=nz, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10")
    Q.load_state_dict(synth.module_state_like(Q, prefix="Q."))
    Q = Q.to(dev).eval()
    for B in (1, 128, 130, 300):
        xemb = (0.5 * synth.det_normal("xe", (B, 1024))).to(dev)
        zT = synth.det_normal("zT", (B, nz))
        res = {}
        for prec in ("fp32", "fp16", "bf16"):
            for seq in ("1", "0"):
                if prec == "fp32" and seq == "0":
                    continue
                os.environ["DAMC_DEN_SEQ"] = seq
                res[prec, seq] = MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision=prec).cpu()
                torch.cuda.synchronize()
        f32 = res["fp32", "1"]
        print(f"B={B}: fp16 seq-vs-fp32 {relmax(res['fp16', '1'], f32):.3e}  old-vs-fp32 {relmax(res['fp16', '0'], f32):.3e}   "
              f"bf16 seq-vs-fp32 {relmax(res['bf16', '1'], f32):.3e}  old-vs-fp32 {relmax(res['bf16', '0'], f32):.3e}   "
              f"finite {bool(torch.isfinite(res['fp16', '1']).all())}", flush=True)
if what in ("all", "time"):
    torch.manual_seed(1)
    Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=100, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
    for B in (128, 1024, 4096, 16384):
        xemb = torch.randn(B, 1024, device=dev) * 0.5
        zT = torch.randn(B, 128, device=dev)
        line = f"B={B} T=100 (xemb given):"
        for seq in ("1", "0"):
            os.environ["DAMC_DEN_SEQ"] = seq
            ms = timed(lambda: MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision="fp16"))
            line += f"  seq={seq}: {ms:.3f} ms ({B * 100 / ms * 1e3 / 1e6:.2f} M steps/s, {B * 100 * 2.949e6 / ms * 1e3 / 1e12:.0f} TFLOP/s)"
        print(line, flush=True)
