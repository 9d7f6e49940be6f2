"""Prior Langevin K = 60: fp32 persistent kernel vs the tensor-core form (precision="fp16")."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import bench
from damc_b200 import MCMC
dev = torch.device("cuda:0")
_, E = bench.make_nets(dev)
out = {}
for B in (256, 1024, 4096, 16384, 65536):
    z0 = torch.randn(B, 128, device=dev)
    for prec in ("fp32", "fp16"):
        f = lambda: MCMC.sample_langevin_prior_z(z0.clone().requires_grad_(True), E, 60, 0.4, True, seed=1, precision=prec)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[f"prior_B{B}_K60_{prec}"] = {"ms": round(ms, 3), "chain_steps_per_s": round(B * 60 / ms * 1e3)}
print(json.dumps(out))
