"""Debug: damc_gemm_tf32 against torch.matmul over the shapes denoiser_train uses."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import denoiser_train as dt  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
shapes = [(12, 1408, 1152), (128, 1408, 1152), (128, 64, 128), (128, 256, 256), (128, 512, 512), (128, 512, 256), (12, 256, 512),
          (512, 256, 32), (512, 256, 128), (256, 128, 32), (256, 128, 128), (1408, 1152, 32), (1408, 1152, 128), (128, 1152, 1408),
          (256, 512, 128), (512, 512, 128), (300, 272, 96)]
for M, N, K in shapes:
    A, W, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.randn(N, device=dev)
    ref = A.double() @ W.double().t()
    for bias in (None, b):
        out = dt.gemm(A, W, bias)
        r = ref + (bias.double() if bias is not None else 0)
        e = float((out.double() - r).abs().max() / r.abs().max())
        bad = (out.double() - r).abs() > 0.05 * r.abs().max()
        print(f"M={M} N={N} K={K} bias={bias is not None}: rel err {e:.3e}" + (f"  BAD rows {sorted(set(bad.nonzero()[:, 0].tolist()))[:8]} cols {sorted(set(bad.nonzero()[:, 1].tolist()))[:8]}" if e > 0.02 else ""))
# strided output (column slice of a wider tensor)
big = torch.zeros(128, 512, device=dev)
A, W = torch.randn(128, 256, device=dev), torch.randn(256, 256, device=dev)
dt.gemm(A, W, None, out=big[:, 256:])
print("strided out err", float((big[:, 256:].double() - A.double() @ W.double().t()).abs().max()), "left half untouched", float(big[:, :256].abs().max()))
