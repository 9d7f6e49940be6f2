import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
from damc_b200 import diffusion_net as dn
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.manual_seed(4)
B = 128
Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=100, logsnr_min=-5.1,
               logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev)
x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
z = torch.randn(B, 128, device=dev)
mask = (torch.rand(B, device=dev) >= 0.2).float().unsqueeze(-1)
res = {}
from damc_b200 import denoiser_train as dt
real_gemm = dt.gemm
for engine in ("torch", "library", "fake"):
    Qe = Q
    Qe.zero_grad(set_to_none=True)
    torch.manual_seed(77)
    def checked(A, W, bias=None, out=None):
        r = real_gemm(A, W, bias, out)
        ref = A.double() @ W.double().t() + (bias.double() if bias is not None else 0)
        e = float((r.double() - ref).abs().max() / (ref.abs().max() + 1e-30))
        if e > 5e-3:
            print(f"   gemm M={A.shape[0]} N={W.shape[0]} K={A.shape[1]} A.contig={A.is_contiguous()} W.contig={W.is_contiguous()} "
                  f"|A|max {float(A.abs().max()):.2e} |W|max {float(W.abs().max()):.2e} |ref|max {float(ref.abs().max()):.2e} rel err {e:.3e}")
        return r
    dt.gemm = checked if engine == "library" else (lambda A, W, bias=None, out=None: (A @ W.t() + bias) if bias is not None else A @ W.t())
    loss = Qe.calculate_loss(x=x, z=z, mask=mask, engine="torch" if engine == "torch" else "library").mean()
    loss.backward()
    res[engine] = (float(loss), {n: p.grad.detach().double().clone() for n, p in Qe.named_parameters() if p.grad is not None})
print({k: v[0] for k, v in res.items()})
g0, g1, g64 = res["torch"][1], res["library"][1], res["fake"][1]
rows = []
for n in g0:
    rows.append((float((g1[n] - g0[n]).abs().max()), float(g0[n].abs().max()), float((g0[n] - g64[n]).abs().max()) if n in g64 else -1, n))
rows.sort(key=lambda r: -r[0] / (r[1] + 1e-30))
for r in rows[:12]:
    print(f"{r[3]:45s} |lib - torch| {r[0]:.3e}  |torch| {r[1]:.3e}  |fake-gemm engine - torch| {r[2]:.3e}")
