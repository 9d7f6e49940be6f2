"""GPU debugging aid: run ONE posterior step and compare every intermediate tensor in the library's workspace
(activations NHWC, parity-planar gradients, dz partial sums) with the CPU oracle.  Usage:
   python tools/debug_layers.py cifar10 128 128 3 4 [fp32|bf16] [sigma]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "diffusion-amortized-mcmc_b200")]
import __graft_entry__ as ge  # noqa: E402

ge.build()
from damc_b200 import MCMC, diffusion_net as dn  # noqa: E402
from oracle import synth  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def main():
    dataset, nz, ngf, nc, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    prec = sys.argv[6] if len(sys.argv) > 6 else "fp32"
    sigma = float(sys.argv[7]) if len(sys.argv) > 7 else 0.1
    dev = torch.device("cuda:0")
    layers = synth.gen_layers(dataset, nz, ngf, nc)
    gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, 1, sigma)
    start_k = int(os.environ.get("START_K", "0"))
    if start_k:  # start from the oracle's z after start_k noise-free steps
        from oracle import damc_oracle as O
        z0 = O.langevin_posterior_analytic(z0.double(), x.double(), synth.gen_list_from_state(gsd, layers, torch.float64),
                                           None, start_k, sigma, False, 0.1).float()
    G = dn._netG(dataset, nz, ngf, nc)
    G.load_state_dict(gsd)
    G = G.to(dev)
    # oracle intermediates (fp64)
    gen = synth.gen_list_from_state(gsd, layers, torch.float64)
    acts, h = [], z0.double().reshape(B, nz, 1, 1)
    for i, (W, b, s, p) in enumerate(gen):
        h = F.conv_transpose2d(h, W, b, stride=s, padding=p)
        if i + 1 < len(gen):
            h = torch.where(h > 0, h, 0.2 * h)
            acts.append(h)
    xhat = torch.tanh(h)
    g = (xhat - x.double()) / sigma ** 2 * (1 - xhat ** 2)
    grads = [None] * (len(gen) - 1)
    for i in range(len(gen) - 1, -1, -1):
        W, b, s, p = gen[i]
        g = F.conv2d(g, W, None, stride=s, padding=p)
        if i > 0:
            g = g * torch.where(acts[i - 1] > 0, 1.0, 0.2)
            grads[i - 1] = g
    gz = g.reshape(B, nz)
    # ours: one step, no noise, no EBM
    z = z0.to(dev).clone().requires_grad_(True)
    xh = torch.empty(B, nc, x.shape[2], x.shape[3], device=dev)
    MCMC.sample_langevin_post_z_with_prior(z, x.to(dev), G, None, 1, sigma, False, 0.1, precision=prec, x_hat_out=xh)
    torch.cuda.synchronize()
    print("xhat rel err", rel(xh.cpu(), xhat))
    ws = MCMC._workspaces[(dev.type, dev.index)]
    es = 2 if prec == "bf16" else 4
    tdt = torch.bfloat16 if prec == "bf16" else torch.float32
    al = lambda v: (v + 255) // 256 * 256
    nz_p = (nz + 63) // 64 * 64
    off = al(es * B * nz_p)
    for l, (cin, cout, k, s, p) in enumerate(layers[:-1]):
        H = acts[l].shape[2]
        n = B * H * H * cout
        a = ws[off:off + es * n].view(tdt).float().reshape(B, H, H, cout).permute(0, 3, 1, 2).cpu()
        off += al(es * n)
        graw = ws[off:off + es * n].view(tdt).float()
        off += al(es * n)
        if prec == "bf16" and not os.environ.get("DAMC_TC_NOBITS") and os.environ.get("DAMC_TC", "1") != "0":
            off += al(n // 8)  # 1-bit LeakyReLU masks of this layer (tcgen05 engine)
        if l == 0:
            gg = graw.reshape(B, H, H, cout).permute(0, 3, 1, 2).cpu()
        else:  # planar [py][px][B][H/2][W/2][C]
            gp = graw.reshape(2, 2, B, H // 2, H // 2, cout).cpu()
            gg = torch.zeros(B, H, H, cout)
            for py in range(2):
                for px in range(2):
                    gg[:, py::2, px::2, :] = gp[py, px]
            gg = gg.permute(0, 3, 1, 2)
        d = (gg.double() - grads[l]).abs()
        nbad = int((d > 1e-4 * grads[l].abs().max()).sum())
        print(f"layer {l}: act rel err {rel(a, acts[l]):.3e}   grad rel err {rel(gg, grads[l]):.3e}  "
              f"(elements off by >1e-4 of max: {nbad} of {d.numel()}; |act|max {float(acts[l].abs().max()):.2f})")
    znew = z.detach().cpu().double()
    gz_ours = (z0.double() - znew) / (0.5 * 0.1 * 0.1) - z0.double()
    print("dz rel err", rel(gz_ours, gz), " |gz|max", float(gz.abs().max()))


main()
