"""The caller on either side of the hot path: one training iteration and the eval consumers, on a GPU."""
import copy

import pytest
import torch

from oracle import damc_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


def test_training_iteration_updates_all_three_networks(dev):
    from damc_b200 import diffusion_net as dn, train
    torch.manual_seed(0)
    nz, ngf, B = 128, 16, 12
    G, E = dn._netG_cifar10(nz, ngf, 3).to(dev), dn._netE(nz).to(dev)
    Q = dn._netQ_U(nc=3, nz=nz, nxemb=256, ntemb=128, nif=16, diffusion_residual=True, n_interval=8, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev)
    Q_dummy = copy.deepcopy(Q)
    opts = [torch.optim.Adam(G.parameters(), lr=2e-4, betas=(0.5, 0.999)),
            torch.optim.Adam(E.parameters(), lr=1e-4, betas=(0.5, 0.999)),
            torch.optim.AdamW(Q.parameters(), lr=2e-4, weight_decay=1e-2, betas=(0.5, 0.999))]
    before = [[p.detach().clone() for p in m.parameters()] for m in (G, E, Q)]
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    cfg = train.TrainConfig(g_l_steps=5, e_l_steps=8, q_updates=2, precision="fp32")   # ngf = 16: below tensor-core granularity
    out = train.training_iteration(x, G, E, Q, Q_dummy, *opts, cfg=cfg)
    for k in ("q_loss", "g_loss", "e_loss"):
        assert torch.isfinite(out[k]), k
    assert out["zk_pos"].shape == (B, nz) and out["zk_neg"].shape == (2 * B, nz)
    for m, old in zip((G, E, Q), before):
        assert all(p.requires_grad for p in m.parameters())          # samplers restore requires_grad (MCMC.py:72-73)
        assert any(not torch.equal(p.detach(), o) for p, o in zip(m.parameters(), old))
    # the packed-weight cache must see the optimiser's in-place updates AND .data EMA copies on the next call
    train.ema_update(Q, Q_dummy, rho=0.5)
    out2 = train.training_iteration(x, G, E, Q, Q_dummy, *opts, cfg=cfg)
    assert torch.isfinite(out2["g_loss"])


def test_sampler_sees_in_place_weight_updates(dev):
    """param.data.copy_() does not bump tensor versions; the samplers must still use the new values."""
    from damc_b200 import MCMC, diffusion_net as dn
    nz, B, K = 100, 6, 5
    E = dn._netE(nz).to(dev)
    z0, noise = synth.det_normal("u.z0", (B, nz)), synth.det_normal("u.n", (K, B, nz))
    for seed in (1, 2):
        sd = synth.ebm_state(nz, seed=seed)
        for name, p in E.state_dict().items():
            p.data.copy_(sd[name])                                    # in place, no version bump
        out = MCMC.sample_langevin_prior_z(z0.to(dev).clone().requires_grad_(True), E, K, 0.4, True, noise=noise.to(dev))
        ref = O.langevin_prior_analytic(z0.double(), synth.ebm_list_from_state(sd, torch.float64), K, 0.4, True,
                                        noise.double())
        assert float((out.cpu().double() - ref).abs().max()) < 1e-4, seed


def test_eval_consumers_match_oracle(dev):
    from damc_b200 import MCMC, diffusion_net as dn
    nz, ngf, nc, B = 8, 16, 1, 7
    layers = synth.gen_layers("mnist", nz, ngf, nc)
    gsd, esd, z0, x, _ = synth.synth_problem(layers, nz, B, 1, 1.0, seed=9)
    G, E = dn._netG("mnist", nz, ngf, nc), dn._netE(nz)
    G.load_state_dict(gsd)
    E.load_state_dict(esd)
    G, E = G.to(dev), E.to(dev)
    gen, ebm = synth.gen_list_from_state(gsd, layers, torch.float64), synth.ebm_list_from_state(esd, torch.float64)
    xh = O.gen_forward(gen, z0.double())
    score_ref = ((xh - x.double()) ** 2).sum((1, 2, 3)) + O.ebm_forward(ebm, z0.double()) + 0.5 * (z0.double() ** 2).sum(1)
    score = MCMC.anomaly_score(x.to(dev), z0.to(dev), G, E, precision="fp32")
    assert float((score.cpu().double() - score_ref).abs().max() / score_ref.abs().max()) < 1e-5
    mse_ref = ((xh - x.double()) ** 2).mean((1, 2, 3)).sum()
    mse = MCMC.recon_mse(x.to(dev), z0.to(dev), G, precision="fp32")
    assert abs(float(mse) - float(mse_ref)) < 1e-5 * float(mse_ref)


@pytest.mark.parametrize("B", [12, 128])
def test_q_loss_on_library_gemms_matches_autograd(B, dev):
    """Q.calculate_loss(engine="library") -- the 35 Linear layers of Q.p forward and backward on damc_gemm_tf32 (tcgen05, TF32
    operands, fp32 accumulate) -- against the reference formulation in PyTorch autograd (diffusion_net.py:624-646 / :463-533),
    same random draws.  The loss must agree to TF32 rounding.  The gradients cannot agree element-wise: the network's
    LeakyReLU(0.01) kinks turn any 1e-3 perturbation of the pre-activations into ~30 derivative flips (1 <-> 0.01) per layer at
    128 chains, each moving one row of a weight gradient by ~10 % -- torch's own TF32 mode (allow_tf32) shows the same against
    its fp32 mode.  So the bar is: relative L2 error of the whole gradient, and of every tensor, within 3x of what torch's TF32
    matmuls produce on the same inputs (and the exact-GEMM stand-in of the same orchestration reproduces autograd to 1e-6)."""
    from damc_b200 import diffusion_net as dn, denoiser_train as dt
    torch.manual_seed(4)
    Q = dn._netQ_U(nc=3, nz=128, nxemb=1024, ntemb=128, nif=64, diffusion_residual=True, n_interval=100, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev)
    x = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
    z = torch.randn(B, 128, device=dev)
    mask = (torch.rand(B, device=dev) >= 0.2).float().unsqueeze(-1)

    def run(engine, allow_tf32=False, gemm=None):
        old_gemm, old_flag = dt.gemm, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = allow_tf32
        if gemm is not None:
            dt.gemm = gemm
        try:
            Q.zero_grad(set_to_none=True)
            torch.manual_seed(77)
            n0 = dt.lib().damc_launch_count()
            loss = Q.calculate_loss(x=x, z=z, mask=mask, engine=engine).mean()
            loss.backward()
            launches = dt.lib().damc_launch_count() - n0
            return float(loss), {n: p.grad.detach().double().clone() for n, p in Q.named_parameters() if p.grad is not None}, launches
        finally:
            dt.gemm, torch.backends.cuda.matmul.allow_tf32 = old_gemm, old_flag

    l0, g0, _ = run("torch")
    lt, gt, _ = run("torch", allow_tf32=True)                         # the reference under torch's TF32 matmuls
    l1, g1, launches = run("library")
    exact = lambda A, W, bias=None, out=None: (A @ W.t() + bias) if bias is not None else A @ W.t()
    le, ge, _ = run("library", gemm=exact)                            # same orchestration, exact fp32 products
    l2, g2, _ = run("library_graphed")                                # captures the graphs
    l3, g3, _ = run("library_graphed")                                # replays them
    assert l2 == l1 and l3 == l1                                      # graph replay == eager launches: the same loss bits ...
    gmax = max(float(v.abs().max()) for v in g1.values())
    for n in g1:                                                      # ... and the same gradients
        if n.startswith("p.in_layers") or n.startswith("p.mid_layers") or n.startswith("p.out_layers"):
            assert torch.equal(g3[n], g1[n]), n                       # the core's own outputs: bit for bit
        else:                                                         # downstream of the core through cuDNN / cuBLAS backward kernels
            assert float((g3[n] - g1[n]).abs().max()) <= 1e-5 * gmax, n
    assert set(g0) == set(g1) == set(ge) and launches == 46           # 16 forward + 30 backward GEMM launches on the library (rounding passes not counted)
    norm = lambda g: sum(float((v ** 2).sum()) for v in g.values()) ** 0.5
    diff = lambda a, b: sum(float(((a[n] - b[n]) ** 2).sum()) for n in a) ** 0.5
    e_exact, e_lib, e_tf = diff(ge, g0) / norm(g0), diff(g1, g0) / norm(g0), diff(gt, g0) / norm(g0)
    print(f"B={B}: loss fp32 {l0:.6f} / torch-TF32 {lt:.6f} / library {l1:.6f}; relative L2 error of the whole gradient: exact-GEMM "
          f"stand-in {e_exact:.2e}, library {e_lib:.3e}, torch TF32 {e_tf:.3e}")
    assert abs(le - l0) <= 1e-5 * abs(l0) and e_exact < 1e-5
    assert abs(l1 - l0) <= 2e-3 * abs(l0), (l0, l1)
    assert e_lib < max(1.5 * e_tf, 2e-3), (e_lib, e_tf)
    big = [n for n in g0 if float(g0[n].abs().max()) > 1e-3 * max(float(v.abs().max()) for v in g0.values())]
    for n in big:
        a = float(((g1[n] - g0[n]) ** 2).sum()) ** 0.5 / (float((g0[n] ** 2).sum()) ** 0.5)
        b = float(((gt[n] - g0[n]) ** 2).sum()) ** 0.5 / (float((g0[n] ** 2).sum()) ** 0.5)
        assert a < max(3 * b, 0.08), (n, a, b)


@pytest.mark.parametrize("M,N,K", [(128, 1408, 1152), (12, 256, 512), (512, 256, 32), (300, 272, 96), (1408, 1152, 128)])
def test_gemm_tf32_abi_matches_matmul(M, N, K, dev):
    """damc_gemm_tf32 (the tcgen05 engine's plain-GEMM plan, kind::tf32) through the C ABI: D = A W^T (+ bias), ragged M, N that is
    not a multiple of the 256-column tile, one 32-wide k-block, a strided output (column slice of a wider tensor)."""
    from damc_b200 import denoiser_train as dt
    torch.manual_seed(M + N + K)
    A, W, b = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.randn(N, device=dev)
    ref = A.double() @ W.double().t()
    for bias in (None, b):
        out = dt.gemm(A, W, bias)
        r = ref + (bias.double() if bias is not None else 0)
        assert float((out.double() - r).abs().max() / r.abs().max()) < 1e-3       # TF32 operands (rounded), fp32 accumulation
    wide = torch.full((M, N + 64), 7.0, device=dev)
    dt.gemm(A, W, None, out=wide[:, 32:32 + N])
    assert float((wide[:, 32:32 + N].double() - ref).abs().max() / ref.abs().max()) < 1e-3
    assert bool((wide[:, :32] == 7.0).all()) and bool((wide[:, 32 + N:] == 7.0).all())   # nothing outside the slice is touched
    t = dt._tf32(A)
    assert float((t - A).abs().max() / A.abs().max()) < 2.0 ** -11 and bool(((t.view(torch.int32) & 0x1FFF) == 0).all())
