"""Round-2 GPU parity tests: the wrappers around the samplers (gen_samples, gen_samples_with_diffusion_prior, Q(x=None)),
x_hat_out, the CUDA-graph LRU of the posterior sampler, the change-detecting weight repack, and the statistical parity
(MMD / energy) of the tensor-core precisions against the reference algorithm.

Reference lines: src/MCMC.py:119-128 (gen_samples), :146-150 (gen_samples_with_diffusion_prior),
src/diffusion_net.py:585-595 (_netQ_U.forward, x=None branch), src/MCMC.py:55 (x_hat = netG(z))."""
import contextlib
import os

import numpy as np
import pytest
import torch

from oracle import damc_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def relmax(a, b):
    a = a.detach().double().cpu().numpy() if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    b = b.detach().double().cpu().numpy() if torch.is_tensor(b) else np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


@contextlib.contextmanager
def injected_randn(draws):
    """torch.randn -> pre-drawn tensors (moved to the requested device), like oracle/ref_shim.injected_noise."""
    it = iter(draws)
    real = torch.randn

    def fake(*a, **k):
        t = next(it).clone()
        return t.to(k["device"]) if k.get("device") is not None else t

    torch.randn = fake
    try:
        yield
    finally:
        torch.randn = real


def _nets(dataset, nz, ngf, nc, gsd, esd, dev):
    from damc_b200 import diffusion_net as dn
    G, E = dn._netG(dataset, nz, ngf, nc), dn._netE(nz)
    G.load_state_dict(gsd)
    E.load_state_dict(esd)
    return G.to(dev), E.to(dev)


def _Q(g, dev):
    from damc_b200 import diffusion_net as dn
    nz, nxemb, T, B = (int(v) for v in g["cfg"])
    Q = dn._netQ_U(nc=3, nz=nz, nxemb=nxemb, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type=str(g["var_type"]), with_noise=bool(g["with_noise"]), dataset="cifar10")
    sd = synth.module_state_like(Q, prefix="Q.")
    Q.load_state_dict(sd)
    return Q.to(dev).eval(), sd, nz, nxemb, T, B


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("tf32", 3e-3), ("bf16", 3e-2)])
def test_gen_samples_matches_reference_algorithm(prec, tol, dev):
    """gen_samples (MCMC.py:119-128): z ~ randn (torch CPU generator, as the reference's torch.randn(...).cuda()),
    noise-free prior Langevin, x = G(z) -- against the oracle from the same initial draw."""
    from damc_b200 import MCMC
    nz, ngf, nc, bs, K, s = 128, 64, 3, 10, 12, 0.4
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd = synth.generator_state(layers, seed=3, gain=0.85), synth.ebm_state(nz, seed=3)
    G, E = _nets("cifar10", nz, ngf, nc, gsd, esd, dev)
    torch.manual_seed(4321)
    x = MCMC.gen_samples(bs, nz, E, G, K, s, False, precision=prec)
    torch.manual_seed(4321)
    z_init = torch.randn(bs, nz)
    zk = O.langevin_prior(z_init.double(), synth.ebm_list_from_state(esd, torch.float64), K, s, False)
    ref = O.gen_forward(synth.gen_list_from_state(gsd, layers, torch.float64), zk)
    assert tuple(x.shape) == (bs, nc, 32, 32) and x.is_cuda
    e = relmax(x, ref)
    print(f"gen_samples[{prec}] vs oracle: {e:.3e}")
    assert e < tol
    assert all(p.requires_grad for p in list(E.parameters()))


@pytest.mark.parametrize("name", ["damc_cifar10_T10", "damc_cifar10_T100", "damc_cifar10_small_T20"])
def test_amortizer_prior_branch_golden(name, dev):
    """Q(x=None, b, device) (diffusion_net.py:591-595): xemb = prior_emb(randn), z_T = randn, T reverse steps -- against
    the reference's own output for the same draws (golden z_prior_f32) and the fp64 oracle."""
    from damc_b200 import MCMC
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    Q, sd, nz, nxemb, T, B = _Q(g, dev)
    pz, zT = synth.det_normal("prior_z", (B, nz)), synth.det_normal("zT", (B, nz))
    noise = synth.det_normal("qnoise", (T - 1, B, nz))
    with injected_randn([pz, zT]):
        z = MCMC.damc_sample(Q, x=None, b=B, device=dev, noise=noise.to(dev))
    with torch.no_grad():
        assert relmax(Q.prior_emb(pz.to(dev)), g["xemb_prior"]) < 1e-5
    P64 = synth.denoiser_params_from_state(sd, True, 128, torch.float64)
    z64 = O.damc_sample(P64, torch.from_numpy(g["xemb_prior"]).double(), zT.double(), T, -5.1, 9.8, str(g["var_type"]),
                        bool(g["with_noise"]), noise.double())
    ref_err, err = relmax(g["z_prior_f32"], z64), relmax(z, z64)
    print(f"{name} prior branch: ours-vs-fp64 {err:.3e}   reference fp32-vs-fp64 {ref_err:.3e}")
    assert tuple(z.shape) == (B, nz)
    assert err < max(2.0 * ref_err, 1e-3), (err, ref_err)
    # the module's forward is the same call
    with injected_randn([pz, zT]):
        z2 = Q(x=None, b=B, device=dev, noise=noise.to(dev))
    assert torch.equal(z, z2)


def test_gen_samples_with_diffusion_prior_golden(dev):
    """gen_samples_with_diffusion_prior (MCMC.py:146-150): (G(z), z) with z = Q(x=None, b, device).  The 'small'-variance
    golden has with_noise = False, so the two randn draws determine the result."""
    from damc_b200 import MCMC
    g = np.load(os.path.join(GOLDEN, "damc_cifar10_small_T20.npz"), allow_pickle=True)
    Q, sd, nz, nxemb, T, B = _Q(g, dev)
    ngf, nc = 64, 3
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd = synth.generator_state(layers, seed=5, gain=0.85)
    G, _ = _nets("cifar10", nz, ngf, nc, gsd, synth.ebm_state(nz), dev)
    pz, zT = synth.det_normal("prior_z", (B, nz)), synth.det_normal("zT", (B, nz))
    with injected_randn([pz, zT]):
        x, z = MCMC.gen_samples_with_diffusion_prior(B, dev, Q, G, precision="fp32")
    P64 = synth.denoiser_params_from_state(sd, True, 128, torch.float64)
    z64 = O.damc_sample(P64, torch.from_numpy(g["xemb_prior"]).double(), zT.double(), T, -5.1, 9.8, "small", False)
    ref_err, err = relmax(g["z_prior_f32"], z64), relmax(z, z64)
    print(f"gen_samples_with_diffusion_prior: z ours-vs-fp64 {err:.3e}   reference fp32-vs-fp64 {ref_err:.3e}")
    assert err < max(2.0 * ref_err, 1e-3)
    xref = O.gen_forward(synth.gen_list_from_state(gsd, layers, torch.float64), z.double().cpu())
    assert tuple(x.shape) == (B, nc, 32, 32) and relmax(x, xref) < 1e-4


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("tf32", 3e-3), ("fp16", 3e-3), ("bf16", 3e-2)])
def test_x_hat_out_is_generator_output_of_last_pre_update_z(prec, tol, dev):
    """x_hat_out = G(z_{K-1}) (the x_hat the reference computes in its last iteration, MCMC.py:55): values against the
    oracle, on direct launches (injected noise) and bit-identical through the captured CUDA graph (Philox)."""
    from damc_b200 import MCMC
    nz, ngf, nc, B, K, sigma = 128, 64, 3, 6, 4, 0.3
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, K, sigma, seed=9, gain=0.0)
    G, E = _nets("cifar10", nz, ngf, nc, gsd, esd, dev)
    gen, ebm = synth.gen_list_from_state(gsd, layers, torch.float64), synth.ebm_list_from_state(esd, torch.float64)
    xh = torch.full((B, nc, 32, 32), float("nan"), device=dev)
    MCMC.sample_langevin_post_z_with_prior(z0.to(dev).clone().requires_grad_(True), x.to(dev), G, E, K, sigma, True, 0.1,
                                           noise=noise.to(dev), precision=prec, x_hat_out=xh)
    z_prev = O.langevin_posterior_analytic(z0.double(), x.double(), gen, ebm, K - 1, sigma, True, 0.1, noise.double())
    e = relmax(xh, O.gen_forward(gen, z_prev))
    print(f"x_hat_out[{prec}] vs oracle G(z_(K-1)): {e:.3e}")
    assert e < tol
    if prec == "fp32":
        return
    outs = []
    for _ in range(3):   # direct, capturing, replay
        xo = torch.zeros(B, nc, 32, 32, device=dev)
        zo = MCMC.sample_langevin_post_z_with_prior(z0.to(dev).clone().requires_grad_(True), x.to(dev), G, E, K, sigma, True,
                                                    0.1, seed=77, precision=prec, x_hat_out=xo)
        outs.append((zo.clone(), xo))
    for zo, xo in outs[1:]:
        assert torch.equal(zo, outs[0][0]) and torch.equal(xo, outs[0][1])


def test_posterior_graph_cache_survives_alternating_batches(dev):
    """train_gen_recon.py alternates 128 training chains (:203) and 500 evaluation chains (:331) and advances the seed every
    call: from the third sighting on, every call of either configuration must be a graph replay, with results equal to
    direct launches (DAMC_GRAPH cannot be toggled in-process, so the direct result comes from the very first call)."""
    from damc_b200 import MCMC, _lib
    nz, ngf, nc, sigma = 128, 64, 3, 0.3
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd, z0, x, _ = synth.synth_problem(layers, nz, 160, 1, sigma, seed=2, gain=0.0)
    G, E = _nets("cifar10", nz, ngf, nc, gsd, esd, dev)
    z0, x = z0.to(dev), x.to(dev)
    big = torch.empty(1, device=dev)  # grow the shared workspace first so its base is stable across both batch sizes
    MCMC.sample_langevin_post_z_with_prior(z0.clone().requires_grad_(True), x, G, E, 2, sigma, True, 0.1, seed=1, precision="bf16")
    gh = MCMC.pack_generator(G, "bf16")

    def run(B, K, seed, step0=0):
        return MCMC.sample_langevin_post_z_with_prior(z0[:B].clone().requires_grad_(True), x[:B].contiguous(), G, E, K, sigma,
                                                      True, 0.1, seed=seed, step0=step0, precision="bf16").clone()

    first = {96: run(96, 5, 10), 160: run(160, 3, 10)}          # direct launches
    run(96, 5, 11); run(160, 3, 11)                             # second sighting: captured
    r0 = _lib.lib().damc_graph_replays(gh.ptr)
    for it in range(3):
        a, b = run(96, 5, 10), run(160, 3, 10)
        assert torch.equal(a, first[96]) and torch.equal(b, first[160])
        run(96, 5, 20 + it, step0=7 * it)                       # new seed / step offset: still the same graph
    assert _lib.lib().damc_graph_replays(gh.ptr) - r0 == 9
    # step0 shifts the Philox stream exactly as for direct launches: steps [0,5) == steps [0,2) then [2,5)
    whole = run(96, 5, 33)
    part = MCMC.sample_langevin_post_z_with_prior(z0[:96].clone().requires_grad_(True), x[:96].contiguous(), G, E, 2, sigma,
                                                  True, 0.1, seed=33, precision="bf16")
    part = MCMC.sample_langevin_post_z_with_prior(part.clone().requires_grad_(True), x[:96].contiguous(), G, E, 3, sigma,
                                                  True, 0.1, seed=33, step0=2, precision="bf16")
    assert torch.equal(whole, part)
    del big


def test_repack_detects_data_updates_without_version_bumps(dev):
    """The reference updates Q_dummy through param.data.copy_ (train_gen_recon.py:258-261), which bumps no version counter.
    damc_repack hashes the source tensors on the device and re-packs only when they changed: a .data update must be seen,
    an unchanged model must give bit-identical results, and restoring the weights must restore the result."""
    from damc_b200 import MCMC
    nz, ngf, nc, B, K, sigma = 128, 64, 3, 9, 3, 0.3
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, K, sigma, seed=4, gain=0.0)
    G, E = _nets("cifar10", nz, ngf, nc, gsd, esd, dev)

    def run(prec):
        return MCMC.sample_langevin_post_z_with_prior(z0.to(dev).clone().requires_grad_(True), x.to(dev), G, E, K, sigma, True,
                                                      0.1, noise=noise.to(dev), precision=prec).clone()

    for prec in ("tf32", "bf16", "fp32"):
        a, a2 = run(prec), run(prec)
        assert torch.equal(a, a2)
        old = G.gen[2].weight.data.clone()
        G.gen[2].weight.data.copy_(old * 1.05)           # no version bump
        b = run(prec)
        assert not torch.equal(a, b)
        Gf, _ = _nets("cifar10", nz, ngf, nc, {k: v.clone() for k, v in G.state_dict().items()}, esd, dev)   # fresh pack
        bf = MCMC.sample_langevin_post_z_with_prior(z0.to(dev).clone().requires_grad_(True), x.to(dev), Gf, E, K, sigma, True,
                                                    0.1, noise=noise.to(dev), precision=prec)
        assert torch.equal(b, bf)
        G.gen[2].weight.data.copy_(old)
        assert torch.equal(run(prec), a)
        oldb = E.ebm[0].bias.data.clone()
        E.ebm[0].bias.data.add_(0.3)
        assert not torch.equal(run(prec), a)
        E.ebm[0].bias.data.copy_(oldb)
        assert torch.equal(run(prec), a)


def test_shape_mismatches_raise_before_any_kernel(dev):
    """Mis-shaped z / xemb / x_hat_out must raise (the reference would fail in reshape / matmul), never reach a kernel."""
    from damc_b200 import MCMC, diffusion_net as dn
    G = dn._netG("cifar10", 128, 64, 3).to(dev)
    with pytest.raises(RuntimeError, match="shape"):
        MCMC.generator_forward(G, torch.zeros(4, 100, device=dev))
    Q = dn._netQ_U(nc=3, nz=128, nxemb=256, ntemb=128, nif=64, diffusion_residual=True, n_interval=8, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10").to(dev).eval()
    with pytest.raises(RuntimeError, match="shape"):
        MCMC.damc_sample(Q, xemb=torch.zeros(4, 128, device=dev))
    with pytest.raises(RuntimeError, match="shape"):
        MCMC.damc_sample(Q, xemb=torch.zeros(4, 256, device=dev), z_init=torch.zeros(4, 64))
    with pytest.raises(RuntimeError, match="shape"):
        MCMC.denoiser_eps(Q, torch.zeros(4, 128, device=dev), 0.0, torch.zeros(3, 256, device=dev))
    with pytest.raises(RuntimeError):
        MCMC.sample_langevin_post_z_with_prior(torch.zeros(2, 128, device=dev, requires_grad=True),
                                               torch.zeros(2, 3, 32, 32, device=dev), G, None, 2, 0.3, True, 0.1,
                                               x_hat_out=torch.zeros(2, 3, 32, 32))   # wrong device


# ----------------------------------------------------------------------------------------------------------------------
def mmd2_unbiased(x, y, bw):
    def k(a, b):
        return torch.exp(-(torch.cdist(a, b) ** 2) / (2 * bw * bw))
    n, m = len(x), len(y)
    kxx, kyy, kxy = k(x, x), k(y, y), k(x, y)
    return float((kxx.sum() - kxx.diag().sum()) / (n * (n - 1)) + (kyy.sum() - kyy.diag().sum()) / (m * (m - 1))
                 - 2 * kxy.mean())


def test_posterior_mmd_and_energy_for_every_precision_at_tensor_core_width(dev):
    """north_star: "long-chain energy and MMD statistics must also agree with the reference" -- for the precisions that are
    benchmarked.  CIFAR-shaped generator at ngf = 64 (tensor-core granularity), trained-like weights, K = 40 noisy steps.
    The reference algorithm (oracle, torch RNG, true fp32: TF32 off) runs on the GPU as the checker: at this width the CPU
    would need minutes per run.  Null distribution = oracle-vs-oracle with different seeds."""
    from damc_b200 import MCMC
    nz, ngf, nc, B, K, sigma, s = 128, 64, 3, 384, 40, 0.3, 0.1
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd, z0, x, _ = synth.synth_problem(layers, nz, B, 1, sigma, seed=6, gain=0.85)
    G, E = _nets("cifar10", nz, ngf, nc, gsd, esd, dev)
    gen = [(w.to(dev), b.to(dev), st, p) for w, b, st, p in synth.gen_list_from_state(gsd, layers)]
    ebm = [(w.to(dev), b.to(dev)) for w, b in synth.ebm_list_from_state(esd)]
    z0d, xd = z0.to(dev), x.to(dev)

    def oracle_run(seed):
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        return O.langevin_posterior(z0d, xd, gen, ebm, K, sigma, True, s)

    refs = [oracle_run(sd) for sd in (21, 22, 23, 24)]
    zz = torch.cat([refs[0], refs[1]])
    d = torch.cdist(zz, zz)
    bw = float(d[d > 0].median())
    null = [mmd2_unbiased(refs[i], refs[j], bw) for i in range(4) for j in range(i + 1, 4)]

    def U(z):
        with torch.no_grad():
            xh = O.gen_forward(gen, z)
            return (((xh - xd) ** 2).sum((1, 2, 3)) / (2 * sigma ** 2) + O.ebm_forward(ebm, z) + 0.5 * (z ** 2).sum(1)).mean().item()

    u_ref = np.array([U(r) for r in refs])
    u0 = U(z0d)
    for prec in ("tf32", "bf16", "fp16", "fp32"):
        zs = [MCMC.sample_langevin_post_z_with_prior(z0d.clone().requires_grad_(True), xd, G, E, K, sigma, True, s, seed=sd,
                                                     precision=prec) for sd in (201, 202)]
        stat = [mmd2_unbiased(o, r, bw) for o in zs for r in refs]
        u_our = np.array([U(o) for o in zs])
        print(f"posterior[{prec}] MMD^2: null max {max(null):.3e} mean {np.mean(null):.3e}; ours max {max(stat):.3e} mean "
              f"{np.mean(stat):.3e}; U start {u0:.1f} ref {u_ref.round(2)} ours {u_our.round(2)}")
        assert max(stat) < max(null) + 3 * (np.std(null) + 1e-5), (prec, stat, null)
        assert abs(u_our.mean() - u_ref.mean()) < 4 * u_ref.std() + 0.01 * abs(u_ref.mean()), (prec, u_our, u_ref)
        assert u_our.mean() < u0   # the chains did descend from the initial energy
