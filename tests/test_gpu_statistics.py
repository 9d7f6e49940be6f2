"""Statistical parity of the Philox-driven samplers with the reference algorithm (CPU oracle, torch RNG): long-chain
energy traces and an MMD two-sample test (north_star "long-chain energy and MMD statistics").

The reference has no MMD code, so the test defines it: RBF kernel, median-heuristic bandwidth, unbiased MMD^2; the null
distribution comes from oracle-vs-oracle runs with different seeds, and the GPU-vs-oracle statistic must not exceed the
largest null value by more than a small margin."""
import numpy as np
import pytest
import torch

from oracle import damc_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu


def mmd2_unbiased(x, y, bw):
    def k(a, b):
        d = torch.cdist(a, b) ** 2
        return torch.exp(-d / (2 * bw * bw))
    n, m = len(x), len(y)
    kxx, kyy, kxy = k(x, x), k(y, y), k(x, y)
    return float((kxx.sum() - kxx.diag().sum()) / (n * (n - 1)) + (kyy.sum() - kyy.diag().sum()) / (m * (m - 1))
                 - 2 * kxy.mean())


def median_bw(x, y):
    z = torch.cat([x, y])
    d = torch.cdist(z, z)
    return float(d[d > 0].median())


@pytest.fixture(scope="module")
def dev():
    import __graft_entry__ as ge
    ge.build()
    return torch.device("cuda:0")


def test_prior_langevin_energy_trace_and_mmd(dev):
    from damc_b200 import MCMC, diffusion_net as dn
    nz, B, K, s = 128, 768, 100, 0.4
    esd = synth.ebm_state(nz, seed=4)
    E = dn._netE(nz)
    E.load_state_dict(esd)
    E = E.to(dev)
    ebm = synth.ebm_list_from_state(esd, torch.float64)
    z0 = synth.det_normal("stat.z0", (B, nz)).double()

    def oracle_run(seed):
        torch.manual_seed(seed)
        tr = []
        z = O.langevin_prior(z0, ebm, K, s, True, trace=tr)
        return z.float(), np.array([[t[1], t[2]] for t in tr])

    refs = [oracle_run(sd) for sd in (11, 12, 13, 14)]
    # ours, Philox noise; trace via verbose capture is exercised in the parity tests -> read energies directly here
    import ctypes as C
    outs = []
    for sd in (101, 102):
        z = z0.float().to(dev).clone().requires_grad_(True)
        outs.append(MCMC.sample_langevin_prior_z(z, E, K, s, True, seed=sd).cpu())
    bw = median_bw(refs[0][0], refs[1][0])
    null = [mmd2_unbiased(refs[i][0], refs[j][0], bw) for i in range(4) for j in range(i + 1, 4)]
    stat = [mmd2_unbiased(o, refs[i][0], bw) for o in outs for i in range(4)]
    print(f"prior MMD^2: null max {max(null):.3e} mean {np.mean(null):.3e}; ours-vs-oracle max {max(stat):.3e} mean {np.mean(stat):.3e}")
    assert max(stat) < max(null) + 3 * (np.std(null) + 1e-5), (stat, null)
    # energy statistics of the final samples agree (mean E and mean |z|^2/2 per chain)
    e_ref = np.array([O.ebm_forward(ebm, r[0].double()).mean().item() for r in refs])
    e_our = np.array([O.ebm_forward(ebm, o.double()).mean().item() for o in outs])
    n_ref = np.array([0.5 * (r[0].double() ** 2).sum(1).mean().item() for r in refs])
    n_our = np.array([0.5 * (o.double() ** 2).sum(1).mean().item() for o in outs])
    print("mean E ref", e_ref, "ours", e_our, " mean |z|^2/2 ref", n_ref, "ours", n_our)
    assert abs(e_our.mean() - e_ref.mean()) < 4 * e_ref.std() + 0.02 * abs(e_ref.mean()) + 0.02
    assert abs(n_our.mean() - n_ref.mean()) < 4 * n_ref.std() + 0.01 * n_ref.mean()


def test_posterior_langevin_mmd_and_energy(dev):
    """Posterior chains on a small CIFAR-shaped generator: z_K distribution (per fixed x) pooled over chains."""
    from damc_b200 import MCMC, diffusion_net as dn
    nz, ngf, nc, B, K, sigma, s = 128, 16, 3, 384, 40, 0.3, 0.1
    layers = synth.gen_layers("cifar10", nz, ngf, nc)
    gsd, esd, z0, x, _ = synth.synth_problem(layers, nz, B, 1, sigma, seed=6, gain=0.85)
    G, E = dn._netG("cifar10", nz, ngf, nc), dn._netE(nz)
    G.load_state_dict(gsd)
    E.load_state_dict(esd)
    G, E = G.to(dev), E.to(dev)
    gen, ebm = synth.gen_list_from_state(gsd, layers), synth.ebm_list_from_state(esd)

    def oracle_run(seed):
        torch.manual_seed(seed)
        return O.langevin_posterior(z0, x, gen, ebm, K, sigma, True, s)

    refs = [oracle_run(sd) for sd in (21, 22, 23)]
    outs = {}
    for prec in ("fp32", "bf16"):
        if prec == "bf16":
            continue  # ngf = 16 is below the tensor-core engine's 64-channel granularity; bf16 is covered at full width
        outs[prec] = [MCMC.sample_langevin_post_z_with_prior(z0.to(dev).clone().requires_grad_(True), x.to(dev), G, E, K,
                                                             sigma, True, s, seed=sd, precision=prec).cpu()
                      for sd in (201, 202)]
    bw = median_bw(refs[0], refs[1])
    null = [mmd2_unbiased(refs[i], refs[j], bw) for i in range(3) for j in range(i + 1, 3)]

    def U(z):
        xh = O.gen_forward(gen, z)
        return (((xh - x) ** 2).sum((1, 2, 3)) / (2 * sigma ** 2) + O.ebm_forward(ebm, z) + 0.5 * (z ** 2).sum(1)).mean().item()

    u_ref = np.array([U(r) for r in refs])
    for prec, zs in outs.items():
        stat = [mmd2_unbiased(o, r, bw) for o in zs for r in refs]
        u_our = np.array([U(o) for o in zs])
        print(f"posterior[{prec}] MMD^2: null max {max(null):.3e}; ours max {max(stat):.3e}; U ref {u_ref} ours {u_our}")
        assert max(stat) < max(null) + 3 * (np.std(null) + 1e-5), (prec, stat, null)
        assert abs(u_our.mean() - u_ref.mean()) < 4 * u_ref.std() + 0.01 * abs(u_ref.mean())
