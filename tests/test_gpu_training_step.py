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
