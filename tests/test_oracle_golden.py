"""Pin the CPU oracle (oracle/damc_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import damc_oracle as O
from oracle import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = {"post_cifar10_full", "post_cifar10_full_k5", "post_celebaHQ_w64"}  # minutes of CPU: covered by the GPU parity tests instead


def relmax(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def _names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


@pytest.mark.parametrize("name", [n for n in _names("post_") if n not in FULL])
def test_posterior_oracle_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    nz, ngf, nc, B, K = (int(v) for v in g["cfg"])
    sigma, step, noise_on = float(g["sigma"]), float(g["step"]), bool(g["noise_on"])
    layers = synth.gen_layers(str(g["dataset"]), nz, ngf, nc)
    gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, K, sigma, gain=float(g["gain"]))
    for tag, dt, tol in (("f64", torch.float64, 1e-9), ("f32", torch.float32, 2e-3)):
        gen, ebm = synth.gen_list_from_state(gsd, layers, dt), synth.ebm_list_from_state(esd, dt)
        z = O.langevin_posterior(z0.to(dt), x.to(dt), gen, ebm, K, sigma, noise_on, step, noise.to(dt))
        assert relmax(z, g["z_" + tag]) < tol, (name, tag)
        if tag == "f64":  # analytic restatement == autograd restatement (this is what the kernels compute)
            za = O.langevin_posterior_analytic(z0.to(dt), x.to(dt), gen, ebm, K, sigma, noise_on, step, noise.to(dt))
            assert relmax(za, g["z_f64"]) < 1e-7, name
            xh = O.gen_forward(gen, z)[:, :, :4, :4]
            assert relmax(xh, g["xhat_f64"]) < 1e-9


@pytest.mark.parametrize("name", _names("prior_"))
def test_prior_oracle_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    nz, B, K = (int(v) for v in g["cfg"])
    step, noise_on = float(g["step"]), bool(g["noise_on"])
    esd = synth.ebm_state(nz)
    z0, noise = synth.det_normal("z0", (B, nz)), synth.det_normal("noise", (K, B, nz))
    for tag, dt, tol in (("f64", torch.float64, 1e-10), ("f32", torch.float32, 1e-3)):
        ebm = synth.ebm_list_from_state(esd, dt)
        trace = []
        z = O.langevin_prior(z0.to(dt), ebm, K, step, noise_on, noise.to(dt), trace)
        assert relmax(z, g["z_" + tag]) < tol, (name, tag)
        assert relmax(O.ebm_forward(ebm, z), g["en_" + tag]) < max(tol, 1e-6) * 10
        if tag == "f64":
            za = O.langevin_prior_analytic(z0.to(dt), ebm, K, step, noise_on, noise.to(dt))
            assert relmax(za, g["z_f64"]) < 1e-7
            # the logged scalars are the ones the reference prints (MCMC.py:40-41)
            log = str(g["log_f64"])
            for i, en, zn in trace:
                assert "{}/{:.3f}/{:.3f}".format(i, en, zn) in log


@pytest.mark.parametrize("name", _names("damc_"))
def test_damc_oracle_matches_reference(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=True)
    nz, nxemb, T, B = (int(v) for v in g["cfg"])
    var_type, with_noise = str(g["var_type"]), bool(g["with_noise"])
    from damc_b200 import diffusion_net as dn
    Q = dn._netQ_U(nc=3, nz=nz, nxemb=nxemb, ntemb=128, nif=64, diffusion_residual=True, n_interval=T,
                   logsnr_min=-5.1, logsnr_max=9.8, var_type=var_type, with_noise=with_noise, dataset="cifar10")
    sd = synth.module_state_like(Q, prefix="Q.")
    Q.load_state_dict(sd)  # same key set as the reference's Q: the mirror is checkpoint-compatible
    P = synth.denoiser_params_from_state(sd, True, 128)
    zT, noise = synth.det_normal("zT", (B, nz)), synth.det_normal("qnoise", (T - 1, B, nz))
    xemb = torch.from_numpy(g["xemb"])
    # per-step eps prediction: tight
    for j, i in enumerate((T - 1, T // 2, 1, 0)):
        lt = O.logsnr_schedule(torch.ones(B) * float(i) / (T - 1.0), -5.1, 9.8)
        eps = O.denoiser_eps(P, zT * (0.3 + 0.2 * i / T), lt, xemb)
        assert relmax(eps, g["eps_steps"][j]) < 1e-5, (name, i)
    # encoder mirror reproduces the reference's xemb
    x = torch.tanh(synth.det_normal("x", (B, 3, 32, 32)))
    with torch.no_grad():
        assert relmax(Q.encoder(x), g["xemb"]) < 1e-5
        assert relmax(Q.prior_emb(synth.det_normal("prior_z", (B, nz))), g["xemb_prior"]) < 1e-5
    # end to end: fp32 chaotic amplification -> compare at the reference's own fp32 noise level
    z = O.damc_sample(P, xemb, zT, T, -5.1, 9.8, var_type, with_noise, noise)
    P64 = synth.denoiser_params_from_state(sd, True, 128, torch.float64)
    z64 = O.damc_sample(P64, xemb.double(), zT.double(), T, -5.1, 9.8, var_type, with_noise, noise.double())
    ref_err = relmax(g["z_x_f32"], z64)
    assert relmax(z, z64) < max(2.0 * ref_err, 1e-4), (relmax(z, z64), ref_err)
    zp = O.damc_sample(P, torch.from_numpy(g["xemb_prior"]), zT, T, -5.1, 9.8, var_type, with_noise, noise)
    zp64 = O.damc_sample(P64, torch.from_numpy(g["xemb_prior"]).double(), zT.double(), T, -5.1, 9.8, var_type,
                         with_noise, noise.double())
    assert relmax(zp, zp64) < max(2.0 * relmax(g["z_prior_f32"], zp64), 1e-4)


def test_toy_oracle_matches_reference():
    g = np.load(os.path.join(GOLDEN, "toy.npz"), allow_pickle=True)
    B, K = (int(v) for v in g["cfg"])
    dims = [(128, 2), (128, 128), (128, 128), (2, 128)]
    z0, x, noise = synth.det_normal("toy.z0", (B, 2)), synth.det_normal("toy.x", (B, 2)), \
        synth.det_normal("toy.noise", (K, B, 2))
    for tag, dt, tol in (("f64", torch.float64, 1e-10), ("f32", torch.float32, 1e-3)):
        mlp = [(synth.det_normal(f"toy.w{i}", d, 0, 0.2).to(dt), synth.det_normal(f"toy.b{i}", (d[0],), 0, 0.1).to(dt))
               for i, d in enumerate(dims)]
        z = O.toy_langevin_posterior(z0.to(dt), x.to(dt), mlp, K, True, float(g["step"]), noise.to(dt))
        assert relmax(z, g["z_" + tag]) < tol, tag


def test_calculate_loss_mirror_matches_reference():
    """Training-side mirror (_netQ_U.calculate_loss) against the reference with injected randn / rand draws."""
    from damc_b200 import diffusion_net as dn
    g = np.load(os.path.join(GOLDEN, "qloss_cifar10.npz"))
    B, nz, nxemb, T = (int(v) for v in g["cfg"])
    Q = dn._netQ_U(nc=3, nz=nz, nxemb=nxemb, ntemb=128, nif=64, diffusion_residual=True, n_interval=T,
                   logsnr_min=-5.1, logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10")
    Q.load_state_dict(synth.module_state_like(Q, prefix="Q."))
    Q.eval()
    x = torch.tanh(synth.det_normal("x", (B, 3, 32, 32)))
    z, mask = synth.det_normal("lz", (B, nz)), (synth.det_normal("lmask", (B, 1)) > 0).float()
    draws = iter([synth.det_normal("prior_z", (B, nz)), synth.det_normal("leps", (B, nz))])
    u = torch.sigmoid(synth.det_normal("lu", (B,)))
    real = (torch.randn, torch.randn_like, torch.rand)
    torch.randn = lambda *a, **k: next(draws).clone()
    torch.randn_like = lambda t, **k: next(draws).clone()
    torch.rand = lambda *a, **k: u.clone()
    try:
        with torch.no_grad():
            loss = Q.calculate_loss(x=x, z=z, mask=mask)
    finally:
        torch.randn, torch.randn_like, torch.rand = real
    assert relmax(loss, g["loss"]) < 1e-5
