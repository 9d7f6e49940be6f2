"""Out-of-bounds guard bands (compute-sanitizer is not available on the GPU pool): every library call that takes a caller
workspace runs with the workspace embedded in a larger buffer filled with a canary pattern; nothing outside the
``*_workspace_bytes`` range -- and nothing outside the z / x_hat / xemb outputs -- may change."""
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu
GUARD = 1 << 20


@pytest.fixture()
def guarded(monkeypatch):
    from damc_b200 import MCMC
    dev = torch.device("cuda:0")
    state = {}

    def ws(device, nbytes):
        buf = torch.full((int(nbytes) + 2 * GUARD,), 0xA5, dtype=torch.uint8, device=device)
        state["buf"], state["n"] = buf, int(nbytes)
        return buf[GUARD:GUARD + int(nbytes)]

    monkeypatch.setattr(MCMC, "_workspace", ws)

    def check():
        torch.cuda.synchronize()
        buf, n = state["buf"], state["n"]
        assert bool((buf[:GUARD] == 0xA5).all()), "write below the workspace"
        assert bool((buf[GUARD + n:] == 0xA5).all()), "write above the workspace"

    return dev, check


def _padded(t, pad=4096):
    """t's values inside a larger canary-filled allocation; returns (view, checker)."""
    flat = torch.full((t.numel() + 2 * pad,), 7.25, dtype=t.dtype, device=t.device)
    flat[pad:pad + t.numel()] = t.flatten()
    view = flat[pad:pad + t.numel()].view(t.shape)

    def check():
        torch.cuda.synchronize()
        assert bool((flat[:pad] == 7.25).all()) and bool((flat[pad + t.numel():] == 7.25).all()), "write outside the tensor"

    return view, check


@pytest.mark.parametrize("dataset,nz,ngf,nc,B", [("cifar10", 128, 64, 3, 37), ("svhn", 100, 64, 3, 130), ("mnist", 8, 64, 1, 5),
                                                 ("celeba64", 100, 64, 3, 3)])
@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
def test_posterior_and_generator_stay_inside_their_buffers(dataset, nz, ngf, nc, B, prec, guarded):
    from damc_b200 import MCMC, diffusion_net as dn
    dev, check = guarded
    layers = synth.gen_layers(dataset, nz, ngf, nc)
    gsd, esd, z0, x, noise = synth.synth_problem(layers, nz, B, 2, 0.3, seed=3, gain=0.0)
    G, E = dn._netG(dataset, nz, ngf, nc), dn._netE(nz)
    G.load_state_dict(gsd)
    E.load_state_dict(esd)
    G, E = G.to(dev), E.to(dev)
    z, zc = _padded(z0.to(dev))
    xh, xc = _padded(torch.zeros_like(x).to(dev))
    out = MCMC.sample_langevin_post_z_with_prior(z.requires_grad_(True), x.to(dev), G, E, 2, 0.3, True, 0.1, seed=1,
                                                 precision=prec, x_hat_out=xh)
    check(); zc(); xc()
    assert torch.isfinite(out).all() and torch.isfinite(xh).all()
    MCMC.generator_forward(G, z0.to(dev), precision=prec)
    check()


@pytest.mark.parametrize("B", [1, 130, 2100])
@pytest.mark.parametrize("prec", ["fp32", "bf16", "fp16"])
@pytest.mark.parametrize("schedule", ["hoisted", "round1"])
def test_denoiser_and_encoder_stay_inside_their_buffers(B, prec, schedule, guarded, monkeypatch):
    from damc_b200 import MCMC, diffusion_net as dn
    dev, check = guarded
    if schedule == "round1":   # cluster kernel / per-layer launches + graph instead of the hoisted-context schedule (16-bit modes)
        if prec == "fp32":
            pytest.skip("the fp32 kernel has one schedule")
        monkeypatch.setenv("DAMC_DEN_SEQ", "0")
    T = 6
    Q = dn._netQ_U(nc=3, nz=128, nxemb=256, ntemb=128, nif=64, diffusion_residual=True, n_interval=T, logsnr_min=-5.1,
                   logsnr_max=9.8, var_type="large", with_noise=True, dataset="cifar10")
    Q.load_state_dict(synth.module_state_like(Q, prefix="Q."))
    Q = Q.to(dev).eval()
    xemb = (0.5 * synth.det_normal("xe", (B, 256))).to(dev)
    zT, zc = _padded(synth.det_normal("zT", (B, 128)).to(dev))
    for _ in range(3):   # direct launches, graph capture, graph replay (per-layer tcgen05 path) / cluster kernel (B >= 2048)
        out = MCMC.damc_sample(Q, xemb=xemb, z_init=zT, seed=5, precision=prec)
        check(); zc()
    assert torch.isfinite(out).all()
    eps = MCMC.denoiser_eps(Q, zT, 0.3, xemb, precision=prec)
    check()
    assert torch.isfinite(eps).all()
    if B <= 130:
        x = torch.tanh(synth.det_normal("xg", (B, 3, 32, 32))).to(dev)
        xe = MCMC.encoder_forward(Q.encoder, x, precision=prec)
        check()
        assert torch.isfinite(xe).all()
