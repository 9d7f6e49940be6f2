"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/damc.h declares, and the
host-side mirror validates its inputs without touching a GPU (no compute calls here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    from damc_b200 import _lib
    return _lib


def test_header_symbols_exported(built_lib):
    header = open(os.path.join(ROOT, "include", "damc.h")).read()
    declared = set(re.findall(r"\b(damc_[a-z0-9_]+)\s*\(", header))
    declared -= {"damc_handle"}
    assert len(declared) >= 15
    h = ctypes.CDLL(built_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(h, name), f"{name} declared in damc.h but not exported"
    assert declared == set(built_lib.SIGNATURES), declared ^ set(built_lib.SIGNATURES)
    assert built_lib.lib().damc_version() >= 100


def test_no_cpu_fallback(built_lib):
    from damc_b200 import MCMC, diffusion_net as dn
    E = dn._netE(nz=16)
    z = torch.zeros(4, 16, requires_grad=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        MCMC.sample_langevin_prior_z(z, E, 3, 0.1, True)
    G = dn._netG_cifar10(nz=16, ngf=16)
    with pytest.raises(RuntimeError, match="CUDA"):
        MCMC.sample_langevin_post_z_with_prior(z, torch.zeros(4, 3, 32, 32), G, E, 3, 0.1, True, 0.1)


def test_structure_validation(built_lib):
    from damc_b200 import MCMC, diffusion_net as dn
    E = dn._netE(nz=16)
    E.ebm[1] = torch.nn.ReLU()
    with pytest.raises(RuntimeError, match="LeakyReLU"):
        MCMC.pack_ebm(E)
    E2 = dn._netE(nz=16)
    E2.ebm[0] = torch.nn.utils.spectral_norm(E2.ebm[0])
    with pytest.raises(RuntimeError, match="spectral"):
        MCMC.pack_ebm(E2)
    G = dn._netG_svhn(nz=16, ngf=16)
    G.gen[7] = torch.nn.Sigmoid()
    with pytest.raises(RuntimeError, match="Tanh"):
        MCMC.pack_generator(G)


def test_set_requires_grad_semantics(built_lib):
    from damc_b200 import MCMC, diffusion_net as dn
    E, G = dn._netE(nz=8), dn._netG_mnist(nz=8, ngf=16)
    MCMC.set_requires_grad([E, None, G], False)
    assert not any(p.requires_grad for p in list(E.parameters()) + list(G.parameters()))
    MCMC.set_requires_grad(E, True)
    assert all(p.requires_grad for p in E.parameters())


def test_state_dict_keys_match_reference_layout():
    from damc_b200 import diffusion_net as dn
    G = dn._netG_cifar10()
    assert list(G.state_dict())[:2] == ["gen.0.weight", "gen.0.bias"]
    assert tuple(G.gen[0].weight.shape) == (128, 1024, 8, 8) and tuple(G.gen[6].weight.shape) == (256, 3, 3, 3)
    E = dn._netE()
    assert set(E.state_dict()) == {f"ebm.{i}.{w}" for i in (0, 2, 4) for w in ("weight", "bias")}
    assert sum(p.numel() for p in E.parameters()) == 66201
    Q = dn._netQ_U(nxemb=1024, n_interval=100, dataset="cifar10")
    keys = set(Q.state_dict())
    for k in ("encoder.net.0.weight", "encoder.net.1.weight", "encoder.net.12.weight", "p.time_mlp.1.weight", "p.B",
              "p.in_layers.0._layer.0.weight", "p.out_layers.2._hyper_bias.weight", "p.mid_layers.0._layer_ctx.1.bias",
              "xemb", "prior_emb.2.weight"):
        assert k in keys, k
    assert sum(p.numel() for p in Q.p.parameters()) == 3143424


def test_encoder_family_detection(built_lib):
    """Host-side structure check that decides whether Q.encoder runs on the library (even-sized maps down to the final
    k x k) or stays in torch (the 28x28 MNIST encoder)."""
    from damc_b200 import MCMC, diffusion_net as dn
    assert MCMC._encoder_on_library(dn.Encoder("cifar10", nc=3, nemb=1024, nif=64), torch.zeros(1, 3, 32, 32))
    assert MCMC._encoder_on_library(dn.Encoder("celeba64", nc=3, nemb=128, nif=64), torch.zeros(1, 3, 64, 64))
    assert MCMC._encoder_on_library(dn.Encoder("celebaHQ", nc=3, nemb=128, nif=64), torch.zeros(1, 3, 256, 256))
    assert MCMC._encoder_on_library(dn.Encoder("mnist", nc=1, nemb=128, nif=64), torch.zeros(1, 1, 28, 28))   # odd maps: padded planes
    assert not MCMC._encoder_on_library(dn.Encoder("cifar10", nc=3, nemb=128, nif=64), torch.zeros(1, 3, 48, 48))
    enc = dn.Encoder("cifar10", nc=3, nemb=128, nif=64)
    enc.net[1] = torch.nn.BatchNorm2d(64)
    assert not MCMC._encoder_on_library(enc, torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        MCMC.encoder_forward(dn.Encoder("cifar10", nc=3, nemb=128, nif=64), torch.zeros(1, 3, 32, 32), precision="bf16")


def test_library_selftest(built_lib):
    """Host-side check of the magic-number division behind the kernels' tile / row decoding (70 k divisors x 80 dividends
    up to 2^31 - 1, against integer division)."""
    built_lib.check(built_lib.lib().damc_selftest(), "damc_selftest")


def test_c_abi_argument_validation_without_a_gpu(built_lib):
    """Entry points reject bad handles / shapes with a status code and a message before any CUDA work (no GPU needed)."""
    import ctypes as C
    L = built_lib.lib()
    err = lambda: L.damc_last_error().decode()
    # null / wrong handles
    assert L.damc_denoise(None, None, None, 4, 10, None, 1, 1, None, 0, 0, 0, None, 0, None) == 1 and "handle" in err()
    assert L.damc_encoder_forward(None, None, None, 1, None, 0, None) == 1
    assert L.damc_posterior_langevin(None, None, None, None, 1, 1, 0.1, 0.1, 1, None, 0, 0, 0, None, None, None, 0, None) == 1
    assert L.damc_denoise_workspace_bytes(None, 4, 10, 1) == 0 and L.damc_encoder_workspace_bytes(None, 4) == 0
    # encoder shapes outside the supported family: UNSUPPORTED (2) with the offending layer named
    fake = C.c_void_p(0x1000)   # never dereferenced: shape validation comes first
    def layers(specs):
        arr = (built_lib.ConvLayer * len(specs))()
        for i, (cin, cout, k, s, p) in enumerate(specs):
            last = i == len(specs) - 1
            arr[i] = built_lib.ConvLayer(cin, cout, k, s, p, fake, fake, None if last else fake, None if last else fake)
        return arr
    out = C.c_void_p()
    bad_first = layers([(3, 64, 5, 1, 2), (64, 128, 4, 2, 1), (128, 128, 4, 1, 0)])
    assert L.damc_pack_encoder(C.byref(out), 3, bad_first, 8, 8, 0.2, 1e-5, 1, None) == 2 and "layer 0" in err()
    tiny = layers([(1, 64, 3, 1, 1), (64, 128, 4, 2, 1), (128, 256, 4, 2, 1), (256, 128, 1, 1, 0)])   # 2 -> 1 -> stride-2 conv of a 1 x 1 map
    assert L.damc_pack_encoder(C.byref(out), 4, tiny, 2, 2, 0.2, 1e-5, 1, None) == 2 and "2 x 2" in err()
    assert L.damc_pack_encoder(C.byref(out), 3, bad_first, 8, 8, 0.2, 1e-5, 7, None) == 1 and "precision" in err()
