"""Import the UNMODIFIED reference read-only from /root/reference -- TEST INFRASTRUCTURE, build container only.

/root/reference does not exist on the GPU box, so nothing that runs there may call this; it is used by
``oracle/make_golden.py`` (fixture generation) and by CPU tests that skip when the reference is absent.
``src/MCMC.py:8`` imports ``pytorch_fid_wrapper`` (not installed) at module load -> an empty stub is registered.
"""
import contextlib
import importlib
import importlib.util
import io
import os
import sys
import types

REF_ROOT = "/root/reference/workspace"


def available():
    return os.path.isdir(REF_ROOT)


def load():
    """Returns (MCMC module, diffusion_net module) of the reference."""
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    sys.modules.setdefault("pytorch_fid_wrapper", types.ModuleType("pytorch_fid_wrapper"))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("src.MCMC"), importlib.import_module("src.diffusion_net")


def load_toy_net():
    """toy_example/src/diffusion_net.py under a private package name (it clashes with ``src``)."""
    root = os.path.join(REF_ROOT, "toy_example", "src")
    if "ref_toy_src" not in sys.modules:
        pkg = types.ModuleType("ref_toy_src")
        pkg.__path__ = [root]
        sys.modules["ref_toy_src"] = pkg
    return importlib.import_module("ref_toy_src.diffusion_net")


@contextlib.contextmanager
def injected_noise(draws):
    """Replace torch.randn / torch.randn_like with an iterator over pre-drawn tensors (the reference calls them
    through the module-global ``torch``: MCMC.py:38,64; diffusion_net.py:593,595,616)."""
    import torch
    it = iter(draws)
    real_randn, real_like = torch.randn, torch.randn_like

    def fake_randn(*a, **k):
        return next(it).clone()

    def fake_like(t, **k):
        return next(it).clone().to(t.dtype)

    torch.randn, torch.randn_like = fake_randn, fake_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = real_randn, real_like


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
