"""Drop-in mirror of the reference's sampler surface (workspace/src/MCMC.py and _netQ_U.forward), backed by
libdamc_b200 (hand-written sm_100a CUDA behind the C ABI of include/damc.h).

Same names, argument order and side effects as the reference:
  sample_langevin_prior_z            src/MCMC.py:27-46
  sample_langevin_post_z_with_prior  src/MCMC.py:48-74
  gen_samples                        src/MCMC.py:119-128
  gen_samples_with_diffusion_prior   src/MCMC.py:146-150
  set_requires_grad                  src/MCMC.py:12-25
  damc_sample  (= _netQ_U.forward)   src/diffusion_net.py:585-622
Extra keyword-only arguments (noise=, chain0=, precision=) are additions; defaults reproduce the reference call.
There is no CPU or PyTorch fallback: tensors must be CUDA fp32, and a missing library raises.
"""
import ctypes as C
import math
import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, check

_PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16, "tf32": _lib.PREC_TF32}
_DEN_PRECISIONS = {k: v for k, v in _PRECISIONS.items() if k != "tf32"}   # denoiser / encoder GEMMs: no tf32 variant
# 'tf32' is what the unmodified reference computes on a GPU: torch runs its (transposed) convolutions on the tensor cores
# in TF32 by default (torch.backends.cudnn.allow_tf32 = True; reference src/MCMC.py:55-60 goes through cuDNN).
DEFAULT_PRECISION = "tf32"
DEFAULT_DENOISER_PRECISION = "fp32"


def set_default_precision(name):
    """Generator GEMM arithmetic: 'tf32' (default: tcgen05 tensor cores on tf32-rounded fp32 tensors, rel 1e-3 parity),
    'bf16' (tcgen05, 16-bit operands, rel 2e-2, 2x the tf32 throughput), 'fp16' (same engine as bf16 with fp16 operands:
    ~8x smaller operand rounding at the same speed) or 'fp32' (CUDA-core fp32 FMA GEMMs: exact-fp32 cross-check)."""
    global DEFAULT_PRECISION
    if name not in _PRECISIONS:
        raise ValueError(f"precision must be one of {list(_PRECISIONS)}")
    DEFAULT_PRECISION = name


def set_default_denoiser_precision(name):
    """Arithmetic of the DAMC denoiser GEMMs: 'fp32' (one persistent CUDA-core kernel for all T steps; parity mode) or
    'bf16' / 'fp16' (per-layer tcgen05 GEMMs; the throughput mode for large batches of chains)."""
    global DEFAULT_DENOISER_PRECISION
    if name not in _DEN_PRECISIONS:
        raise ValueError(f"precision must be one of {list(_DEN_PRECISIONS)}")
    DEFAULT_DENOISER_PRECISION = name


def set_requires_grad(nets, requires_grad=False):
    """Toggle requires_grad on every parameter of a module or list of modules (None entries are skipped)."""
    for net in nets if isinstance(nets, list) else [nets]:
        if net is not None:
            for p in net.parameters():
                p.requires_grad = requires_grad


# ----------------------------------------------------------------------------------------------------------------------
# module -> packed-weight handle (validated, cached on parameter identity + version)
# ----------------------------------------------------------------------------------------------------------------------
class _Handle:
    def __init__(self, ptr):
        self.ptr = ptr
        self._fin = weakref.finalize(self, lib().damc_free, ptr)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32_cuda(t, what, strict=False):
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor (damc_b200 has no CPU fallback); got device {t.device}")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what} must be float32, got {t.dtype}")
    if strict and not t.is_contiguous():  # packed handles keep the pointer and re-read it on every call
        raise RuntimeError(f"{what} must be contiguous")
    return t.detach().contiguous()


def _no_hooks(m, what):
    if hasattr(m, "weight_orig") or hasattr(m, "parametrizations"):
        raise RuntimeError(f"{what}: spectral-norm / parametrized layers are not supported by the packed kernels")


def _key(params):
    return tuple((p.data_ptr(), tuple(p.shape), p.dtype, p.is_contiguous()) for p in params)


_cache = weakref.WeakKeyDictionary()  # module -> {tag: (key, handle, keepalive)}


def _cached(module, tag, params, build):
    """Packed-weight handle of a module.  Allocations are cached on the identity (storage pointers, shapes) of the
    parameters; their VALUES are re-read on every call (damc_repack: a few async copy/transpose kernels), because
    ``param.data.copy_()`` style updates -- e.g. the reference's EMA of Q_dummy, train_gen_recon.py:258-261 -- do not
    bump tensor version counters and would otherwise leave stale packed weights."""
    slot = _cache.setdefault(module, {})
    key = _key(params)
    hit = slot.get(tag)
    if hit is not None and hit[0] == key:
        check(lib().damc_repack(hit[1].ptr, _stream(params[0].device)), "damc_repack")
        return hit[1]
    handle, keep = build()
    slot[tag] = (key, handle, keep)
    return handle


def _slope_of(act, what):
    if not isinstance(act, nn.LeakyReLU):
        raise RuntimeError(f"{what}: expected nn.LeakyReLU, found {type(act).__name__}")
    return float(act.negative_slope)


def pack_ebm(netE):
    """netE.ebm = Sequential(Linear, LeakyReLU, Linear, LeakyReLU, Linear(ndf,1))  (reference diffusion_net.py:212-220)."""
    seq = netE.ebm
    if len(seq) != 5 or not all(isinstance(seq[i], nn.Linear) for i in (0, 2, 4)):
        raise RuntimeError("netE.ebm must be Sequential(Linear, LeakyReLU, Linear, LeakyReLU, Linear)")
    s1, s2 = _slope_of(seq[1], "netE.ebm[1]"), _slope_of(seq[3], "netE.ebm[3]")
    if s1 != s2:
        raise RuntimeError("netE.ebm: both LeakyReLU slopes must be equal")
    lins = [seq[0], seq[2], seq[4]]
    for i, l in enumerate(lins):
        _no_hooks(l, f"netE.ebm[{2 * i}]")
        if l.bias is None:
            raise RuntimeError("netE.ebm: Linear layers need a bias")
    if lins[2].out_features != 1 or lins[1].in_features != lins[0].out_features or lins[1].out_features != lins[2].in_features:
        raise RuntimeError("netE.ebm: expected nz -> ndf -> ndf -> 1")
    params = [p for l in lins for p in (l.weight, l.bias)]

    def build():
        ts = [_f32_cuda(p, "netE parameter", True) for p in params]
        out = C.c_void_p()
        check(lib().damc_pack_mlp(C.byref(out), lins[0].in_features, lins[0].out_features,
                                  *[C.c_void_p(t.data_ptr()) for t in ts], s1, _stream(ts[0].device)), "damc_pack_mlp")
        return _Handle(out), ts

    return _cached(netE, "ebm", params, build)


def pack_generator(netG, precision=None):
    """netG.gen = Sequential(ConvTranspose2d, LeakyReLU, ..., ConvTranspose2d, Tanh)  (reference diffusion_net.py:26-45)."""
    precision = precision or DEFAULT_PRECISION
    seq = netG.gen
    if len(seq) % 2 or len(seq) < 4 or not isinstance(seq[len(seq) - 1], nn.Tanh):
        raise RuntimeError("netG.gen must alternate ConvTranspose2d / LeakyReLU and end with Tanh")
    convs, slope = [], None
    for i in range(0, len(seq), 2):
        c = seq[i]
        if not isinstance(c, nn.ConvTranspose2d):
            raise RuntimeError(f"netG.gen[{i}]: expected ConvTranspose2d, found {type(c).__name__}")
        _no_hooks(c, f"netG.gen[{i}]")
        sq = lambda v: v[0] == v[1]
        if not (sq(c.kernel_size) and sq(c.stride) and sq(c.padding)) or c.groups != 1 or c.dilation != (1, 1) or \
                c.output_padding != (0, 0):
            raise RuntimeError(f"netG.gen[{i}]: only square, ungrouped, undilated ConvTranspose2d is supported")
        convs.append(c)
        if i + 1 < len(seq) - 1:
            s = _slope_of(seq[i + 1], f"netG.gen[{i + 1}]")
            if slope is not None and s != slope:
                raise RuntimeError("netG.gen: all LeakyReLU slopes must be equal")
            slope = s
    params = [p for c in convs for p in ([c.weight] + ([c.bias] if c.bias is not None else []))]

    def build():
        keep, arr = [], (_lib.ConvTLayer * len(convs))()
        for i, c in enumerate(convs):
            w = _f32_cuda(c.weight, "netG weight", True)
            b = _f32_cuda(c.bias, "netG bias", True) if c.bias is not None else None
            keep += [w, b]
            arr[i] = _lib.ConvTLayer(c.in_channels, c.out_channels, c.kernel_size[0], c.stride[0], c.padding[0],
                                     w.data_ptr(), b.data_ptr() if b is not None else None)
        out = C.c_void_p()
        check(lib().damc_pack_generator(C.byref(out), len(convs), arr, slope if slope is not None else 0.2,
                                        _PRECISIONS[precision], _stream(keep[0].device)), "damc_pack_generator")
        return _Handle(out), keep

    return _cached(netG, "gen_" + precision, params, build)


_workspaces = {}


def _workspace(device, nbytes):
    key = (device.type, device.index)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(int(nbytes), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _draw_seed():
    # host-side draw from torch's CPU generator: torch.manual_seed() controls it, and no device sync is needed
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _noise_ptr(noise, shape, device):
    if noise is None:
        return None, C.c_void_p(None)
    n = _f32_cuda(noise, "noise")
    if tuple(n.shape) != tuple(shape) or n.device != device:
        raise RuntimeError(f"noise must have shape {tuple(shape)} on {device}, got {tuple(n.shape)} on {n.device}")
    return n, C.c_void_p(n.data_ptr())


def _chain_tensor(z, what="z"):
    if not z.is_cuda or z.dtype != torch.float32 or z.dim() != 2:
        raise RuntimeError(f"{what} must be a 2-D CUDA float32 tensor [B, nz] (got {tuple(z.shape)}, {z.dtype}, {z.device})")
    if not z.data.is_contiguous():
        raise RuntimeError(f"{what} must be contiguous: the update is applied in place on {what}.data")
    return z.data


# ----------------------------------------------------------------------------------------------------------------------
# samplers
# ----------------------------------------------------------------------------------------------------------------------
def sample_langevin_prior_z(z, netE, e_l_steps, e_l_step_size, e_l_with_noise, verbose=False, *, noise=None,
                            seed=None, chain0=0, step0=0, precision=None):
    """K-step Langevin on E(z) + |z|^2/2, all steps in one persistent kernel.  Updates ``z.data`` in place and returns
    ``z.detach()`` like the reference (MCMC.py:36,46).  noise: optional injected normals [K,B,nz].
    precision: None / 'fp32' = the fp32 kernel with the MLP resident in shared memory (parity mode, best up to a few thousand
    chains); 'fp16' = the MLP's mat-mat products on the tensor cores, one CTA per 128 chains (damc_prior_langevin_tc: the
    large-batch form, 16-bit accuracy of dE/dz, no verbose trace)."""
    zd = _chain_tensor(z)
    B, nz = zd.shape
    h = pack_ebm(netE)
    set_requires_grad(netE, requires_grad=False)
    K = int(e_l_steps)
    keep, nptr = _noise_ptr(noise, (K, B, nz), zd.device)
    trace = torch.empty(K, 2, dtype=torch.float32, device=zd.device) if verbose and K > 0 else None
    if precision not in (None, "fp32", "fp16"):
        raise ValueError("prior sampler precision must be None / 'fp32' or 'fp16'")
    if precision == "fp16":
        if verbose:
            raise RuntimeError("precision='fp16' (tensor-core prior sampler) has no verbose trace; use the fp32 kernel")
        with torch.cuda.device(zd.device):
            check(lib().damc_prior_langevin_tc(h.ptr, C.c_void_p(zd.data_ptr()), B, K, float(e_l_step_size),
                                               int(bool(e_l_with_noise)), nptr,
                                               _draw_seed() if (seed is None and noise is None and e_l_with_noise) else int(seed or 0),
                                               int(chain0), int(step0), _stream(zd.device)), "damc_prior_langevin_tc")
        set_requires_grad(netE, requires_grad=True)
        return z.detach()
    with torch.cuda.device(zd.device):
        check(lib().damc_prior_langevin(h.ptr, C.c_void_p(zd.data_ptr()), B, K, float(e_l_step_size),
                                        int(bool(e_l_with_noise)), nptr,
                                        _draw_seed() if (seed is None and noise is None and e_l_with_noise) else int(seed or 0),
                                        int(chain0), int(step0),
                                        C.c_void_p(trace.data_ptr()) if trace is not None else None, _stream(zd.device)),
              "damc_prior_langevin")
    if verbose:
        mystr = "Step/en/z_norm: "
        t = trace.cpu().tolist() if trace is not None else []
        for i in range(K):
            if i % 5 == 0 or i == K - 1:
                mystr += "{}/{:.3f}/{:.3f}  ".format(i, t[i][0], t[i][1])
        print("Log prior sampling.")
        print(mystr)
    set_requires_grad(netE, requires_grad=True)
    return z.detach()


def sample_langevin_post_z_with_prior(z, x, netG, netE, g_l_steps, g_llhd_sigma, g_l_with_noise, g_l_step_size,
                                      verbose=False, *, noise=None, seed=None, chain0=0, step0=0, precision=None,
                                      x_hat_out=None):
    """K-step Langevin on |G(z)-x|^2/(2 sigma^2) + E(z) + |z|^2/2.  In-place on ``z.data``; returns ``z.detach()``."""
    zd = _chain_tensor(z)
    B, nz = zd.shape
    gh = pack_generator(netG, precision)
    eh = pack_ebm(netE) if netE is not None else None
    xs = _f32_cuda(x, "x")
    nc, H, W = C.c_int(), C.c_int(), C.c_int()
    gnz = C.c_int()
    check(lib().damc_generator_shape(gh.ptr, C.byref(gnz), C.byref(nc), C.byref(H), C.byref(W)))
    if gnz.value != nz or tuple(xs.shape) != (B, nc.value, H.value, W.value) or xs.device != zd.device:
        raise RuntimeError(f"shape mismatch: z {tuple(zd.shape)}, x {tuple(xs.shape)}, generator nz={gnz.value} "
                           f"-> [{nc.value},{H.value},{W.value}]")
    set_requires_grad(netG, requires_grad=False)
    set_requires_grad(netE, requires_grad=False)
    K = int(g_l_steps)
    keep, nptr = _noise_ptr(noise, (K, B, nz), zd.device)
    trace = torch.empty(K, 4, dtype=torch.float32, device=zd.device) if verbose and K > 0 else None
    if x_hat_out is not None:
        if x_hat_out.shape != xs.shape or x_hat_out.dtype != torch.float32 or not x_hat_out.is_contiguous() or \
                x_hat_out.device != zd.device:
            raise RuntimeError(f"x_hat_out must be a contiguous float32 tensor shaped like x on {zd.device}")
    nbytes = lib().damc_generator_workspace_bytes(gh.ptr, B)
    ws = _workspace(zd.device, nbytes)
    with torch.cuda.device(zd.device):
        check(lib().damc_posterior_langevin(
            gh.ptr, eh.ptr if eh is not None else None, C.c_void_p(zd.data_ptr()), C.c_void_p(xs.data_ptr()), B, K,
            float(g_l_step_size), float(g_llhd_sigma), int(bool(g_l_with_noise)), nptr,
            _draw_seed() if (seed is None and noise is None and g_l_with_noise) else int(seed or 0), int(chain0),
            int(step0), C.c_void_p(trace.data_ptr()) if trace is not None else None,
            C.c_void_p(x_hat_out.data_ptr()) if x_hat_out is not None else None, C.c_void_p(ws.data_ptr()), nbytes,
            _stream(zd.device)), "damc_posterior_langevin")
    if verbose:
        mystr = "Step/cross_entropy/recons_loss: "
        for i, (en, llhd, zn, gm) in enumerate(trace.cpu().tolist() if trace is not None else []):
            mystr += "{}/{:.3f}/{:.3f}/{:.3f}/{:.8f}  ".format(i, en, llhd, zn, gm)
        print("Log posterior sampling.")
        print(mystr)
    set_requires_grad(netG, requires_grad=True)
    set_requires_grad(netE, requires_grad=True)
    return z.detach()


def _den_prec(precision):
    name = precision or DEFAULT_DENOISER_PRECISION
    if name not in _DEN_PRECISIONS:
        raise ValueError(f"denoiser / encoder precision must be one of {list(_DEN_PRECISIONS)} (got {name!r})")
    return name


def _check_rows(t, what, rows, cols, device=None):
    """Raise (like the reference's matmul / reshape would) instead of letting a kernel run past a mis-shaped buffer."""
    if t.dim() != 2 or t.shape[0] != rows or t.shape[1] != cols:
        raise RuntimeError(f"{what} must have shape [{rows}, {cols}], got {tuple(t.shape)}")
    if device is not None:
        device = torch.device(device)
        if t.device.type != device.type or (device.index is not None and t.device.index != device.index):
            raise RuntimeError(f"{what} is on {t.device}, expected {device}")


def generator_forward(netG, z, precision=None):
    """x_hat = netG(z) through the packed kernels (no autograd graph)."""
    zd = _f32_cuda(z, "z")
    gh = pack_generator(netG, precision)
    nz, nc, H, W = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().damc_generator_shape(gh.ptr, C.byref(nz), C.byref(nc), C.byref(H), C.byref(W)))
    if zd.dim() != 2 or zd.shape[1] != nz.value:
        raise RuntimeError(f"z must have shape [B, {nz.value}] for this generator, got {tuple(zd.shape)}")
    B = zd.shape[0]
    out = torch.empty(B, nc.value, H.value, W.value, dtype=torch.float32, device=zd.device)
    nbytes = lib().damc_generator_workspace_bytes(gh.ptr, B)
    ws = _workspace(zd.device, nbytes)
    with torch.cuda.device(zd.device):
        check(lib().damc_generator_forward(gh.ptr, C.c_void_p(zd.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                           C.c_void_p(ws.data_ptr()), nbytes, _stream(zd.device)),
              "damc_generator_forward")
    return out


def gen_samples(bs, nz, netE, netG, e_l_steps, e_l_step_size, e_l_with_noise, *, precision=None):
    """z ~ N(0,I) -> prior Langevin -> G(z).  The reference hard-codes .cuda() (MCMC.py:120); so does this."""
    zk_prior = torch.randn(bs, nz).cuda()
    zk_prior.requires_grad = True
    zk_prior = sample_langevin_prior_z(z=zk_prior, netE=netE, e_l_steps=e_l_steps, e_l_step_size=e_l_step_size,
                                       e_l_with_noise=e_l_with_noise, verbose=False)
    return generator_forward(netG, zk_prior, precision)


def gen_samples_with_diffusion_prior(b, device, netQ, netG, *, precision=None):
    with torch.no_grad():
        zk_prior = netQ(x=None, b=b, device=device)
        x = generator_forward(netG, zk_prior, precision)
    return x, zk_prior


def posterior_score(x, z, netG, netE=None, *, precision=None):
    """(score [B], sqerr [B]) of the eval scripts in ONE library pass: sqerr_b = |G(z_b) - x_b|^2 is reduced where x_hat is
    formed (no x_hat tensor, no separate netG forward), score_b = sqerr_b + E(z_b) + |z_b|^2/2 (no torch netE call).
    Reference: eval_anomaly_det.py:114-117, eval_gen_recon.py:192-193."""
    zd, xs = _f32_cuda(z, "z"), _f32_cuda(x, "x")
    gh = pack_generator(netG, precision)
    eh = pack_ebm(netE) if netE is not None else None
    nz, nc, H, W = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().damc_generator_shape(gh.ptr, C.byref(nz), C.byref(nc), C.byref(H), C.byref(W)))
    if zd.dim() != 2 or zd.shape[1] != nz.value or tuple(xs.shape) != (zd.shape[0], nc.value, H.value, W.value) or \
            xs.device != zd.device:
        raise RuntimeError(f"shape mismatch: z {tuple(zd.shape)}, x {tuple(xs.shape)}, generator nz={nz.value} "
                           f"-> [{nc.value},{H.value},{W.value}]")
    B = zd.shape[0]
    score = torch.empty(B, dtype=torch.float32, device=zd.device)
    sqerr = torch.empty(B, dtype=torch.float32, device=zd.device)
    nbytes = lib().damc_generator_workspace_bytes(gh.ptr, B)
    ws = _workspace(zd.device, nbytes)
    with torch.cuda.device(zd.device):
        check(lib().damc_posterior_score(gh.ptr, eh.ptr if eh is not None else None, C.c_void_p(zd.data_ptr()),
                                         C.c_void_p(xs.data_ptr()), B, C.c_void_p(score.data_ptr()),
                                         C.c_void_p(sqerr.data_ptr()), C.c_void_p(ws.data_ptr()), nbytes,
                                         _stream(zd.device)), "damc_posterior_score")
    return score, sqerr


def recon_mse(x, z, netG, *, precision=None):
    """sum_b mean_pixels (G(z_b) - x_b)^2 -- the quantity accumulated at eval_gen_recon.py:192-194 /
    train_gen_recon.py:340-343 after the noise-free Langevin refinement (fused residual reduction, see posterior_score)."""
    _, sqerr = posterior_score(x, z, netG, None, precision=precision)
    return (sqerr / float(x[0].numel())).sum()


def anomaly_score(x, z, netG, netE, *, precision=None):
    """Per-sample score |G(z)-x|^2 + E(z) + |z|^2/2 of eval_anomaly_det.py:114-119, in one library pass."""
    return posterior_score(x, z, netG, netE, precision=precision)[0]


# ----------------------------------------------------------------------------------------------------------------------
# DAMC ancestral sampler
# ----------------------------------------------------------------------------------------------------------------------
_LAYER_ORDER = (("in_layers", 0), ("in_layers", 1), ("in_layers", 2), ("mid_layers", 0), ("out_layers", 0),
                ("out_layers", 1), ("out_layers", 2))


def pack_denoiser(Q):
    """Q.p = Diffusion_UnetA (reference diffusion_net.py:463-495): time_mlp, B and 7 ConcatSquashLinearSkipCtx."""
    p = Q.p
    layers = [getattr(p, grp)[i] for grp, i in _LAYER_ORDER]
    if len(p.in_layers) != 3 or len(p.mid_layers) != 1 or len(p.out_layers) != 3:
        raise RuntimeError("Q.p must have 3 in_layers, 1 mid_layer and 3 out_layers")
    lin = lambda m: m[0] if isinstance(m, nn.Sequential) else m
    tensors = [p.time_mlp[1].weight, p.time_mlp[1].bias, p.time_mlp[3].weight, p.time_mlp[3].bias, p.B]
    per = []
    for L in layers:
        main, ctx = lin(L._layer), L._layer_ctx[1]
        for m in (main, ctx, L._hyper_bias, L._hyper_gate, L._skip):
            _no_hooks(m, "Q.p layer")
        per.append((main.weight, main.bias, ctx.weight, ctx.bias, L._hyper_gate.weight, L._hyper_gate.bias,
                    L._hyper_bias.weight, L._skip.weight, L._skip.bias))
        tensors += list(per[-1])

    def build():
        keep = []

        def ptr(t):
            c = _f32_cuda(t, "Q.p parameter", True)
            keep.append(c)
            return c.data_ptr()

        d = _lib.DenoiserDesc()
        d.nz, d.nxemb, d.ntemb, d.nf, d.residual = p.nz, p.nxemb, p.ntemb, getattr(p, "nf", 4), int(bool(p.residual))
        d.time_w1, d.time_b1, d.time_w2, d.time_b2 = (ptr(t) for t in tensors[:4])
        d.Bproj = ptr(p.B)
        for i, (W, b, Wc, bc, Wg, bg, Wb, Ws, bs) in enumerate(per):
            d.dim_in[i], d.dim_out[i] = W.shape[1], W.shape[0]
            d.W[i], d.b[i], d.Wc[i], d.bc[i] = ptr(W), ptr(b), ptr(Wc), ptr(bc)
            d.Wg[i], d.bg[i], d.Wb[i], d.Ws[i], d.bs[i] = ptr(Wg), ptr(bg), ptr(Wb), ptr(Ws), ptr(bs)
        out = C.c_void_p()
        check(lib().damc_pack_denoiser(C.byref(out), C.byref(d), _stream(keep[0].device)), "damc_pack_denoiser")
        return _Handle(out), keep

    return _cached(Q, "den", tensors, build)


def _encoder_layers(enc):
    """(conv, instnorm | None) pairs of an Encoder_* module: enc.net = Sequential(Conv2d, InstanceNorm2d, LeakyReLU, ...,
    Conv2d)  (reference diffusion_net.py:233-263 / :274-310 / :321-369).  Raises if the stack has another shape."""
    mods = list(enc.net)
    pairs, i, slope = [], 0, None
    while i < len(mods):
        conv = mods[i]
        if not isinstance(conv, nn.Conv2d):
            raise RuntimeError(f"encoder.net[{i}]: expected Conv2d, found {type(conv).__name__}")
        _no_hooks(conv, f"encoder.net[{i}]")
        if conv.groups != 1 or conv.dilation != (1, 1) or conv.kernel_size[0] != conv.kernel_size[1] or \
                conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] or conv.padding_mode != "zeros":
            raise RuntimeError(f"encoder.net[{i}]: only square, ungrouped, undilated, zero-padded Conv2d is supported")
        if i + 1 == len(mods):
            pairs.append((conv, None))
            break
        norm, act = mods[i + 1], mods[i + 2] if i + 2 < len(mods) else None
        if not isinstance(norm, nn.InstanceNorm2d) or not norm.affine or norm.track_running_stats:
            raise RuntimeError(f"encoder.net[{i + 1}]: expected InstanceNorm2d(affine=True) without running statistics")
        s = _slope_of(act, f"encoder.net[{i + 2}]")
        if slope is not None and s != slope:
            raise RuntimeError("encoder: all LeakyReLU slopes must be equal")
        slope = s
        pairs.append((conv, norm))
        i += 3
    return pairs, (slope if slope is not None else 0.2)


def pack_encoder(enc, height, width, precision):
    pairs, slope = _encoder_layers(enc)
    eps = {float(n.eps) for _, n in pairs if n is not None}
    if len(eps) > 1:
        raise RuntimeError("encoder: InstanceNorm2d layers must share one eps")
    params = [p for c, n in pairs for p in ([c.weight] + ([c.bias] if c.bias is not None else []) +
                                            ([n.weight, n.bias] if n is not None else []))]

    def build():
        keep, arr = [], (_lib.ConvLayer * len(pairs))()

        def ptr(t):
            if t is None:
                return None
            c = _f32_cuda(t, "encoder parameter", True)
            keep.append(c)
            return c.data_ptr()

        for i, (c, n) in enumerate(pairs):
            arr[i] = _lib.ConvLayer(c.in_channels, c.out_channels, c.kernel_size[0], c.stride[0], c.padding[0],
                                    ptr(c.weight), ptr(c.bias), ptr(n.weight if n is not None else None),
                                    ptr(n.bias if n is not None else None))
        out = C.c_void_p()
        check(lib().damc_pack_encoder(C.byref(out), len(pairs), arr, int(height), int(width), slope,
                                      eps.pop() if eps else 1e-5, _PRECISIONS[precision], _stream(keep[0].device)),
              "damc_pack_encoder")
        return _Handle(out), keep

    return _cached(enc, f"enc_{precision}_{height}x{width}", params, build)


def encoder_forward(enc, x, precision=None):
    """xemb = enc(x) [B, nemb] through libdamc_b200: direct first convolution, k4-s2-p1 convolutions as the generator
    engines' stride-2 GEMMs (tcgen05 for 'bf16' / 'fp16', CUDA-core for 'fp32'), fused InstanceNorm + LeakyReLU kernels.
    Odd-sized maps (the 28x28 MNIST encoder: 28 -> 14 -> 7 -> 3 -> 1) are handled on zero-padded parity planes.
    Raises for encoders outside that family."""
    precision = _den_prec(precision)
    xs = _f32_cuda(x, "x")
    if xs.dim() != 4:
        raise RuntimeError("x must be [B, nc, H, W]")
    B, _, H, W = xs.shape
    h = pack_encoder(enc, H, W, precision)
    out = torch.empty(B, enc.nemb, dtype=torch.float32, device=xs.device)
    nbytes = lib().damc_encoder_workspace_bytes(h.ptr, B)
    ws = _workspace(xs.device, nbytes)
    with torch.cuda.device(xs.device):
        check(lib().damc_encoder_forward(h.ptr, C.c_void_p(xs.data_ptr()), C.c_void_p(out.data_ptr()), B,
                                         C.c_void_p(ws.data_ptr()), nbytes, _stream(xs.device)), "damc_encoder_forward")
    return out


def _encoder_on_library(enc, x):
    """True if Q.encoder is a stack damc_pack_encoder accepts for this image size (even maps down to the final k x k)."""
    try:
        pairs, _ = _encoder_layers(enc)
    except (RuntimeError, AttributeError):
        return False
    H, W = x.shape[-2:]
    for i, (c, n) in enumerate(pairs):
        k, s, p = c.kernel_size[0], c.stride[0], c.padding[0]
        if i == 0:
            ok = (k, s, p) == (3, 1, 1) and c.in_channels <= 4 and c.out_channels % 64 == 0
        elif i < len(pairs) - 1:
            ok = (k, s, p) == (4, 2, 1) and H >= 2 and W >= 2 and c.in_channels % 64 == 0 and c.out_channels % 64 == 0
            H, W = (H - 2) // 2 + 1, (W - 2) // 2 + 1   # odd maps (MNIST: 7 -> 3) run on zero-padded parity planes
        else:
            ok = s == 1 and p == 0 and k == H == W and (k * k * c.in_channels) % 64 == 0 and c.out_channels % 16 == 0
        if not ok:
            return False
    return len(pairs) >= 3


def logsnr_table(T, logsnr_min, logsnr_max):
    """lambda(t_i), i = 0..T-1, evaluated in fp32 with the reference's formula (diffusion_helper_func.py:41-50,
    called at diffusion_net.py:599 with t = i/(T-1))."""
    t = torch.arange(T, dtype=torch.float32) / (T - 1.0)
    b = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_max)))
    a = torch.arctan(torch.exp(-0.5 * torch.full_like(t, logsnr_min))) - b
    return (-2.0 * torch.log(torch.tan(a * t + b))).numpy().astype(np.float32)


def damc_sample(Q, x=None, b=None, device=None, cond_w=-1, noise=None, *, z_init=None, seed=None, chain0=0,
                precision=None, xemb=None, encoder_precision=None):
    """DAMC ancestral sampler: z_T ~ N(0,I), T reverse steps of the latent denoiser, returns z_0 [b,nz].
    The image encoder / prior embedding run once in PyTorch; the T-step loop runs in libdamc_b200.
    noise: optional [T-1,b,nz] injected normals; z_init: optional z_T (otherwise torch.randn as the reference)."""
    if x is not None and cond_w is not None and cond_w > 0:
        raise NotImplementedError("classifier-free guidance (cond_w > 0) is never taken by the reference's callers")
    prec = _PRECISIONS[_den_prec(precision)]
    with torch.no_grad():
        if xemb is not None:  # precomputed context embedding (skips the encoder / prior_emb)
            b, device = len(xemb), xemb.device
        elif x is not None:
            assert b is None and device is None
            b, device = len(x), x.device
            # tensor-core modes: the encoder runs on the library too (same operand precision as the denoiser GEMMs);
            # the fp32 parity mode keeps torch's fp32 convolutions, as do encoders outside the supported family
            # (encoder_precision: run the encoder on the library whatever the denoiser mode -- e.g. MNIST, whose nz = 8
            #  denoiser is below the tensor-core granularity and runs in the fp32 persistent kernel)
            enc_prec = encoder_precision or (None if prec == _lib.PREC_FP32 else (precision or DEFAULT_DENOISER_PRECISION))
            if enc_prec is not None and x.is_cuda and _encoder_on_library(Q.encoder, x):
                xemb = encoder_forward(Q.encoder, x, enc_prec)
            else:
                xemb = Q.encoder(x)
        else:
            device = torch.device(device)
            xemb = Q.prior_emb(torch.randn(b, Q.nz, device=device))
        zt = torch.randn(b, Q.nz).to(device) if z_init is None else z_init.detach().clone().to(device)
    if device.type != "cuda":
        raise RuntimeError("damc_sample needs a CUDA device (damc_b200 has no CPU fallback)")
    xemb = _f32_cuda(xemb, "xemb")
    zt = _f32_cuda(zt, "z_T")
    _check_rows(xemb, "xemb", b, int(Q.nxemb), device)
    _check_rows(zt, "z_T", b, int(Q.nz), device)
    T = int(Q.n_interval)
    h = pack_denoiser(Q)
    keep, nptr = _noise_ptr(noise, (T - 1, b, Q.nz), device)
    lam = logsnr_table(T, Q.logsnr_min, Q.logsnr_max)
    lam_c = (C.c_float * (T + 1))(*lam.tolist(), 0.0)
    var_type = {"small": 0, "large": 1}[Q.var_type]
    nbytes = lib().damc_denoise_workspace_bytes(h.ptr, b, T, prec)
    ws = _workspace(device, nbytes)
    with torch.cuda.device(device):
        check(lib().damc_denoise(h.ptr, C.c_void_p(zt.data_ptr()), C.c_void_p(xemb.data_ptr()), b, T, lam_c, var_type,
                                 int(bool(Q.with_noise)), nptr,
                                 _draw_seed() if (seed is None and noise is None) else int(seed or 0), int(chain0),
                                 prec, C.c_void_p(ws.data_ptr()), nbytes, _stream(device)), "damc_denoise")
    return zt


def denoiser_eps(Q, z, logsnr, xemb, *, precision=None):
    """One eps-prediction Q.p(z, logsnr, xemb) with a batch-constant logsnr (float) through the CUDA library."""
    zd, xe = _f32_cuda(z, "z"), _f32_cuda(xemb, "xemb")
    prec = _PRECISIONS[_den_prec(precision)]
    if zd.dim() != 2:
        raise RuntimeError(f"z must be [B, {int(Q.nz)}], got {tuple(zd.shape)}")
    _check_rows(zd, "z", zd.shape[0], int(Q.nz))
    _check_rows(xe, "xemb", zd.shape[0], int(Q.nxemb), zd.device)
    h = pack_denoiser(Q)
    out = torch.empty_like(zd)
    nbytes = lib().damc_denoise_workspace_bytes(h.ptr, zd.shape[0], 1, prec)
    ws = _workspace(zd.device, nbytes)
    with torch.cuda.device(zd.device):
        check(lib().damc_denoiser_eps(h.ptr, C.c_void_p(zd.data_ptr()), C.c_void_p(xe.data_ptr()), float(logsnr),
                                      C.c_void_p(out.data_ptr()), zd.shape[0], prec, C.c_void_p(ws.data_ptr()), nbytes,
                                      _stream(zd.device)), "damc_denoiser_eps")
    return out


# ----------------------------------------------------------------------------------------------------------------------
# toy example (reference toy_example/toy_example.py:110-131)
# ----------------------------------------------------------------------------------------------------------------------
def pack_toy_generator(netG):
    """netG.net = Sequential(Linear, ReLU, Linear, ReLU, Linear, ReLU, Linear)  (toy_example.py:26-34)."""
    seq = netG.net
    lins = [m for m in seq if isinstance(m, nn.Linear)]
    if len(lins) != 4 or len(seq) != 7 or not all(isinstance(seq[i], nn.ReLU) for i in (1, 3, 5)):
        raise RuntimeError("toy netG.net must be Linear-ReLU-Linear-ReLU-Linear-ReLU-Linear")
    params = [p for l in lins for p in (l.weight, l.bias)]

    def build():
        ws = [_f32_cuda(l.weight, "toy weight", True) for l in lins]
        bs = [_f32_cuda(l.bias, "toy bias", True) for l in lins]
        Wp = (C.c_void_p * 4)(*[w.data_ptr() for w in ws])
        bp = (C.c_void_p * 4)(*[t.data_ptr() for t in bs])
        out = C.c_void_p()
        check(lib().damc_pack_toy_mlp(C.byref(out), lins[0].in_features, lins[0].out_features, lins[3].out_features,
                                      Wp, bp, _stream(ws[0].device)), "damc_pack_toy_mlp")
        return _Handle(out), ws + bs

    return _cached(netG, "toy", params, build)


def sample_langevin_post_z(z, x, netG, g_l_steps, g_l_with_noise, g_l_step_size, verbose=False, *, sigma=0.25,
                           noise=None, seed=None, chain0=0, step0=0):
    """Toy posterior sampler: the reference defines it as a closure over netG with sigma = .25 hard-coded
    (toy_example.py:117); here netG is an explicit argument.  All K steps run in one persistent kernel."""
    zd = _chain_tensor(z)
    B, nz = zd.shape
    h = pack_toy_generator(netG)
    xs = _f32_cuda(x, "x")
    K = int(g_l_steps)
    keep, nptr = _noise_ptr(noise, (K, B, nz), zd.device)
    with torch.cuda.device(zd.device):
        check(lib().damc_toy_posterior_langevin(
            h.ptr, C.c_void_p(zd.data_ptr()), C.c_void_p(xs.data_ptr()), B, K, float(g_l_step_size), float(sigma),
            int(bool(g_l_with_noise)), nptr,
            _draw_seed() if (seed is None and noise is None and g_l_with_noise) else int(seed or 0), int(chain0),
            int(step0), _stream(zd.device)), "damc_toy_posterior_langevin")
    if verbose:
        print("Log posterior sampling.")
    return z.detach()
