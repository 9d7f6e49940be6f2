"""Q.calculate_loss with the Linear layers of the epsilon-network on the library's tcgen05 GEMMs.

Reference: workspace/src/diffusion_net.py:624-646 (calculate_loss), :463-533 (Diffusion_UnetA), :417-445
(ConcatSquashLinearSkipCtx); called six times per training iteration at train_gen_recon.py:211-220.

What runs where.  The 35 Linear layers of the seven ConcatSquashLinearSkipCtx blocks are 16 forward + 30 backward GEMM launches
of ``damc_gemm_tf32`` (TF32 operands, fp32 accumulate -- the arithmetic torch uses for these layers when
``torch.backends.cuda.matmul.allow_tf32`` is on):
  * the seven ctx projections share their input SiLU([temb, xemb]) and are ONE GEMM with the concatenated weight [1408, 1152];
  * per block, (main | skip) share the input h and (gate | hyper-bias) share c: two GEMMs with concatenated weights;
  * backward: input-gradients are the same GEMM against transposed weight copies, weight-gradients are dY^T X with the batch
    as the K axis (operands transposed and zero-padded to a multiple of 32 rows).
The gating / SiLU / LeakyReLU algebra between the GEMMs is elementwise torch inside ONE autograd.Function (no autograd graph is
recorded for it); the encoder, prior_emb, time_mlp and the loss stay ordinary autograd -- their gradients flow through the
Function's ``ctx`` input.  No CPU or cuBLAS fallback: the GEMMs fail loudly without the library.
"""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from ._lib import lib, check

_LAYERS = (("in_layers", 0), ("in_layers", 1), ("in_layers", 2), ("mid_layers", 0), ("out_layers", 0), ("out_layers", 1),
           ("out_layers", 2))


def _tf32(t):
    """Contiguous copy of t rounded to the nearest TF32 values (damc_round_tf32): the MMA would otherwise truncate the operand."""
    src = t.contiguous()
    dst = torch.empty_like(src)
    with torch.cuda.device(t.device):
        check(lib().damc_round_tf32(C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), src.numel(),
                                    C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)), "damc_round_tf32")
    return dst


def gemm(A, W, bias=None, out=None):
    """out[m, n] = sum_k A[m, k] W[n, k] (+ bias[n]) through damc_gemm_tf32.  A [M, K], W [N, K] fp32 CUDA tensors (rounded to
    TF32 copies here)."""
    assert A.is_cuda and A.dtype == torch.float32 and W.dtype == torch.float32 and A.dim() == 2 and W.dim() == 2
    A, W = _tf32(A), _tf32(W)
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise RuntimeError(f"gemm: A is [{M}, {K}], W is {tuple(W.shape)}")
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    ldd = out.stride(0)
    with torch.cuda.device(A.device):
        check(lib().damc_gemm_tf32(C.c_void_p(A.data_ptr()), C.c_void_p(W.data_ptr()),
                                   C.c_void_p(bias.data_ptr()) if bias is not None else None, C.c_void_p(out.data_ptr()), M, N, K,
                                   ldd, C.c_void_p(torch.cuda.current_stream(A.device).cuda_stream)), "damc_gemm_tf32")
    return out


def _t_pad(X):
    """X [B, F] -> X^T [F, B'] contiguous with B' = B rounded up to 32 (zero columns): the K-major operand of a weight-gradient."""
    B, Fd = X.shape
    Bp = (B + 31) // 32 * 32
    out = X.new_zeros(Fd, Bp)
    out[:, :B] = X.t()
    return out


def _dsilu(x):
    s = torch.sigmoid(x)
    return s * (1 + x * (1 - s))


class _CoreFn(torch.autograd.Function):
    """eps_hat - z (the seven-block U-net of Diffusion_UnetA) as a function of (z, ctx = [temb, xemb], p.B, block parameters)."""

    @staticmethod
    def forward(fctx, z, ctx, Bproj, *params):
        L = [params[9 * i:9 * i + 9] for i in range(7)]   # (W, b, Wc, bc, Wg, bg, Wb, Ws, bs) per block
        douts = [l[0].shape[0] for l in L]
        offs = [sum(douts[:i]) for i in range(7)]
        s = F.silu(ctx)
        Wc_all = torch.cat([l[2] for l in L], 0)
        c_pre = gemm(s, Wc_all, torch.cat([l[3] for l in L], 0))
        c_all = F.silu(c_pre)
        proj = (2 * math.pi) * gemm(z, Bproj.t())
        sn, cs = torch.sin(proj), torch.cos(proj)
        outs, saved = [], []
        for i, (W, b, Wc, bc, Wg, bg, Wb, Ws, bs) in enumerate(L):
            d = douts[i]
            # block input (reference forward :515-528): the embedding, the activated previous output, or -- out blocks -- the
            # activated concatenation of the previous output with the matching in-block output
            if i == 0:
                x = torch.cat([sn, cs, z], 1)
            elif i < 4:
                x = F.leaky_relu(outs[i - 1], 0.01)
            else:
                x = F.leaky_relu(torch.cat([outs[i - 1], outs[6 - i]], 1), 0.01)
            c = c_all[:, offs[i]:offs[i] + d].contiguous()
            W_ms, W_gb = torch.cat([W, Ws], 0), torch.cat([Wg, Wb], 0)
            mk = gemm(x, W_ms, torch.cat([b, bs], 0))                          # [main | skip]
            gh = gemm(c, W_gb, torch.cat([bg, torch.zeros_like(bg)], 0))       # [gate | hyper-bias]
            m, gate = mk[:, :d], torch.sigmoid(gh[:, :d])
            outs.append(m * gate + gh[:, d:] + mk[:, d:])
            saved += [x, c, m, gate, W_ms, W_gb]
        fctx.douts = douts
        fctx.save_for_backward(z, ctx, s, c_pre, Wc_all, sn, cs, *saved)
        return outs[6]

    @staticmethod
    def backward(fctx, d_out):
        z, ctx, s, c_pre, Wc_all, sn, cs, *flat = fctx.saved_tensors
        saved = [flat[6 * i:6 * i + 6] for i in range(7)]
        douts = fctx.douts
        offs = [sum(douts[:i]) for i in range(7)]
        dc_pre = torch.empty_like(c_pre)
        grads = [None] * 63
        g_o = [None] * 7                      # gradient with respect to every block's output
        g_o[6] = d_out.contiguous()
        acc = lambda cur, new: new if cur is None else cur + new
        d_emb = None
        for i in range(6, -1, -1):
            x, c, m, gate, W_ms, W_gb = saved[i]
            d, g = douts[i], g_o[i]
            d_mk = torch.cat([g * gate, g], 1)                                # d[main | skip]
            d_gh = torch.cat([g * m * gate * (1 - gate), g], 1)               # d[gate pre-activation | hyper-bias]
            dx = gemm(d_mk, W_ms.t())
            dc = gemm(d_gh, W_gb.t())
            dW_ms = gemm(_t_pad(d_mk), _t_pad(x))
            dW_gb = gemm(_t_pad(d_gh), _t_pad(c))
            db_ms, db_g = d_mk.sum(0), d_gh[:, :d].sum(0)
            dc_pre[:, offs[i]:offs[i] + d] = dc * _dsilu(c_pre[:, offs[i]:offs[i] + d])
            grads[9 * i + 0], grads[9 * i + 7] = dW_ms[:d], dW_ms[d:]
            grads[9 * i + 1], grads[9 * i + 8] = db_ms[:d], db_ms[d:]
            grads[9 * i + 4], grads[9 * i + 6] = dW_gb[:d], dW_gb[d:]
            grads[9 * i + 5] = db_g
            if i == 0:
                d_emb = dx
                continue
            dpre = dx * torch.where(x > 0, 1.0, 0.01).to(dx.dtype)            # x = leaky_relu(pre): same sign as pre
            if i < 4:
                g_o[i - 1] = acc(g_o[i - 1], dpre)
            else:
                w = douts[i - 1]
                g_o[i - 1] = acc(g_o[i - 1], dpre[:, :w])
                g_o[6 - i] = acc(g_o[6 - i], dpre[:, w:])
        d_ctx = gemm(dc_pre, Wc_all.t()) * _dsilu(ctx)
        dWc_all = gemm(_t_pad(dc_pre), _t_pad(s))
        dbc_all = dc_pre.sum(0)
        for i in range(7):
            grads[9 * i + 2] = dWc_all[offs[i]:offs[i] + douts[i]]
            grads[9 * i + 3] = dbc_all[offs[i]:offs[i] + douts[i]]
        half = sn.shape[1]
        d_proj = (2 * math.pi) * (d_emb[:, :half] * cs - d_emb[:, half:2 * half] * sn)
        dB = z.t() @ d_proj                                 # [nz, nz/2]: one tiny product, left to torch
        return (None, d_ctx, dB, *grads)


_graphed = {}   # (device, shapes) -> CUDA-graphed forward / backward of _CoreFn (torch.cuda.make_graphed_callables)


def eps_network(p, z, logsnr, xemb, graphed=False):
    """Q.p(z, logsnr, xemb) (reference Diffusion_UnetA.forward, diffusion_net.py:501-533) with the block GEMMs on the library.
    graphed=True: the ~350 launches of the core's forward and of its backward (46 library GEMMs + the elementwise algebra) are
    each captured once per shape into a CUDA graph and replayed -- at 128 chains the eager form is bound by launch overhead."""
    if not z.is_cuda:
        raise RuntimeError("denoiser_train.eps_network needs CUDA tensors (damc_b200 has no CPU fallback)")
    widths = [p.nz, p.nxemb + p.ntemb] + [getattr(p, g)[i]._skip.in_features for g, i in _LAYERS] + \
             [getattr(p, g)[i]._skip.out_features for g, i in _LAYERS]
    if any(w % 32 for w in widths):
        raise RuntimeError(f"engine='library' needs layer widths that are multiples of 32 (TF32 k-blocks); got {sorted(set(widths))} -- "
                           "use engine='torch' for this network")
    u = torch.arctan(torch.exp(-0.5 * torch.clamp(logsnr, min=-20.0, max=20.0))) / (0.5 * math.pi)
    ctx = torch.cat([p.time_mlp(u), xemb], dim=1)
    params = []
    for grp, i in _LAYERS:
        Lr = getattr(p, grp)[i]
        main = Lr._layer[0] if isinstance(Lr._layer, torch.nn.Sequential) else Lr._layer
        cl = Lr._layer_ctx[1]
        params += [main.weight, main.bias, cl.weight, cl.bias, Lr._hyper_gate.weight, Lr._hyper_gate.bias, Lr._hyper_bias.weight,
                   Lr._skip.weight, Lr._skip.bias]
    args = (z, ctx, p.B, *params)
    if graphed:
        key = (z.device, tuple(z.shape), tuple(ctx.shape), tuple(tuple(t.shape) for t in params))
        fn = _graphed.get(key)
        if fn is None:
            sample = tuple(a.detach().clone().requires_grad_(a.requires_grad) for a in args)
            fn = torch.cuda.make_graphed_callables(lambda *a: _CoreFn.apply(*a), sample)
            _graphed[key] = fn
        out = fn(*args)
    else:
        out = _CoreFn.apply(*args)
    return z + out if p.residual else out
