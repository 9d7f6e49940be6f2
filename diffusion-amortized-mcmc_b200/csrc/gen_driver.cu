// gen_driver.cu -- host-side planning for the generator: weight packing, workspace carving and the per-step chain of
// shifted-window GEMMs (forward G(z), likelihood gradient, backward with respect to z).
//
// Replaces, per Langevin step, netG(z) + autograd.grad through netG of sample_langevin_post_z_with_prior (reference
// workspace/src/MCMC.py:55-60) for the generator families of workspace/src/diffusion_net.py:20-203.
#include <stdlib.h>

#include <algorithm>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

// split-K factor of the first layer's dgrad (M = chains only): enough tiles to fill the SMs, no empty K range
static int dz_splits_for(const GenPack* g, int B) {
  const int mt = ceil_div(B, 128);
  const GenLayer& f = g->layers[0];
  const int kb = f.k * f.k * f.cout / 64;
  // (128 splits at 128 chains were tried twice: round 1 the GEMM dropped 46 -> 38 us but the update kernel's partial-sum reads grew
  //  by as much; round 2, with the partials pre-summed by dz_reduce_kernel, the step time did not move (17.2 ms either way) -- and a
  //  cap that binds at every tested batch size is what keeps the split count, hence the summation order, shard-invariant)
  int s = std::max(1, std::min(std::min(32, kb), 296 / mt));
  while (s > 1 && (long long)ceil_div(kb, s) * (s - 1) >= kb) --s;
  return s;
}

static int dev_alloc(GenPack* g, void** p, size_t bytes) {
  DAMC_CUDA(cudaMalloc(p, bytes));
  g->allocs.push_back(*p);
  return DAMC_OK;
}

int build_generator(GenPack* g, int nlayers, const damc_convt_layer* L, float slope, int precision,
                    cudaStream_t stream) {
  if (nlayers < 2 || nlayers > 8) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "generator needs 2..8 ConvTranspose2d layers (got %d)", nlayers);
  if (precision != DAMC_PREC_FP32 && !is_gen_tc_precision(precision)) DAMC_FAIL(DAMC_ERR_INVALID, "unknown precision %d", precision);
  g->kind = H_GEN;
  g->precision = precision;
  g->nlayers = nlayers;
  g->slope = slope;
  g->nz = L[0].cin;
  g->nz_p = (int)align_up(g->nz, 64);
  const int cmult = is_gen_tc_precision(precision) ? 64 : 16;
  int H = 1, W = 1;
  g->layers.resize(nlayers);
  for (int i = 0; i < nlayers; ++i) {
    GenLayer& y = g->layers[i];
    const damc_convt_layer& s = L[i];
    if (s.weight == nullptr) DAMC_FAIL(DAMC_ERR_INVALID, "layer %d: null weight", i);
    if (i > 0 && s.cin != L[i - 1].cout) DAMC_FAIL(DAMC_ERR_INVALID, "layer %d: cin %d != previous cout %d", i, s.cin, L[i - 1].cout);
    y.cin = s.cin; y.cout = s.cout; y.k = s.k; y.stride = s.stride; y.pad = s.pad; y.Hin = H; y.Win = W;
    if (i == 0) {
      if (s.stride != 1 || s.pad != 0) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "first layer must be 1x1 -> kxk with stride 1, padding 0");
      y.type = L_FIRST;
      y.Hout = y.Wout = s.k;
    } else if (s.k == 4 && s.stride == 2 && s.pad == 1) {
      y.type = L_UP;
      y.Hout = 2 * H; y.Wout = 2 * W;
    } else if (s.k == 3 && s.stride == 1 && s.pad == 1 && i == nlayers - 1) {
      y.type = L_SAME;
      y.Hout = H; y.Wout = W;
    } else {
      DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "layer %d: ConvTranspose2d(k=%d,s=%d,p=%d) is not one of the reference's shapes", i, s.k, s.stride, s.pad);
    }
    if (i < nlayers - 1 && s.cout % cmult) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "layer %d: %d channels; this precision needs a multiple of %d", i, s.cout, cmult);
    if (i == nlayers - 1 && (s.cout > 4 || s.k * s.k > 16)) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "last layer: nc <= 4 and k <= 4 supported (nc=%d k=%d)", s.cout, s.k);
    y.cin_p = i == 0 ? g->nz_p : s.cin;
    H = y.Hout; W = y.Wout;
  }
  g->nc = L[nlayers - 1].cout; g->H = H; g->W = W;

  g->src.assign(L, L + nlayers);
  const char* env = getenv("DAMC_TC");
  g->use_tc = is_gen_tc_precision(precision) && !(env && env[0] == '0');
  g->use_bits = g->use_tc && (precision == DAMC_PREC_TF32 || !getenv("DAMC_TC_NOBITS"));
  if (g->use_tc && !tc_available()) DAMC_FAIL(DAMC_ERR_CUDA, "bf16 mode needs cuTensorMapEncodeTiled from the driver (no fallback)");
  DAMC_TRY(g->refill(stream, nullptr));
  g->last_fused = last_fused_supported(g);
  return handle_hash_init(g, stream);
}

void GenPack::sources(std::vector<HashSrc>& out) const {
  for (const damc_convt_layer& s : src) {
    out.push_back(HashSrc{s.weight, (unsigned long long)s.cin * s.cout * s.k * s.k, 0ull});
    if (s.bias) out.push_back(HashSrc{s.bias, (unsigned long long)s.cout, 0ull});
  }
}

// (re)pack every layer's GEMM operands from the caller's ConvTranspose2d tensors; buffers are allocated on first use.
// Only the layouts the chosen engine reads are packed: [tap][n][c] K-major for tcgen05, [tap][c][n] for the CUDA-core engine.
int GenPack::refill(cudaStream_t stream, const int* dirty) {
  GenPack* g = this;
  const size_t es = elem_size(precision);
  const bool simt = !use_tc;
  for (int i = 0; i < nlayers; ++i) {
    GenLayer& y = g->layers[i];
    const damc_convt_layer& s = g->src[i];
    const bool last = i == nlayers - 1;
    // bias (fp32 device copy)
    if (!y.bias) {
      DAMC_TRY(dev_alloc(g, (void**)&y.bias, sizeof(float) * s.cout));
      DAMC_CUDA(cudaMemsetAsync(y.bias, 0, sizeof(float) * s.cout, stream));
    }
    if (s.bias) DAMC_TRY(launch_gated_copy(y.bias, s.bias, s.cout, dirty, stream));
    // forward operands
    int ntaps, Cs, mode, ncls = 1;
    if (y.type == L_FIRST) { ntaps = 1; Cs = y.cin_p; mode = PK_FIRST_FWD; y.n_fwd = s.k * s.k * s.cout; }
    else if (y.type == L_UP) { ntaps = 4; Cs = y.cin; mode = PK_UP_FWD; ncls = 4; y.n_fwd = s.cout; }
    else { ntaps = 9; Cs = y.cin; mode = PK_SAME_FWD; y.n_fwd = s.cout; }
    y.np_fwd = (int)align_up(y.n_fwd, 16);
    if (last) {  // decided before the forward operands: the scatter form replaces the multi-tap forward of the last layer
      y.n_sc = s.k * s.k * s.cout;
      y.np_sc = (int)align_up(y.n_sc, 16);
      const char* env = getenv("DAMC_LAST_SCATTER");
      g->last_scatter = last_finish_smem(y) <= 200 * 1024 && !(env && env[0] == '0');
    }
    for (int c = 0; c < ncls && !(last && g->last_scatter); ++c) {
      if (simt) {
        if (!y.w_fwd[c]) DAMC_TRY(dev_alloc(g, &y.w_fwd[c], es * (size_t)ntaps * Cs * y.np_fwd));
        DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, mode, c, ntaps, Cs, y.np_fwd, 0,
                                   precision, y.w_fwd[c], stream, dirty));
      } else {
        if (!y.w_fwd_tc[0]) DAMC_TRY(dev_alloc(g, &y.w_fwd_tc[0], es * (size_t)ncls * ntaps * Cs * y.np_fwd));
        y.w_fwd_tc[c] = (char*)y.w_fwd_tc[0] + es * (size_t)c * ntaps * Cs * y.np_fwd;  // class blocks back to back
        DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, mode, c, ntaps, Cs, y.np_fwd, 1,
                                   precision, y.w_fwd_tc[c], stream, dirty));
      }
    }
    // dgrad operands
    if (last) { ntaps = 1; Cs = 64; mode = PK_LAST_DGRAD_COL; }
    else if (y.type == L_FIRST) { ntaps = 1; Cs = s.k * s.k * s.cout; mode = PK_FIRST_DGRAD; }
    else { ntaps = 16; Cs = s.cout; mode = PK_UP_DGRAD; }
    y.n_dg = s.cin;
    y.np_dg = (int)align_up(y.n_dg, 16);
    if (simt) {
      if (!y.w_dgrad) DAMC_TRY(dev_alloc(g, &y.w_dgrad, es * (size_t)ntaps * Cs * y.np_dg));
      DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, mode, 0, ntaps, Cs, y.np_dg, 0,
                                 precision, y.w_dgrad, stream, dirty));
    } else {
      if (!y.w_dgrad_tc) DAMC_TRY(dev_alloc(g, &y.w_dgrad_tc, es * (size_t)ntaps * Cs * y.np_dg));
      DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, mode, 0, ntaps, Cs, y.np_dg, 1,
                                 precision, y.w_dgrad_tc, stream, dirty));
    }
    if (last && g->last_scatter) {  // scatter-form operands: one tap, N = k*k*nc
      if (simt) {
        if (!y.w_scatter) DAMC_TRY(dev_alloc(g, &y.w_scatter, es * (size_t)y.cin * y.np_sc));
        DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, PK_LAST_FWD_SCATTER, 0, 1, y.cin,
                                   y.np_sc, 0, precision, y.w_scatter, stream, dirty));
      } else {
        if (!y.w_scatter_tc) DAMC_TRY(dev_alloc(g, &y.w_scatter_tc, es * (size_t)y.cin * y.np_sc));
        DAMC_TRY(launch_pack_convt(s.weight, s.cin, s.cout, s.k, s.stride, s.pad, PK_LAST_FWD_SCATTER, 0, 1, y.cin,
                                   y.np_sc, 1, precision, y.w_scatter_tc, stream, dirty));
      }
    }
  }
  return DAMC_OK;
}

int plan_workspace(const GenPack* g, int B, void* base, GenWorkspace* ws) {
  const size_t es = elem_size(g->precision);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align_up(bytes, 256); return base ? (void*)((char*)base + r) : nullptr; };
  const int L = g->nlayers;
  ws->zin = take(es * (size_t)B * g->nz_p);
  ws->act.assign(L - 1, nullptr);
  ws->grad.assign(L - 1, nullptr);
  ws->mask.assign(L - 1, nullptr);
  for (int l = 0; l < L - 1; ++l) {
    const GenLayer& y = g->layers[l];
    const size_t n = (size_t)B * y.Hout * y.Wout * y.cout;
    ws->act[l] = take(es * n);
    ws->grad[l] = take(es * n);
    if (g->use_bits) ws->mask[l] = (uint32_t*)take(n / 8);
  }
  const GenLayer& last = g->layers[L - 1];
  ws->gcol = take(es * (size_t)B * last.Hin * last.Win * 64);
  ws->ybuf = g->last_scatter ? (float*)take(sizeof(float) * (size_t)B * last.Hin * last.Win * last.np_sc) : nullptr;
  ws->dz_part = (float*)take(sizeof(float) * (size_t)dz_splits_for(g, B) * B * g->nz_p);
  ws->zbuf = (float*)take(sizeof(float) * (size_t)B * g->nz);
  ws->xbuf = (float*)take(sizeof(float) * (size_t)B * g->nc * g->H * g->W);
  ws->xhat_buf = (float*)take(sizeof(float) * (size_t)B * g->nc * g->H * g->W);
  ws->sq_part = (float*)take(sizeof(float) * (size_t)B * score_parts(g));
  ws->seed_dev = (unsigned long long*)take(32);
  ws->base = base;
  ws->bytes = o;
  return DAMC_OK;
}

static int run_gemm(const GenPack* g, const GemmPlan& p, cudaStream_t stream) {
  profile_mark(stream, true);
  if ((g->use_tc ? p.Wtc : p.W) == nullptr) DAMC_FAIL(DAMC_ERR_INVALID, "generator GEMM: operand layout of the active engine was not packed");
  const int r = g->use_tc ? launch_gemm_tc(p, g->precision, stream) : launch_gemm_simt(p, g->precision, stream);
  profile_mark(stream, false);
  count_launch();
  return r;
}

static void up_fwd_taps(int cls, GemmPlan& p) {
  const int py = cls >> 1, px = cls & 1;
  p.ntaps = 4;
  for (int ty = 0; ty < 2; ++ty)
    for (int tx = 0; tx < 2; ++tx) {
      Tap& t = p.taps[ty * 2 + tx];
      t.plane = 0;
      t.dy = (signed char)(py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0));
      t.dx = (signed char)(px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0));
    }
}

int generator_forward(const GenPack* g, const GenWorkspace& ws, const float* z, int B, const float* x, float sigma,
                      float* xhat, float* loss, cudaStream_t stream, float* sq_part) {
  DAMC_TRY(launch_stage_z(z, ws.zin, B, g->nz, g->nz_p, g->precision, stream));
  count_launch();
  const int L = g->nlayers;
  for (int l = 0; l < L; ++l) {
    const GenLayer& y = g->layers[l];
    const bool last = l == L - 1;
    GemmPlan p{};
    p.A = l == 0 ? ws.zin : ws.act[l - 1];
    p.plane_stride = 0;
    p.B = B; p.Hm = y.Hin; p.Wm = y.Win; p.Cs = y.cin_p;
    p.N = y.n_fwd; p.Np = y.np_fwd; p.ksplit = 1;
    Epilogue& e = p.epi;
    e.slope = g->slope;
    e.bias = y.bias;
    e.sy = e.sx = y.type == L_UP ? 2 : 1;
    if (last && g->last_fused && x != nullptr) {   // forward + likelihood gradient + this layer's dgrad in one launch
      profile_mark(stream, true);
      DAMC_TRY(launch_last_fused(g, ws, B, x, sigma, xhat, loss, sq_part, stream));
      profile_mark(stream, false);
      count_launch();
      continue;
    }
    if (last && g->last_scatter) {
      p.N = y.n_sc; p.Np = y.np_sc;
      p.ntaps = 1; p.taps[0] = Tap{0, 0, 0, 0};
      p.W = y.w_scatter; p.Wtc = y.w_scatter_tc;
      e.kind = EPI_STORE_F32;
      e.out = ws.ybuf; e.nz_out = y.np_sc;
      DAMC_TRY(run_gemm(g, p, stream));
      DAMC_TRY(launch_last_finish(y, g->precision, ws.ybuf, B, x, xhat, 1.0f / (sigma * sigma),
                                  generator_grad_scale(g, sigma), loss, ws.gcol, stream));
      count_launch();
      continue;
    }
    if (last) {
      e.kind = EPI_FWD_LAST;
      e.x = x; e.xhat = xhat; e.loss = loss; e.gcol = ws.gcol;
      e.inv_sigma2 = 1.0f / (sigma * sigma);
      e.gscale = generator_grad_scale(g, sigma);
      e.nc = y.cout; e.k = y.k; e.stride = y.stride; e.padding = y.pad;
      e.Hi = y.Hin; e.Wi = y.Win; e.Ho = y.Hout; e.Wo = y.Wout;
    } else {
      e.kind = EPI_FWD_ACT;
      e.out = ws.act[l];
      e.maskbits = ws.mask[l];
      e.bias_mod = y.cout;
      if (y.type == L_FIRST) { e.o_b = (long long)y.n_fwd; e.o_y = 0; e.o_x = 0; }
      else { e.o_b = (long long)y.Hout * y.Wout * y.cout; e.o_y = (long long)y.Wout * y.cout; e.o_x = y.cout; }
    }
    if (y.type == L_FIRST) {
      p.ntaps = 1; p.taps[0] = Tap{0, 0, 0, 0};
      p.W = y.w_fwd[0]; p.Wtc = y.w_fwd_tc[0];
      DAMC_TRY(run_gemm(g, p, stream));
    } else if (y.type == L_UP && g->use_tc && !last && !getenv("DAMC_TC_NOMERGE")) {
      up_fwd_taps(0, p);           // (taps are derived from the class inside the kernel)
      p.W = y.w_fwd[0]; p.Wtc = y.w_fwd_tc[0];
      p.ncls = 4;
      DAMC_TRY(run_gemm(g, p, stream));
    } else if (y.type == L_UP) {
      for (int cls = 0; cls < 4; ++cls) {
        up_fwd_taps(cls, p);
        p.W = y.w_fwd[cls]; p.Wtc = y.w_fwd_tc[cls];
        e.py = cls >> 1; e.px = cls & 1;
        DAMC_TRY(run_gemm(g, p, stream));
      }
    } else {  // L_SAME: oh = ih - 1 + kh  ->  source row = oh + 1 - kh
      p.ntaps = 9;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) p.taps[kh * 3 + kw] = Tap{0, (signed char)(1 - kh), (signed char)(1 - kw), 0};
      p.W = y.w_fwd[0]; p.Wtc = y.w_fwd_tc[0];
      DAMC_TRY(run_gemm(g, p, stream));
    }
  }
  return DAMC_OK;
}

int score_parts(const GenPack* g) { return g->last_fused ? last_fused_parts(g) : 1; }

// Eval consumers (reference eval_anomaly_det.py:114-117, eval_gen_recon.py:192-194): G(z) and sum (G(z) - x)^2 per chain.
// 16-bit tensor-core modes: the fused last-layer kernel in score mode reduces the residual where x_hat is formed (x_hat never
// reaches HBM).  Other modes: the ordinary forward into the workspace, then a per-chain reduction kernel.
int generator_score_forward(const GenPack* g, const GenWorkspace& ws, const float* z, int B, const float* x, cudaStream_t stream) {
  if (!g->last_fused) {
    DAMC_TRY(generator_forward(g, ws, z, B, nullptr, 1.0f, ws.xhat_buf, nullptr, stream));
    return launch_sqerr(ws.xhat_buf, x, B, g->nc * g->H * g->W, ws.sq_part, stream);
  }
  return generator_forward(g, ws, z, B, x, 1.0f, nullptr, nullptr, stream, ws.sq_part);
}

int generator_dgrad(const GenPack* g, const GenWorkspace& ws, int B, cudaStream_t stream) {
  const int L = g->nlayers;
  for (int l = g->last_fused ? L - 2 : L - 1; l >= 0; --l) {   // the fused last-layer kernel has already produced grad[L-2]
    const GenLayer& y = g->layers[l];
    GemmPlan p{};
    p.B = B; p.Hm = y.Hin; p.Wm = y.Win;
    p.N = y.n_dg; p.Np = y.np_dg; p.ksplit = 1;
    p.W = y.w_dgrad; p.Wtc = y.w_dgrad_tc;
    Epilogue& e = p.epi;
    e.slope = g->slope;
    if (l == L - 1) {  // im2col'd dL/dh written by the forward epilogue
      p.A = ws.gcol; p.Cs = 64; p.ntaps = 1; p.taps[0] = Tap{0, 0, 0, 0};
    } else if (y.type == L_UP) {  // gin[ih] = sum_kh gout[2 ih - 1 + kh] W[kh]: parity plane (kh+1)&1, shift -1/0/0/+1
      p.A = ws.grad[l]; p.Cs = y.cout; p.ntaps = 16;
      p.plane_stride = (long long)B * y.Hin * y.Win * y.cout;
      for (int kh = 0; kh < 4; ++kh)
        for (int kw = 0; kw < 4; ++kw) {
          Tap& t = p.taps[kh * 4 + kw];
          t.plane = (signed char)((((kh + 1) & 1) << 1) | ((kw + 1) & 1));
          t.dy = (signed char)(kh == 0 ? -1 : (kh == 3 ? 1 : 0));
          t.dx = (signed char)(kw == 0 ? -1 : (kw == 3 ? 1 : 0));
        }
    } else {  // L_FIRST: plain GEMM over (kh,kw,co)
      p.A = ws.grad[0]; p.Cs = y.k * y.k * y.cout; p.ntaps = 1; p.taps[0] = Tap{0, 0, 0, 0};
    }
    if (l == 0) {
      e.kind = EPI_DGRAD_Z;
      e.out = ws.dz_part;
      e.nz_out = g->nz_p;
      p.ksplit = dz_splits_for(g, B);
    } else {
      e.kind = EPI_DGRAD_MASK;
      e.maskbits = ws.mask[l - 1];
      e.act = ws.act[l - 1];
      e.out = ws.grad[l - 1];
      e.planar_out = g->layers[l - 1].type == L_UP;
    }
    DAMC_TRY(run_gemm(g, p, stream));
  }
  return DAMC_OK;
}

// fp16 mode carries dU/dh scaled by sigma^2 (O(1) magnitudes) through the backward chain; the update kernel undoes it
float generator_grad_scale(const GenPack* g, float sigma) { return g->precision == DAMC_PREC_FP16 ? sigma * sigma : 1.0f; }

int dz_splits(const GenPack* g, int B) { return dz_splits_for(g, B); }

}  // namespace damc
