// encoder.cu -- image encoder of the DAMC amortizer: xemb = Q.encoder(x), run once per _netQ_U.forward(x) call
// (reference workspace/src/diffusion_net.py:590; Encoder_cifar10 :227-266, Encoder_celeba64 :268-313,
// Encoder_celebaHQ :315-372: Conv2d(nc,nif,3,1,1) -> [Conv2d(.,.,4,2,1)]* -> Conv2d(.,nemb,4,1,0), each but the last
// followed by InstanceNorm2d(affine) and LeakyReLU(0.2)).  SURVEY.md section 8(f) row 1.
//
// Mapping onto the generator's GEMM engines (no new GEMM kernel):
//   * Conv2d(k4,s2,p1) with weight [Cout,Cin,4,4] IS the input-gradient of ConvTranspose2d(k4,s2,p1) with weight
//     [Cin_t = Cout, Cout_t = Cin, 4, 4] -- the same tensor, read the same way:  out[y] = sum_kh in[2y-1+kh] W[kh].
//     So each down-sampling layer runs the k4-s2-p1 dgrad plan: 16 taps over the 4 parity planes of the (normalised)
//     input, K = 16 Cin, raw fp32 accumulators out (EPI_STORE_F32).
//   * the final Conv2d(k4,s1,p0) on a 4x4 map is a plain GEMM over (kh,kw,c), the first generator layer's dgrad plan,
//     with the bias added in the epilogue (EPI_STORE_F32_BIAS) -> xemb [B,nemb] fp32.
//   * the first Conv2d(k3,s1,p1) has K = 9 nc <= 36: a CUDA-core direct convolution (HBM-bound: it writes 64 channels
//     per pixel).
//   * InstanceNorm + LeakyReLU: one kernel per layer; a CTA owns (image, 32 channels), reduces mean / biased variance over
//     the pixels, then re-reads the slab (L2-hot) and writes the normalised activation in the operand type, already in
//     the layout the next GEMM's TMA boxes want (4 parity planes, or flat NHWC before the last layer).
//     The conv bias of a normalised layer cancels exactly ((x+b) - mean(x+b) = x - mean(x)) and is not added.
//   * odd-sized maps (Encoder_mnist :374-413: 28 -> 14 -> 7 -> 3 -> 1): a k4-s2-p1 convolution of an H x W map equals the
//     same convolution of the map zero-padded to even size, restricted to the first floor((H-2)/2)+1 outputs.  So the
//     parity planes are ceil(H/2) x ceil(W/2) with zeros at the pixels that do not exist (the buffer is cleared per call),
//     the GEMM runs on that padded grid, and the normalisation that follows reads the real outputs only.
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "gen_epilogue.cuh"

namespace damc {

enum EncType { ENC_FIRST = 0, ENC_DOWN = 1, ENC_LAST = 2 };

struct EncLayer {
  int type, cin, cout, k, Hin, Win, Hout, Wout;
  int Hg, Wg;   // pixel grid of the layer's raw fp32 output: = Hout x Wout, or ceil(Hin/2) x ceil(Win/2) for a down layer on an odd map
  damc_conv_layer src;
  void* w_simt = nullptr;   // [ntaps*Cs][N]  (CUDA-core engine)
  void* w_tc = nullptr;     // [ntaps][N][Cs] (tcgen05 engine)
};

struct EncPack : damc_handle {
  int precision = DAMC_PREC_FP32, nlayers = 0, nc = 0, H = 0, W = 0, nemb = 0;
  float slope = 0.2f, eps = 1e-5f;
  bool use_tc = false;
  std::vector<EncLayer> layers;
  std::vector<void*> allocs;
  ~EncPack() override { for (void* p : allocs) cudaFree(p); }
  int refill(cudaStream_t stream, const int* dirty) override;
  void sources(std::vector<HashSrc>& out) const override;
};

struct EncWs {
  float* raw;               // conv output before normalisation, fp32 NHWC (largest layer)
  std::vector<void*> act;   // normalised activations (operand type), input of layer l+1
  size_t bytes;
};

static EncWs enc_ws(const EncPack* e, int B, void* base) {
  EncWs w;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 256); return base ? (void*)((char*)base + r) : nullptr; };
  size_t raw = 0;
  for (int l = 0; l + 1 < e->nlayers; ++l) {
    const EncLayer& y = e->layers[l];
    raw = std::max(raw, sizeof(float) * (size_t)B * y.Hg * y.Wg * y.cout);
  }
  w.raw = (float*)take(raw);
  w.act.assign(e->nlayers - 1, nullptr);
  const size_t es = elem_size(e->precision);
  for (int l = 0; l + 1 < e->nlayers; ++l) {
    const EncLayer& y = e->layers[l];
    const bool planar = e->layers[l + 1].type == ENC_DOWN;   // 4 parity planes of ceil(H/2) x ceil(W/2) pixels
    w.act[l] = take(es * (size_t)B * y.cout * (planar ? 4 * (size_t)((y.Hout + 1) / 2) * ((y.Wout + 1) / 2) : (size_t)y.Hout * y.Wout));
  }
  w.bytes = o;
  return w;
}

// ---- first layer: direct k3-s1-p1 convolution, nc <= 4 input channels ------------------------------------------------
// thread = (pixel, quarter of the output channels): 4 threads write one pixel's contiguous channel row.
__global__ void __launch_bounds__(256) enc_first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, float* __restrict__ raw,
                                                             long long npix, int nc, int H, int W, int C1) {
  extern __shared__ __align__(16) float wsm[];  // [nc*9][C1] then bias [C1]
  const int K = nc * 9;
  for (int i = threadIdx.x; i < K * C1; i += blockDim.x) {
    const int kk = i / C1, c = i - kk * C1;      // w is [C1][nc][3][3]
    wsm[i] = w[(size_t)c * K + kk];
  }
  for (int i = threadIdx.x; i < C1; i += blockDim.x) wsm[K * C1 + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 64 + (threadIdx.x >> 2);
  if (p >= npix) return;
  const int q = threadIdx.x & 3;
  const int px = (int)(p % W), py = (int)((p / W) % H);
  const long long b = p / ((long long)W * H);
  float xin[36];
#pragma unroll
  for (int ci = 0; ci < 4; ++ci)
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = py + t / 3 - 1, xx = px + t % 3 - 1;
      xin[ci * 9 + t] = (ci < nc && yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + ((b * nc + ci) * H + yy) * W + xx) : 0.f;
    }
  const int cq = C1 >> 2;  // channels per thread
  for (int c0 = q * cq; c0 < (q + 1) * cq; c0 += 4) {
    float4 acc = *reinterpret_cast<const float4*>(wsm + K * C1 + c0);
    for (int kk = 0; kk < K; ++kk) {
      const float4 wv = *reinterpret_cast<const float4*>(wsm + kk * C1 + c0);
      const float xv = xin[kk];
      acc.x = fmaf(xv, wv.x, acc.x); acc.y = fmaf(xv, wv.y, acc.y); acc.z = fmaf(xv, wv.z, acc.z); acc.w = fmaf(xv, wv.w, acc.w);
    }
    *reinterpret_cast<float4*>(raw + p * C1 + c0) = acc;
  }
}

// ---- InstanceNorm2d(affine) + LeakyReLU, fp32 in, operand type out ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) enc_instnorm_kernel(const float* __restrict__ raw, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, T* __restrict__ out, int B,
                                                           int H, int W, int Hg, int Wg, int C, int planar, float eps,
                                                           float slope) {
  __shared__ double red[2][8][32];
  __shared__ float stat[2][32];
  const int c = threadIdx.x & 31, pl = threadIdx.x >> 5, c0 = blockIdx.y * 32, b = blockIdx.x;   // images on grid.x (no 65 535 cap)
  const int HW = H * W;   // real pixels; the raw tensor is [B][Hg][Wg][C] with Hg >= H, Wg >= W (padded grid of an odd map)
  const float* src = raw + (size_t)b * Hg * Wg * C + c0 + c;
  auto raw_pix = [&](int p) { return Wg == W ? p : (p / W) * Wg + (p % W); };
  float s = 0.f, s2 = 0.f;
  for (int p = pl; p < HW; p += 8) {
    const float v = src[(size_t)raw_pix(p) * C];
    s += v;
    s2 = fmaf(v, v, s2);
  }
  red[0][pl][c] = (double)s;
  red[1][pl][c] = (double)s2;
  __syncthreads();
  if (pl == 0) {
    double a = 0.0, a2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += red[0][i][c]; a2 += red[1][i][c]; }
    const double mean = a / HW;
    double var = a2 / HW - mean * mean;   // biased variance, as InstanceNorm2d uses
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c0 + c] : 1.f;
    stat[0][c] = g * rstd;
    stat[1][c] = (beta ? beta[c0 + c] : 0.f) - (float)mean * g * rstd;
  }
  __syncthreads();
  const float sc = stat[0][c], sh = stat[1][c];
  const int Hh = (H + 1) >> 1, Wh = (W + 1) >> 1;
  for (int p = pl; p < HW; p += 8) {
    float v = fmaf(src[(size_t)raw_pix(p) * C], sc, sh);
    v = v > 0.f ? v : slope * v;
    size_t o;
    if (planar) {
      const int y = p / W, x = p - y * W;
      o = ((((size_t)((y & 1) * 2 + (x & 1)) * B + b) * Hh + (y >> 1)) * Wh + (x >> 1)) * C + c0 + c;
    } else {
      o = ((size_t)b * HW + p) * C + c0 + c;
    }
    store_t(out + o, v);
  }
}

static int enc_alloc(EncPack* e, void** p, size_t bytes) {
  DAMC_CUDA(cudaMalloc(p, bytes));
  e->allocs.push_back(*p);
  return DAMC_OK;
}

void EncPack::sources(std::vector<HashSrc>& out) const {
  // only the packed convolution weights are derived data; biases / InstanceNorm affines are read from the caller's tensors
  for (int l = 1; l < nlayers; ++l) {
    const EncLayer& y = layers[l];
    out.push_back(HashSrc{y.src.weight, (unsigned long long)y.cout * y.cin * y.k * y.k, 0ull});
  }
}

int EncPack::refill(cudaStream_t stream, const int* dirty) {
  const size_t es = elem_size(precision);
  for (int l = 1; l < nlayers; ++l) {
    EncLayer& y = layers[l];
    // Conv2d weight [Cout,Cin,k,k] read as the ConvTranspose2d weight [Cin_t = Cout, Cout_t = Cin, k, k] of the dgrad plans
    const int mode = y.type == ENC_DOWN ? PK_UP_DGRAD : PK_FIRST_DGRAD;
    const int ntaps = y.type == ENC_DOWN ? 16 : 1;
    const int Cs = y.type == ENC_DOWN ? y.cin : y.k * y.k * y.cin;
    const size_t n = es * (size_t)ntaps * Cs * y.cout;
    if (use_tc) {
      if (!y.w_tc) DAMC_TRY(enc_alloc(this, &y.w_tc, n));
      DAMC_TRY(launch_pack_convt(y.src.weight, y.cout, y.cin, y.k, y.type == ENC_DOWN ? 2 : 1, y.type == ENC_DOWN ? 1 : 0, mode,
                                 0, ntaps, Cs, y.cout, 1, precision, y.w_tc, stream, dirty));
    } else {
      if (!y.w_simt) DAMC_TRY(enc_alloc(this, &y.w_simt, n));
      DAMC_TRY(launch_pack_convt(y.src.weight, y.cout, y.cin, y.k, y.type == ENC_DOWN ? 2 : 1, y.type == ENC_DOWN ? 1 : 0, mode,
                                 0, ntaps, Cs, y.cout, 0, precision, y.w_simt, stream, dirty));
    }
  }
  return DAMC_OK;
}

template <typename T>
static int launch_instnorm(const EncPack* e, const EncLayer& y, const float* raw, void* out, int B, int planar,
                           cudaStream_t s) {
  if (planar && ((y.Hout | y.Wout) & 1))   // odd map: the parity planes keep zeros where no pixel exists
    DAMC_CUDA(cudaMemsetAsync(out, 0, sizeof(T) * 4 * (size_t)B * ((y.Hout + 1) / 2) * ((y.Wout + 1) / 2) * y.cout, s));
  enc_instnorm_kernel<T><<<dim3(B, y.cout / 32), 256, 0, s>>>(raw, y.src.in_weight, y.src.in_bias, (T*)out, B, y.Hout,
                                                               y.Wout, y.Hg, y.Wg, y.cout, planar, e->eps, e->slope);
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}

static int encoder_run(const EncPack* e, const EncWs& w, const float* x, float* xemb, int B, cudaStream_t s) {
  for (int l = 0; l < e->nlayers; ++l) {
    const EncLayer& y = e->layers[l];
    if (y.type == ENC_FIRST) {
      const long long npix = (long long)B * y.Hout * y.Wout;
      const size_t sm = sizeof(float) * ((size_t)y.cin * 9 * y.cout + y.cout);
      enc_first_conv_kernel<<<(unsigned)((npix + 63) / 64), 256, sm, s>>>(x, y.src.weight, y.src.bias, w.raw, npix, y.cin,
                                                                          y.Hout, y.Wout, y.cout);
      DAMC_CUDA(cudaGetLastError());
      count_launch();
    } else {
      GemmPlan p{};
      p.A = w.act[l - 1];
      p.B = B;
      p.N = p.Np = y.cout;
      p.ksplit = 1;
      p.W = y.w_simt; p.Wtc = y.w_tc;
      if (y.type == ENC_DOWN) {   // out[y] = sum_kh in[2y - 1 + kh] W[kh]: parity plane (kh+1)&1, shift -1/0/0/+1
        p.Hm = y.Hg; p.Wm = y.Wg; p.Cs = y.cin; p.ntaps = 16;   // the plane grid (padded for odd maps)
        p.plane_stride = (long long)B * y.Hg * y.Wg * y.cin;
        for (int kh = 0; kh < 4; ++kh)
          for (int kw = 0; kw < 4; ++kw) {
            Tap& t = p.taps[kh * 4 + kw];
            t.plane = (signed char)((((kh + 1) & 1) << 1) | ((kw + 1) & 1));
            t.dy = (signed char)(kh == 0 ? -1 : (kh == 3 ? 1 : 0));
            t.dx = (signed char)(kw == 0 ? -1 : (kw == 3 ? 1 : 0));
            t.pad = 0;
          }
        p.epi.kind = EPI_STORE_F32;
        p.epi.out = w.raw;
        p.epi.nz_out = y.cout;
      } else {                    // k x k map -> 1 x 1: plain GEMM over (kh,kw,c), bias in the epilogue
        p.Hm = 1; p.Wm = 1; p.Cs = y.k * y.k * y.cin; p.ntaps = 1;
        p.taps[0] = Tap{0, 0, 0, 0};
        p.epi.kind = EPI_STORE_F32_BIAS;
        p.epi.bias = y.src.bias;
        p.epi.out = xemb;
        p.epi.nz_out = y.cout;
      }
      profile_mark(s, true);
      const int r = e->use_tc ? launch_gemm_tc(p, e->precision, s) : launch_gemm_simt(p, e->precision, s);
      profile_mark(s, false);
      count_launch();
      DAMC_TRY(r);
    }
    if (l + 1 < e->nlayers) {
      const int planar = e->layers[l + 1].type == ENC_DOWN;
      if (e->precision == DAMC_PREC_FP32) DAMC_TRY(launch_instnorm<float>(e, y, w.raw, w.act[l], B, planar, s));
      else if (e->precision == DAMC_PREC_FP16) DAMC_TRY(launch_instnorm<__half>(e, y, w.raw, w.act[l], B, planar, s));
      else DAMC_TRY(launch_instnorm<__nv_bfloat16>(e, y, w.raw, w.act[l], B, planar, s));
    }
  }
  return DAMC_OK;
}

}  // namespace damc

using namespace damc;

extern "C" int damc_pack_encoder(damc_handle** out, int nlayers, const damc_conv_layer* L, int height, int width,
                                 float negative_slope, float eps, int precision, void* stream) {
  if (!out || !L) DAMC_FAIL(DAMC_ERR_INVALID, "damc_pack_encoder: null argument");
  if (nlayers < 3 || nlayers > 10) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "encoder needs 3..10 Conv2d layers (got %d)", nlayers);
  if (precision != DAMC_PREC_FP32 && !is_tc_precision(precision)) DAMC_FAIL(DAMC_ERR_INVALID, "unknown precision %d", precision);
  EncPack* e = new EncPack();
  e->kind = H_ENC; e->precision = precision; e->nlayers = nlayers; e->slope = negative_slope; e->eps = eps;
  e->nc = L[0].cin; e->H = height; e->W = width;
  e->use_tc = is_tc_precision(precision);
  auto fail = [&](const char* msg, int i) { delete e; set_error("encoder layer %d: %s", i, msg); return DAMC_ERR_UNSUPPORTED; };
  int H = height, W = width;
  e->layers.resize(nlayers);
  for (int i = 0; i < nlayers; ++i) {
    EncLayer& y = e->layers[i];
    const damc_conv_layer& s = L[i];
    if (!s.weight) return fail("null weight", i);
    if (i > 0 && s.cin != L[i - 1].cout) return fail("cin does not match the previous layer's cout", i);
    y.src = s; y.cin = s.cin; y.cout = s.cout; y.k = s.k; y.Hin = H; y.Win = W;
    if (i == 0) {
      if (s.k != 3 || s.stride != 1 || s.pad != 1 || s.cin > 4 || s.cout % 64) return fail("first layer must be Conv2d(nc<=4, 64n, 3, 1, 1)", i);
      y.type = ENC_FIRST; y.Hout = H; y.Wout = W; y.Hg = H; y.Wg = W;
    } else if (i < nlayers - 1) {
      if (s.k != 4 || s.stride != 2 || s.pad != 1) return fail("inner layers must be Conv2d(k=4, s=2, p=1)", i);
      if (H < 2 || W < 2) return fail("k4-s2-p1 layer needs a map of at least 2 x 2", i);
      if (s.cin % 64 || s.cout % 64) return fail("channel counts must be multiples of 64", i);
      y.type = ENC_DOWN; y.Hout = (H - 2) / 2 + 1; y.Wout = (W - 2) / 2 + 1;   // = H/2 for even H
      y.Hg = (H + 1) / 2; y.Wg = (W + 1) / 2;
    } else {
      if (s.stride != 1 || s.pad != 0 || s.k != H || s.k != W) return fail("last layer must reduce the k x k map to 1 x 1 (stride 1, padding 0)", i);
      if ((s.k * s.k * s.cin) % 64 || s.cout % 16) return fail("last layer: k*k*cin must be a multiple of 64 and nemb of 16", i);
      y.type = ENC_LAST; y.Hout = 1; y.Wout = 1; y.Hg = 1; y.Wg = 1;
    }
    if (i < nlayers - 1 && (!s.in_weight || !s.in_bias)) return fail("InstanceNorm2d(affine=True) parameters are required", i);
    H = y.Hout; W = y.Wout;
  }
  e->nemb = L[nlayers - 1].cout;
  if (e->use_tc && !tc_available()) { delete e; DAMC_FAIL(DAMC_ERR_CUDA, "encoder: the tcgen05 engine needs cuTensorMapEncodeTiled from the driver"); }
  int r = e->refill((cudaStream_t)stream, nullptr);
  if (r == DAMC_OK) r = handle_hash_init(e, (cudaStream_t)stream);
  if (r != DAMC_OK) { delete e; return r; }
  *out = e;
  return DAMC_OK;
}

extern "C" size_t damc_encoder_workspace_bytes(const damc_handle* enc, int B) {
  if (!enc || enc->kind != H_ENC || B <= 0) return 0;
  return enc_ws(static_cast<const EncPack*>(enc), B, nullptr).bytes;
}

extern "C" int damc_encoder_forward(const damc_handle* enc, const float* x, float* xemb, int B, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!enc || enc->kind != H_ENC) DAMC_FAIL(DAMC_ERR_INVALID, "damc_encoder_forward: not an encoder handle");
  if (!x || !xemb || B <= 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_encoder_forward: bad arguments");
  const EncPack* e = static_cast<const EncPack*>(enc);
  const EncWs w = enc_ws(e, B, workspace);
  if (!workspace || workspace_bytes < w.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  return encoder_run(e, w, x, xemb, B, (cudaStream_t)stream);
}
