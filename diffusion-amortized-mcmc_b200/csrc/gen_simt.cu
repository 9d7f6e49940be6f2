// gen_simt.cu -- CUDA-core implicit GEMM for the generator's transposed convolutions (forward and input-gradient),
// weight packing, and the shared epilogues.
//
// This is the fp32 parity engine (DAMC_PREC_FP32): exact fp32 FMA accumulation, used for the rel-1e-3 comparison with
// the reference's sample_langevin_post_z_with_prior (workspace/src/MCMC.py:48-74), and the on-device cross-check of
// the tcgen05 engine in gen_tc.cu (instantiated for bf16 storage it computes the same products the tensor cores do).
//
// Formulation (see damc_internal.h): every ConvTranspose2d forward (reference diffusion_net.py:26-45) and its gradient
// with respect to the input is a sum over "taps" of dense [M x Cs] x [Cs x N] products where M enumerates (chain, y, x)
// on a fixed pixel grid and a tap only shifts the source window -- no im2col buffer, no strided gathers:
//   k4,s2,p1 forward : 4 output-parity classes, each a 2x2-tap conv on the INPUT grid (K = 4 Cin)
//   k4,s2,p1 dgrad   : 16 taps over the 4 parity planes of dL/dh, stored plane-major             (K = 16 Cout)
//   1x1 -> kxk first layer: a plain GEMM in both directions; last layer (Cout = nc): forward writes dL/dh already
//   im2col'd ("gcol", 64 columns per input pixel) so its dgrad is a plain K=64 GEMM too.
#include <algorithm>
#include <type_traits>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "gen_epilogue.cuh"

namespace damc {

size_t elem_size(int precision) { return is_tc_precision(precision) ? 2 : 4; }   // fp32 and tf32 modes store fp32 containers

// ---- the SIMT kernel ----------------------------------------------------------------------------------------------
template <typename T, int BN>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const GemmPlan p) {
  constexpr int BM = 128, BK = 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int M = p.B * p.Hm * p.Wm;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const T* A = reinterpret_cast<const T*>(p.A);
  const T* W = reinterpret_cast<const T*>(p.W);

  // A-load role: one row, 8 consecutive channels
  const int lr = tid & 127, lk = (tid >> 7) * 8;
  const int lm = m0 + lr;
  const bool lrow_ok = lm < M;
  const int lx = lm % p.Wm, ly = (lm / p.Wm) % p.Hm, lb = lm / (p.Wm * p.Hm);

  const int chunks_per_tap = p.Cs / BK;
  const int nchunks = p.ntaps * chunks_per_tap;
  const int per_split = (nchunks + p.ksplit - 1) / p.ksplit;
  const int kc_begin = blockIdx.z * per_split, kc_end = min(nchunks, kc_begin + per_split);

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[8];
  constexpr int BQ = BK * BN / 4;             // float4 quads in the B tile
  constexpr int QPT = (BQ + 255) / 256;       // quads per thread
  float rb[QPT][4];

  auto fetch = [&](int kc) {
    const int t = kc / chunks_per_tap, c0 = (kc - t * chunks_per_tap) * BK;
    const Tap tp = p.taps[t];
    const int yy = ly + tp.dy, xx = lx + tp.dx;
    if (lrow_ok && yy >= 0 && yy < p.Hm && xx >= 0 && xx < p.Wm) {
      const T* src = A + (long long)tp.plane * p.plane_stride + (((long long)lb * p.Hm + yy) * p.Wm + xx) * p.Cs + c0 + lk;
      load8(src, ra);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ra[i] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
      const int qi = tid + q * 256;
      if (qi < BQ) {
        const int kr = qi / (BN / 4), nq = (qi - kr * (BN / 4)) * 4;
        if (n0 + nq < p.Np) {
          load4(W + ((long long)t * p.Cs + c0 + kr) * p.Np + n0 + nq, rb[q]);
        } else {
          rb[q][0] = rb[q][1] = rb[q][2] = rb[q][3] = 0.f;
        }
      }
    }
  };

  if (kc_begin < kc_end) fetch(kc_begin);
  for (int kc = kc_begin; kc < kc_end; ++kc) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[lk + i][lr] = ra[i];
#pragma unroll
    for (int q = 0; q < QPT; ++q) {
      const int qi = tid + q * 256;
      if (qi < BQ) {
        const int kr = qi / (BN / 4), nq = (qi - kr * (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[kr][nq]) = make_float4(rb[q][0], rb[q][1], rb[q][2], rb[q][3]);
      }
    }
    __syncthreads();
    if (kc + 1 < kc_end) fetch(kc + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], bv[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[k][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  float loss_acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int x = m % p.Wm, y = (m / p.Wm) % p.Hm, b = m / (p.Wm * p.Hm);
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < p.N) epilogue_elem<T>(p, blockIdx.z, m, b, y, x, n, acc[i][j], loss_acc);
    }
  }
  if (p.epi.kind == EPI_FWD_LAST && p.epi.loss != nullptr) {
    loss_acc = warp_sum(loss_acc);
    if ((tid & 31) == 0 && loss_acc != 0.f) atomicAdd(p.epi.loss, loss_acc);
  }
}

template <typename T>
static int launch_simt_t(const GemmPlan& p, cudaStream_t stream) {
  const int M = p.B * p.Hm * p.Wm;
  const int gm = ceil_div(M, 128);
  if (p.Np <= 16) {
    gemm_simt_kernel<T, 16><<<dim3(gm, ceil_div(p.Np, 16), p.ksplit), 256, 0, stream>>>(p);
  } else if (p.Np <= 64) {
    gemm_simt_kernel<T, 64><<<dim3(gm, ceil_div(p.Np, 64), p.ksplit), 256, 0, stream>>>(p);
  } else {
    gemm_simt_kernel<T, 128><<<dim3(gm, ceil_div(p.Np, 128), p.ksplit), 256, 0, stream>>>(p);
  }
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

int launch_gemm_simt(const GemmPlan& p, int precision, cudaStream_t stream) {
  if (p.Cs % 16 != 0 || p.Np % 4 != 0) DAMC_FAIL(DAMC_ERR_INVALID, "SIMT GEMM needs Cs%%16==0, Np%%4==0 (Cs=%d Np=%d)", p.Cs, p.Np);
  if (precision == DAMC_PREC_FP16) return launch_simt_t<__half>(p, stream);
  if (precision == DAMC_PREC_TF32) return launch_simt_t<tf32_t>(p, stream);   // tf32-rounded storage, fp32 FMA (cross-check engine)
  return precision == DAMC_PREC_BF16 ? launch_simt_t<__nv_bfloat16>(p, stream) : launch_simt_t<float>(p, stream);
}

// ---- weight packing: PyTorch ConvTranspose2d [Cin,Cout,k,k] fp32 -> per-tap GEMM operands ---------------------------
__device__ __forceinline__ long long pack_src_index(int mode, int cls, int t, int c, int n, int cin, int cout, int k) {
  int ci = -1, co = -1, kh = 0, kw = 0;
  switch (mode) {
    case PK_FIRST_FWD: {  // c = ci ; n = (kh,kw,co)
      ci = c; co = n % cout; const int s = n / cout; kh = s / k; kw = s % k;
      if (s >= k * k) return -1;
    } break;
    case PK_FIRST_DGRAD: {  // c = (kh,kw,co) ; n = ci
      ci = n; co = c % cout; const int s = c / cout; kh = s / k; kw = s % k;
      if (s >= k * k) return -1;
    } break;
    case PK_UP_FWD: {  // class (py,px), tap (ty,tx): py==0 -> kh in {1,3}; py==1 -> kh in {0,2}
      const int py = cls >> 1, px = cls & 1, ty = t >> 1, tx = t & 1;
      kh = py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 0 : 2);
      kw = px == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 0 : 2);
      ci = c; co = n;
    } break;
    case PK_UP_DGRAD: { kh = t >> 2; kw = t & 3; co = c; ci = n; } break;
    case PK_SAME_FWD: { kh = t / 3; kw = t % 3; ci = c; co = n; } break;
    case PK_LAST_FWD_SCATTER: {  // c = ci ; n = (kh*k+kw)*cout + co
      ci = c; co = n % cout; const int s = n / cout; kh = s / k; kw = s % k;
      if (s >= k * k) return -1;
    } break;
    case PK_LAST_DGRAD_COL: {  // c = (kh*k+kw)*4 + ch ; n = ci
      const int s = c >> 2; co = c & 3; ci = n; kh = s / k; kw = s % k;
      if (s >= k * k) return -1;
    } break;
  }
  if (ci < 0 || ci >= cin || co < 0 || co >= cout) return -1;
  return (((long long)ci * cout + co) * k + kh) * k + kw;
}

template <typename T>
__global__ void pack_convt_kernel(const float* __restrict__ w, int cin, int cout, int k, int mode, int cls, int ntaps,
                                  int Cs, int Np, int nk_layout, T* __restrict__ dst, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const long long total = (long long)ntaps * Cs * Np;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int t, c, n;
    if (nk_layout) { c = (int)(i % Cs); n = (int)((i / Cs) % Np); t = (int)(i / ((long long)Cs * Np)); }
    else { n = (int)(i % Np); c = (int)((i / Np) % Cs); t = (int)(i / ((long long)Cs * Np)); }
    const long long s = pack_src_index(mode, cls, t, c, n, cin, cout, k);
    store_t(dst + i, s >= 0 ? w[s] : 0.f);
  }
}

int launch_pack_convt(const float* w, int cin, int cout, int k, int stride, int pad, int mode, int cls, int ntaps,
                      int Cs, int Np, int nk_layout, int precision, void* dst, cudaStream_t stream, const int* dirty) {
  (void)stride; (void)pad;
  const long long total = (long long)ntaps * Cs * Np;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (precision == DAMC_PREC_FP16)
    pack_convt_kernel<__half><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                          reinterpret_cast<__half*>(dst), dirty);
  else if (precision == DAMC_PREC_BF16)
    pack_convt_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                                 reinterpret_cast<__nv_bfloat16*>(dst), dirty);
  else if (precision == DAMC_PREC_TF32)
    pack_convt_kernel<tf32_t><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                          reinterpret_cast<tf32_t*>(dst), dirty);
  else
    pack_convt_kernel<float><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                         reinterpret_cast<float*>(dst), dirty);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

// ---- last layer, scatter form: col2im + bias + tanh + likelihood gradient + im2col'd dL/dh ----------------------------------
// Y[b][m_in][(kh*k+kw)*nc + co] holds the per-input-pixel contributions of ConvTranspose2d (reference diffusion_net.py:
// 42-46); out[oy,ox,co] = bias + sum over (kh,kw) with (oy+p-kh, ox+p-kw) divisible by the stride and inside the input.
// One CTA owns a block of IRB input rows of one image: it needs dL/dh on the output rows those inputs touch (a one-row
// halo, recomputed) and Y on the input rows that feed those outputs (a further halo) -- all staged in shared memory.
struct FinishArgs {
  const float* Y; const float* bias; const float* x; float* xhat; float* loss; void* gcol;
  int Hi, Wi, Ho, Wo, k, stride, pad, nc, np, irb, nblk;
  float inv_sigma2, gscale;
};

struct FinishRange { int iy0, iy1, oa, ob, ia, ib; };
__host__ __device__ inline int fdiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }   // floor(a/b), b > 0
__host__ __device__ inline FinishRange finish_range(int blk, int irb, int Hi, int Ho, int k, int s, int p) {
  FinishRange r;
  r.iy0 = blk * irb;
  r.iy1 = r.iy0 + irb < Hi ? r.iy0 + irb : Hi;
  const int oa = r.iy0 * s - p, ob = (r.iy1 - 1) * s - p + k - 1;
  r.oa = oa < 0 ? 0 : oa;
  r.ob = ob > Ho - 1 ? Ho - 1 : ob;
  const int ia = fdiv(r.oa + p - k + 1 + s - 1, s), ib = fdiv(r.ob + p, s);   // ceil / floor
  r.ia = ia < 0 ? 0 : ia;
  r.ib = ib > Hi - 1 ? Hi - 1 : ib;
  return r;
}

// KK / SS: compile-time kernel size and stride (3,1 | 4,2) so the tap arithmetic has no integer divisions; 0,0 = generic.
template <typename T, int KK, int SS>
__global__ void __launch_bounds__(384) last_finish_kernel(const FinishArgs aa) {
  FinishArgs a = aa;
  if (KK) { a.k = KK; a.stride = SS; a.pad = 1; }
  extern __shared__ __align__(16) float fsm[];
  const int b = blockIdx.x / a.nblk, blk = blockIdx.x - b * a.nblk, tid = threadIdx.x;
  const FinishRange R = finish_range(blk, a.irb, a.Hi, a.Ho, a.k, a.stride, a.pad);
  const int nyr = (R.ib - R.ia + 1) * a.Wi, ngr = (R.ob - R.oa + 1) * a.Wo;
  const int pitch = a.np + 1;                   // odd pitch: consecutive pixels (threads) fall into different banks
  float* Ys = fsm;                              // [(ib-ia+1)*Wi][np+1]
  float* gS = fsm + (((size_t)nyr * pitch + 3) & ~(size_t)3);   // [(ob-oa+1)*Wo][4]
  __shared__ float red[12];
  {
    const float4* src = reinterpret_cast<const float4*>(a.Y + ((size_t)b * a.Hi + R.ia) * a.Wi * a.np);
    const int q4 = a.np / 4, n4 = nyr * q4;
    const int q4_shift = (q4 & (q4 - 1)) == 0 ? __ffs(q4) - 1 : -1;
    for (int i0 = tid; i0 < n4; i0 += 4 * blockDim.x) {  // 4 independent 16-byte loads in flight per thread
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        v[u] = i < n4 ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < n4) {
          const int row = q4_shift >= 0 ? (i >> q4_shift) : i / q4, c4 = (i - row * q4) * 4;
          float* d = Ys + (size_t)row * pitch + c4;
          d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
        }
      }
    }
  }
  __syncthreads();
  const int own0 = R.iy0 * a.stride, own1 = R.iy1 * a.stride;  // output rows this block reports (x_hat, loss)
  float loss_acc = 0.f;
  for (int pix = tid; pix < ngr; pix += blockDim.x) {
    const int oyl = pix / a.Wo, ox = pix - oyl * a.Wo, oy = R.oa + oyl;
    float h[4] = {0.f, 0.f, 0.f, 0.f}, xv[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < a.nc; ++c) h[c] = a.bias[c];
    if (a.x)   // issued before the smem sums so that the global-load latency hides behind them
      for (int c = 0; c < a.nc; ++c) xv[c] = __ldg(a.x + (((size_t)b * a.nc + c) * a.Ho + oy) * a.Wo + ox);
    const int kk = KK ? KK : a.k, ss = KK ? SS : a.stride;
#pragma unroll
    for (int kh = 0; kh < (KK ? KK : 4); ++kh) {
      if (kh >= kk) break;
      const int ny = oy + a.pad - kh;
      if (ny < 0 || ny % ss) continue;
      const int iy = ny / ss;
      if (iy >= a.Hi) continue;
#pragma unroll
      for (int kw = 0; kw < (KK ? KK : 4); ++kw) {
        if (kw >= kk) break;
        const int nx = ox + a.pad - kw;
        if (nx < 0 || nx % ss) continue;
        const int ix = nx / ss;
        if (ix >= a.Wi) continue;
        const float* yr = Ys + (size_t)((iy - R.ia) * a.Wi + ix) * pitch + (kh * kk + kw) * a.nc;
        for (int c = 0; c < a.nc; ++c) h[c] += yr[c];
      }
    }
    const bool own = oy >= own0 && oy < own1;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < a.nc; ++c) {
      const float xh = tanhf(h[c]);
      const size_t xi = (((size_t)b * a.nc + c) * a.Ho + oy) * a.Wo + ox;
      if (a.xhat && own) a.xhat[xi] = xh;
      if (a.x) {
        const float r = xh - xv[c];
        g[c] = r * (a.inv_sigma2 * a.gscale) * (1.f - xh * xh);
        if (own) loss_acc += 0.5f * a.inv_sigma2 * r * r;
      }
    }
    *reinterpret_cast<float4*>(gS + (size_t)pix * 4) = make_float4(g[0], g[1], g[2], g[3]);
  }
  if (a.x == nullptr) return;
  if (a.loss != nullptr) {
    loss_acc = warp_sum(loss_acc);
    if ((tid & 31) == 0) red[tid >> 5] = loss_acc;
  }
  __syncthreads();
  if (a.loss != nullptr && tid == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    if (t != 0.f) atomicAdd(a.loss, t);
  }
  // im2col'd gradient rows for the last layer's dgrad: gcol[m_in][(kh*k+kw)*4 + c] = g[(iy*s-p+kh, ix*s-p+kw)][c]
  constexpr int EPC = 16 / (int)sizeof(T);   // entries per 16-byte chunk (8 bf16 | 4 fp32)
  constexpr int CH = 64 / EPC;               // chunks per 64-entry row
  T* gc = reinterpret_cast<T*>(a.gcol) + ((size_t)b * a.Hi + R.iy0) * a.Wi * 64;
  const int nrows = (R.iy1 - R.iy0) * a.Wi;
  for (int i = tid; i < nrows * CH; i += blockDim.x) {
    const int m = i / CH, j = i - m * CH;
    const int iy = R.iy0 + m / a.Wi, ix = m % a.Wi;
    float vals[EPC];
#pragma unroll
    for (int s2 = 0; s2 < EPC / 4; ++s2) {
      const int slot = j * (EPC / 4) + s2;
      float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
      const int kk2 = KK ? KK : a.k, ss2 = KK ? SS : a.stride;
      if (slot < kk2 * kk2) {
        const int kh = slot / kk2, kw = slot - kh * kk2;
        const int oy = iy * ss2 - a.pad + kh, ox = ix * ss2 - a.pad + kw;
        if (oy >= 0 && oy < a.Ho && ox >= 0 && ox < a.Wo)
          gv = *reinterpret_cast<const float4*>(gS + (size_t)((oy - R.oa) * a.Wo + ox) * 4);
      }
      vals[4 * s2] = gv.x; vals[4 * s2 + 1] = gv.y; vals[4 * s2 + 2] = gv.z; vals[4 * s2 + 3] = gv.w;
    }
    T* dst = gc + (size_t)m * 64 + j * EPC;
    if constexpr (sizeof(T) == 4) {
      if constexpr (std::is_same<T, tf32_t>::value) {
#pragma unroll
        for (int q = 0; q < 4; ++q) vals[q] = round_tf32(vals[q]);
      }
      *reinterpret_cast<float4*>(dst) = make_float4(vals[0], vals[1], vals[2], vals[3]);
    } else {
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if constexpr (sizeof(T) == 2 && !std::is_same<T, __half>::value) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(vals[2 * q], vals[2 * q + 1]);
          w[q] = *reinterpret_cast<const uint32_t*>(&hh);
        } else {
          const __half2 hh = __floats2half2_rn(vals[2 * q], vals[2 * q + 1]);
          w[q] = *reinterpret_cast<const uint32_t*>(&hh);
        }
      }
      *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

static size_t finish_smem_for(const GenLayer& y, int irb) {
  size_t mx = 0;
  const int nblk = ceil_div(y.Hin, irb);
  for (int blk = 0; blk < nblk; ++blk) {
    const FinishRange R = finish_range(blk, irb, y.Hin, y.Hout, y.k, y.stride, y.pad);
    mx = std::max(mx, sizeof(float) * ((size_t)(R.ib - R.ia + 1) * y.Win * (y.np_sc + 1) + 4 + (size_t)(R.ob - R.oa + 1) * y.Wout * 4));
  }
  return mx;
}
static int finish_irb(const GenLayer& y) {  // input rows per CTA: ~4+ blocks per image, <= 56 KB so that 4 CTAs share an SM
  int irb = std::max(1, y.Hin / 4);
  while (irb > 1 && finish_smem_for(y, irb) > 56 * 1024) --irb;
  return irb;
}
size_t last_finish_smem(const GenLayer& y) { return finish_smem_for(y, finish_irb(y)); }

int launch_last_finish(const GenLayer& y, int precision, const float* Y, int B, const float* x, float* xhat,
                       float inv_sigma2, float gscale, float* loss, void* gcol, cudaStream_t stream) {
  FinishArgs a{};
  a.Y = Y; a.bias = y.bias; a.x = x; a.xhat = xhat; a.loss = loss; a.gcol = gcol;
  a.Hi = y.Hin; a.Wi = y.Win; a.Ho = y.Hout; a.Wo = y.Wout; a.k = y.k; a.stride = y.stride; a.pad = y.pad;
  a.nc = y.cout; a.np = y.np_sc; a.inv_sigma2 = inv_sigma2; a.gscale = gscale;
  a.irb = finish_irb(y);
  a.nblk = ceil_div(y.Hin, a.irb);
  const size_t smem = finish_smem_for(y, a.irb);
  // one thread per gradient pixel of a block when that fits (k3-s1-p1 at 32 x 32: 10 rows x 32 = 320 pixels -- with 256
  // threads a quarter of the block would idle at the barrier while 64 threads do a second pixel)
  int max_ngr = 0;
  for (int blk = 0; blk < a.nblk; ++blk) {
    const FinishRange R = finish_range(blk, a.irb, y.Hin, y.Hout, y.k, y.stride, y.pad);
    max_ngr = std::max(max_ngr, (R.ob - R.oa + 1) * y.Wout);
  }
  const int threads = (max_ngr > 256 && max_ngr <= 384) ? (max_ngr + 31) / 32 * 32 : 256;
  auto go = [&](auto kern) -> int {
    DAMC_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B * a.nblk, threads, smem, stream>>>(a);
    DAMC_CUDA(cudaGetLastError());
    return DAMC_OK;
  };
  const bool same = y.k == 3 && y.stride == 1 && y.pad == 1, up = y.k == 4 && y.stride == 2 && y.pad == 1;
  if (precision == DAMC_PREC_FP16) {
    if (same) return go(last_finish_kernel<__half, 3, 1>);
    if (up) return go(last_finish_kernel<__half, 4, 2>);
    return go(last_finish_kernel<__half, 0, 0>);
  }
  if (precision == DAMC_PREC_BF16) {
    if (same) return go(last_finish_kernel<__nv_bfloat16, 3, 1>);
    if (up) return go(last_finish_kernel<__nv_bfloat16, 4, 2>);
    return go(last_finish_kernel<__nv_bfloat16, 0, 0>);
  }
  if (precision == DAMC_PREC_TF32) {
    if (same) return go(last_finish_kernel<tf32_t, 3, 1>);
    if (up) return go(last_finish_kernel<tf32_t, 4, 2>);
    return go(last_finish_kernel<tf32_t, 0, 0>);
  }
  if (same) return go(last_finish_kernel<float, 3, 1>);
  if (up) return go(last_finish_kernel<float, 4, 2>);
  return go(last_finish_kernel<float, 0, 0>);
}

// sum over one chain's image of (x_hat - x)^2, fixed reduction order (thread-strided partial sums, shuffle tree, warps in order)
__global__ void __launch_bounds__(256) sqerr_kernel(const float* __restrict__ xhat, const float* __restrict__ x, int n,
                                                    float* __restrict__ sq_part) {
  const size_t base = (size_t)blockIdx.x * n;
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float r = xhat[base + i] - __ldg(x + base + i);
    acc = fmaf(r, r, acc);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    sq_part[blockIdx.x] = t;
  }
}

int launch_sqerr(const float* xhat, const float* x, int B, int n, float* sq_part, cudaStream_t stream) {
  sqerr_kernel<<<B, 256, 0, stream>>>(xhat, x, n, sq_part);
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}

template <typename T>
__global__ void stage_z_kernel(const float* __restrict__ z, T* __restrict__ zin, int B, int nz, int nz_p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nz_p) return;
  const int b = i / nz_p, k = i - b * nz_p;
  store_t(zin + i, k < nz ? z[(size_t)b * nz + k] : 0.f);
}

int launch_stage_z(const float* z, void* zin, int B, int nz, int nz_p, int precision, cudaStream_t stream) {
  const int n = B * nz_p;
  if (precision == DAMC_PREC_FP16)
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<__half*>(zin), B, nz, nz_p);
  else if (precision == DAMC_PREC_BF16)
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<__nv_bfloat16*>(zin), B, nz, nz_p);
  else if (precision == DAMC_PREC_TF32)
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<tf32_t*>(zin), B, nz, nz_p);
  else
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<float*>(zin), B, nz, nz_p);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

}  // namespace damc
