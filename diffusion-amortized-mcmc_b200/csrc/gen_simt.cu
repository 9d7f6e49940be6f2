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
#include "damc_common.cuh"
#include "damc_internal.h"
#include "gen_epilogue.cuh"

namespace damc {

size_t elem_size(int precision) { return precision == DAMC_PREC_BF16 ? 2 : 4; }

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
                                  int Cs, int Np, int nk_layout, T* __restrict__ dst) {
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
                      int Cs, int Np, int nk_layout, int precision, void* dst, cudaStream_t stream) {
  (void)stride; (void)pad;
  const long long total = (long long)ntaps * Cs * Np;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
  if (precision == DAMC_PREC_BF16)
    pack_convt_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                                 reinterpret_cast<__nv_bfloat16*>(dst));
  else
    pack_convt_kernel<float><<<blocks, 256, 0, stream>>>(w, cin, cout, k, mode, cls, ntaps, Cs, Np, nk_layout,
                                                         reinterpret_cast<float*>(dst));
  DAMC_CUDA(cudaGetLastError());
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
  if (precision == DAMC_PREC_BF16)
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<__nv_bfloat16*>(zin), B, nz, nz_p);
  else
    stage_z_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(z, reinterpret_cast<float*>(zin), B, nz, nz_p);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

}  // namespace damc
