// denoiser.cu -- DAMC ancestral sampler: the T reverse steps of the latent diffusion amortizer.
//
// Replaces the loop of _netQ_U.forward (reference workspace/src/diffusion_net.py:597-620) with Q.p = Diffusion_UnetA
// (:463-533), ConcatSquashLinearSkipCtx (:417-445), SinusoidalPosEmb (:447-461) and the scalar algebra of
// workspace/src/diffusion_helper_func.py:36-70.
//
// Restructuring (verified against the reference by tests/golden):
//   * ctx = [temb, xemb]  =>  Linear(SiLU(ctx)) = Wc[:, :ntemb] SiLU(temb) + Wc[:, ntemb:] SiLU(xemb) + bc.
//     The xemb half is constant over the T steps  -> one GEMM per call   ("cx", [B, sum dout]).
//     The temb half is constant over the batch    -> one table per call  ("ct", [T, sum dout]).
//   * every per-step scalar (lambda_t, lambda_s, alpha, r, var) depends on the step only -> 4 coefficients per step.
//   * one fused kernel per reverse step runs the whole 7-layer network for a tile of chains with activations in
//     shared memory; the main/skip and gate/bias matrix pairs are interleaved so each activation load feeds two FMAs.
// All arithmetic is fp32 (the sampler is ill-conditioned: x12.8 amplification at lambda = -5.1).
#include <math.h>

#include <algorithm>
#include <vector>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

// ---- packing ------------------------------------------------------------------------------------------------------
__global__ void pack_interleave_T(const float* __restrict__ A, const float* __restrict__ Bm, int rows, int cols,
                                  float* __restrict__ dst, const int* __restrict__ dirty) {  // A,B: [rows][cols] -> dst[cols][rows][2]
  if (gate_clean(dirty)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int k = i / rows, o = i - k * rows;
  dst[2 * i] = A[(size_t)o * cols + k];
  dst[2 * i + 1] = Bm[(size_t)o * cols + k];
}
__global__ void pack_ctx_T(const float* __restrict__ Wc, int dout, int ntemb, int nxemb, int coff, int csum,
                           float* __restrict__ dT, float* __restrict__ dX, const int* __restrict__ dirty) {  // Wc [dout][ntemb+nxemb]
  if (gate_clean(dirty)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int w = ntemb + nxemb;
  if (i >= dout * w) return;
  const int o = i / w, k = i - o * w;
  if (k < ntemb) dT[(size_t)k * csum + coff + o] = Wc[i];
  else dX[(size_t)(k - ntemb) * csum + coff + o] = Wc[i];
}

__device__ __forceinline__ float silu(float v) { return v / (1.f + expf(-v)); }
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// ---- per-call: cx[b][:] = WcT_x^T SiLU(xemb[b]) + bc   (thread = output column, CTA = 8 chains) ---------------------
__global__ void __launch_bounds__(256) den_hoist_kernel(const float* __restrict__ xemb, const float* __restrict__ WcT_x,
                                                        const float* __restrict__ bc, int B, int nxemb, int csum,
                                                        float* __restrict__ cx) {
  constexpr int TB = 8;
  extern __shared__ float sx[];  // [nxemb][TB]
  const int b0 = blockIdx.y * TB;
  for (int i = threadIdx.x; i < nxemb * TB; i += blockDim.x) {
    const int k = i / TB, c = i - k * TB;
    sx[i] = (b0 + c < B) ? silu(xemb[(size_t)(b0 + c) * nxemb + k]) : 0.f;
  }
  __syncthreads();
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= csum) return;
  float acc[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) acc[c] = bc[o];
  for (int k = 0; k < nxemb; ++k) {
    const float w = WcT_x[(size_t)k * csum + o];
    const float4 s0 = *reinterpret_cast<const float4*>(&sx[k * TB]), s1 = *reinterpret_cast<const float4*>(&sx[k * TB + 4]);
    acc[0] = fmaf(w, s0.x, acc[0]); acc[1] = fmaf(w, s0.y, acc[1]); acc[2] = fmaf(w, s0.z, acc[2]); acc[3] = fmaf(w, s0.w, acc[3]);
    acc[4] = fmaf(w, s1.x, acc[4]); acc[5] = fmaf(w, s1.y, acc[5]); acc[6] = fmaf(w, s1.z, acc[6]); acc[7] = fmaf(w, s1.w, acc[7]);
  }
#pragma unroll
  for (int c = 0; c < TB; ++c)
    if (b0 + c < B) cx[(size_t)(b0 + c) * csum + o] = acc[c];
}

// ---- per-call: ct[t][:] = WcT_t^T SiLU(time_mlp(u(lambda_t)))   (CTA = one step) -------------------------------------
// Mirrors the reference's fp32 op order: u = atan(exp(-clamp(l)/2)) / (pi/2)  (diffusion_net.py:506);
// SinusoidalPosEmb with max_time=1: x = 1000 u ; arg = x * exp(k * -(ln 1e4/(half-1)))  (:454-460).
__global__ void __launch_bounds__(256) den_time_kernel(const float* __restrict__ logsnr,
                                                       const float* __restrict__ tw1, const float* __restrict__ tb1,
                                                       const float* __restrict__ tw2, const float* __restrict__ tb2,
                                                       const float* __restrict__ WcT_t, int ntemb, int csum,
                                                       float* __restrict__ ct) {
  extern __shared__ float sm[];  // pe[ntemb], h[ntemb], s[ntemb]
  float* pe = sm;
  float* h = sm + ntemb;
  float* s = h + ntemb;
  const int t = blockIdx.x;
  const float l = fminf(fmaxf(logsnr[t], -20.f), 20.f);
  const float u = atanf(expf(-0.5f * l)) / 1.5707963267948966f;
  const float xt = u * 1000.0f;
  const int half = ntemb / 2;
  const float dec = (float)(9.210340371976184 / (double)(half - 1));  // ln(10000)/(half-1), rounded as torch does
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float arg = xt * expf((float)k * -dec);
    pe[k] = sinf(arg);
    pe[half + k] = cosf(arg);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ntemb; o += blockDim.x) {
    float acc = tb1[o];
    for (int k = 0; k < ntemb; ++k) acc = fmaf(tw1[(size_t)o * ntemb + k], pe[k], acc);
    h[o] = silu(acc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < ntemb; o += blockDim.x) {
    float acc = tb2[o];
    for (int k = 0; k < ntemb; ++k) acc = fmaf(tw2[(size_t)o * ntemb + k], h[k], acc);
    s[o] = silu(acc);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < csum; o += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < ntemb; ++k) acc = fmaf(WcT_t[(size_t)k * csum + o], s[k], acc);
    ct[(size_t)t * csum + o] = acc;
  }
}

// ---- persistent T-step kernel: fused network + reverse update, weights streamed through a TMA bulk-copy ring ---------------
// One CTA owns TM chains for ALL T reverse steps (z never leaves shared memory).  The per-step weights (5.9 MB fp32 at
// CIFAR-10 shape) do not fit on an SM, so a producer warp streams them from L2 in consumption order with
// cp.async.bulk (32 KB chunks, mbarrier full/empty ring) while 8 consumer warps (thread = output feature) run the
// mat-vecs: acc[c] += w[k][o] * act[k][c].  The (main, skip) and (gate, hyper-bias) matrix pairs are interleaved as
// float2 so each weight fetch feeds two FMAs per chain.
constexpr int DS_CHUNK_BYTES = 32 * 1024;
constexpr int DS_CONSUMERS = 256;

struct DenStreamArgs {
  const float* wstream;   // per step: for L: Wgb rows (padded) then Wms rows (padded), each row [dout][2]
  const float* bias3[DEN_LAYERS];
  int din[DEN_LAYERS], dout[DEN_LAYERS], coff[DEN_LAYERS];
  int rows_gb[DEN_LAYERS], rows_ms[DEN_LAYERS], R[DEN_LAYERS];  // padded row counts, rows per chunk
  const float* Bp;
  const float* cx;     // [B][csum]
  const float* ct;     // [T][csum], row i = reverse-step index
  const float* coef;   // [T][8] in execution order: c_pred, c_eps, c_zt, c_x, c_std, last
  int csum, nz, residual, B, T;
  float* z;            // [B][nz] in/out (unless eps_out)
  float* eps_out;      // non-null: single eps prediction (T = 1), z untouched
  const float* noise;  // [T-1][B][nz] or null
  int use_philox, ring;
  uint64_t seed, chain0;
};

__device__ __forceinline__ uint32_t ds_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ds_mbar_init(uint32_t bar, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void ds_mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void ds_mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void ds_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (spin > (1u << 27)) __trap();  // a protocol bug must fail loudly, not hang the GPU
  }
}
__device__ __forceinline__ void ds_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ds_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(DS_CONSUMERS) : "memory"); }

template <int TM>
__global__ void __launch_bounds__(DS_CONSUMERS + 32, 1) den_stream_kernel(const DenStreamArgs a) {
  extern __shared__ __align__(128) float sm[];
  const int nz = a.nz, half = nz / 2, tid = threadIdx.x;
  float* ring = sm;                                                  // [ring][32 KB]
  float* zs = ring + (size_t)a.ring * (DS_CHUNK_BYTES / 4);          // [nz][TM]
  float* hbuf = zs + nz * TM;                                        // [DEN_MAXW][TM] layer input (concat buffer)
  float* cbuf = hbuf + DEN_MAXW * TM;                                // [256][TM]
  float* skip0 = cbuf + 256 * TM;
  float* skip1 = skip0 + a.dout[0] * TM;
  float* skip2 = skip1 + a.dout[1] * TM;
  float* obuf = skip2 + a.dout[2] * TM;                              // [256][TM] layer output
  __shared__ __align__(8) unsigned long long bars[16];               // full[ring], empty[ring]
  const uint32_t bar0 = ds_smem_u32(bars);
  auto full_bar = [&](int s2) { return bar0 + 8u * s2; };
  auto empty_bar = [&](int s2) { return bar0 + 8u * (a.ring + s2); };
  if (tid == 0) {
    for (int i = 0; i < a.ring; ++i) { ds_mbar_init(full_bar(i), 1); ds_mbar_init(empty_bar(i), DS_CONSUMERS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int nsteps = a.eps_out ? 1 : a.T;

  if (tid >= DS_CONSUMERS) {
    // ===================== producer warp: stream the weights, chunk by chunk, step after step =====================
    if (tid == DS_CONSUMERS) {
      uint32_t cidx = 0;
      for (int st = 0; st < nsteps; ++st) {
        const char* src = reinterpret_cast<const char*>(a.wstream);
        for (int L = 0; L < DEN_LAYERS; ++L) {
          const uint32_t bytes = (uint32_t)a.R[L] * a.dout[L] * 8u;
          const int nch = (a.rows_gb[L] + a.rows_ms[L]) / a.R[L];
          for (int c = 0; c < nch; ++c, ++cidx) {
            const int slot = cidx % a.ring;
            ds_mbar_wait(empty_bar(slot), ((cidx / a.ring) & 1u) ^ 1u);
            ds_mbar_expect(full_bar(slot), bytes);
            ds_bulk_g2s(ds_smem_u32(ring) + (uint32_t)slot * DS_CHUNK_BYTES, src, bytes, full_bar(slot));
            src += bytes;
          }
        }
      }
    }
    return;
  }

  // ===================== consumers =====================
  const int b0 = blockIdx.x * TM, o = tid, lane = tid & 31;
  uint32_t cidx = 0;
  // accA[c] += w.x * vec[k][c], accB[c] += w.y * vec[k][c] over the streamed rows of one matrix
  auto consume = [&](int rows_p, int R, int dout, const float* vec, float (&accA)[TM], float (&accB)[TM]) {
    for (int k0 = 0; k0 < rows_p; k0 += R, ++cidx) {
      const int slot = cidx % a.ring;
      ds_mbar_wait(full_bar(slot), (cidx / a.ring) & 1u);
      if (o < dout) {
        const float2* w = reinterpret_cast<const float2*>(ring + (size_t)slot * (DS_CHUNK_BYTES / 4)) + o;
#pragma unroll 4
        for (int r = 0; r < R; ++r) {
          const float2 wv = w[(size_t)r * dout];
          const float* v = vec + (size_t)(k0 + r) * TM;
#pragma unroll
          for (int c4 = 0; c4 < TM / 4; ++c4) {
            const float4 x4 = *reinterpret_cast<const float4*>(v + 4 * c4);
            accA[4 * c4 + 0] = fmaf(wv.x, x4.x, accA[4 * c4 + 0]); accB[4 * c4 + 0] = fmaf(wv.y, x4.x, accB[4 * c4 + 0]);
            accA[4 * c4 + 1] = fmaf(wv.x, x4.y, accA[4 * c4 + 1]); accB[4 * c4 + 1] = fmaf(wv.y, x4.y, accB[4 * c4 + 1]);
            accA[4 * c4 + 2] = fmaf(wv.x, x4.z, accA[4 * c4 + 2]); accB[4 * c4 + 2] = fmaf(wv.y, x4.z, accB[4 * c4 + 2]);
            accA[4 * c4 + 3] = fmaf(wv.x, x4.w, accA[4 * c4 + 3]); accB[4 * c4 + 3] = fmaf(wv.y, x4.w, accB[4 * c4 + 3]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ds_mbar_arrive(empty_bar(slot));
    }
  };
  // out[b][o] = (W h + b)[o] * sigmoid((Wg c + bg)[o]) + (Wb c)[o] + (Ws h + bs)[o]       (diffusion_net.py:439-445)
  auto css_layer = [&](int L, const float* ctrow, const float* hin, float* hout) {
    const int dout = a.dout[L];
    for (int i = tid; i < dout * TM; i += DS_CONSUMERS) {  // c = SiLU(cx + ct)
      const int k = i / TM, c = i - k * TM;
      const int b = min(b0 + c, a.B - 1);
      cbuf[i] = silu(a.cx[(size_t)b * a.csum + a.coff[L] + k] + ctrow[a.coff[L] + k]);
    }
    ds_consumer_sync();
    float g[TM], hb[TM], m[TM], sk[TM];
    const float bm = o < dout ? a.bias3[L][o] : 0.f, bs = o < dout ? a.bias3[L][dout + o] : 0.f;
    const float bg = o < dout ? a.bias3[L][2 * dout + o] : 0.f;
#pragma unroll
    for (int c = 0; c < TM; ++c) { g[c] = bg; hb[c] = 0.f; m[c] = bm; sk[c] = bs; }
    consume(a.rows_gb[L], a.R[L], dout, cbuf, g, hb);
    consume(a.rows_ms[L], a.R[L], dout, hin, m, sk);
    ds_consumer_sync();  // everyone is done reading hin / cbuf
    if (o < dout) {
#pragma unroll
      for (int c4 = 0; c4 < TM / 4; ++c4) {
        float r4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) r4[e] = m[4 * c4 + e] * sigmoidf_(g[4 * c4 + e]) + hb[4 * c4 + e] + sk[4 * c4 + e];
        *reinterpret_cast<float4*>(&hout[o * TM + c4 * 4]) = make_float4(r4[0], r4[1], r4[2], r4[3]);
      }
    }
    ds_consumer_sync();
  };
  auto lrelu_copy = [&](const float* src, float* dst, int n) {  // dst = leaky_relu(src, 0.01)
    for (int i = tid; i < n * TM; i += DS_CONSUMERS) { const float v = src[i]; dst[i] = v > 0.f ? v : 0.01f * v; }
  };

  for (int i = tid; i < (DEN_MAXW + 256) * TM; i += DS_CONSUMERS) hbuf[i] = 0.f;  // padded rows must stay finite
  for (int i = tid; i < nz * TM; i += DS_CONSUMERS) {
    const int c = i / nz, k = i - c * nz;  // coalesced over k
    zs[k * TM + c] = (b0 + c < a.B) ? a.z[(size_t)(b0 + c) * nz + k] : 0.f;
  }
  ds_consumer_sync();
  float* skips[3] = {skip0, skip1, skip2};
  for (int st = 0; st < nsteps; ++st) {
    const int irev = a.eps_out ? 0 : a.T - 1 - st;
    const float* ctrow = a.ct + (size_t)irev * a.csum;
    // input embedding: [sin(2 pi zB), cos(2 pi zB), z]                                       (diffusion_net.py:497-499)
    for (int i = tid; i < half * TM; i += DS_CONSUMERS) {
      const int j = i / TM, c = i - j * TM;
      float acc = 0.f;
      for (int k = 0; k < nz; ++k) acc = fmaf(zs[k * TM + c], __ldg(a.Bp + (size_t)k * half + j), acc);
      const float pr = 6.283185307179586f * acc;
      hbuf[j * TM + c] = sinf(pr);
      hbuf[(half + j) * TM + c] = cosf(pr);
    }
    for (int i = tid; i < nz * TM; i += DS_CONSUMERS) hbuf[2 * half * TM + i] = zs[i];
    ds_consumer_sync();
    for (int L = 0; L < 3; ++L) {  // in_layers, skip saved pre-activation (:514-519)
      css_layer(L, ctrow, hbuf, skips[L]);
      lrelu_copy(skips[L], hbuf, a.dout[L]);
      ds_consumer_sync();
    }
    css_layer(3, ctrow, hbuf, obuf);  // mid layer (:520)
    for (int L = 4; L < 7; ++L) {     // out_layers on leaky_relu(cat[out, skip]) (:524-527)
      const int wprev = a.dout[L - 1], ws = a.dout[6 - L];
      lrelu_copy(obuf, hbuf, wprev);
      lrelu_copy(skips[6 - L], hbuf + wprev * TM, ws);
      ds_consumer_sync();
      css_layer(L, ctrow, hbuf, obuf);
    }
    // eps = z + out (residual) ; reverse update                                             (:530-531, :610-620)
    const float* cf = a.coef + (size_t)st * 8;
    const float c_pred = cf[0], c_eps = cf[1], c_zt = cf[2], c_x = cf[3], c_std = cf[4];
    const bool last = cf[5] != 0.f;
    for (int i = tid; i < nz * TM; i += DS_CONSUMERS) {
      const int c = i / nz, k = i - c * nz;  // coalesced global accesses over k
      const int b = b0 + c;
      const int si = k * TM + c;
      const float zt = zs[si];
      const float eps = a.residual ? zt + obuf[si] : obuf[si];
      if (a.eps_out != nullptr) {
        if (b < a.B) a.eps_out[(size_t)b * nz + k] = eps;
        continue;
      }
      const float pred = c_pred * (zt - eps * c_eps);
      float zn = pred;
      if (!last) {
        zn = c_zt * zt + c_x * pred;
        if (c_std != 0.f && b < a.B) {
          const float e = a.noise ? a.noise[((size_t)st * a.B + b) * nz + k]
                                  : (a.use_philox ? philox_normal1(a.seed, a.chain0 + b, (uint64_t)st, (uint32_t)k) : 0.f);
          zn = fmaf(c_std, e, zn);
        }
      }
      zs[si] = zn;
    }
    ds_consumer_sync();
  }
  if (a.eps_out == nullptr)
    for (int i = tid; i < nz * TM; i += DS_CONSUMERS) {
      const int c = i / nz, k = i - c * nz;
      if (b0 + c < a.B) a.z[(size_t)(b0 + c) * nz + k] = zs[k * TM + c];
    }
}

template <int TM>
static size_t den_stream_smem(const DenPack* d, int ring) {
  return (size_t)ring * DS_CHUNK_BYTES +
         sizeof(float) * TM * ((size_t)d->nz + DEN_MAXW + 256 + d->dout[0] + d->dout[1] + d->dout[2] + 256) + 256;
}

// host-side scalar algebra of one reverse step, in double (diffusion_helper_func.py:36-70)
static void reverse_coeffs(double lt, double ls, int var_type, float* c_pred, float* c_eps, float* c_zt, float* c_x,
                           float* c_std) {
  auto sigm = [](double v) { return 1.0 / (1.0 + exp(-v)); };
  *c_pred = (float)sqrt(1.0 + exp(-lt));
  *c_eps = (float)(1.0 / sqrt(1.0 + exp(lt)));
  const double alpha_st = sqrt((1.0 + exp(-lt)) / (1.0 + exp(-ls)));
  const double alpha_s = sqrt(sigm(ls));
  const double r = exp(lt - ls), omr = -expm1(lt - ls);
  *c_zt = (float)(r * alpha_st);
  *c_x = (float)(omr * alpha_s);
  double var;
  if (var_type == 1) {
    var = omr * sigm(-lt);
  } else {
    const double a_t = sigm(lt), a_s = sigm(ls);
    var = (1.0 - a_s) / (1.0 - a_t) * (1.0 - a_t / a_s);
  }
  *c_std = (float)sqrt(var > 0.0 ? var : 0.0);
}

static void fill_stream_args(const DenPack* d, DenStreamArgs* a) {
  for (int i = 0; i < DEN_LAYERS; ++i) {
    a->bias3[i] = d->bias3[i];
    a->din[i] = d->din[i]; a->dout[i] = d->dout[i]; a->coff[i] = d->coff[i];
    a->rows_gb[i] = d->rows_gb[i]; a->rows_ms[i] = d->rows_ms[i]; a->R[i] = d->R[i];
  }
  a->wstream = d->wstream; a->Bp = d->Bp; a->csum = d->csum; a->nz = d->nz; a->residual = d->residual;
}

static int launch_den_stream(const DenPack* d, DenStreamArgs& a, int B, cudaStream_t s) {
  // few chains: 8 per CTA (more CTAs, 4-deep ring); many chains: 16 per CTA (half the weight traffic per chain)
  if (B <= 4 * 148) {  // latency regime: spread the chains over as many SMs as possible
    a.ring = 4;
    const size_t smem = den_stream_smem<4>(d, a.ring);
    DAMC_CUDA(cudaFuncSetAttribute(den_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    den_stream_kernel<4><<<ceil_div(B, 4), DS_CONSUMERS + 32, smem, s>>>(a);
  } else if (B <= 8 * 148) {
    a.ring = 4;
    const size_t smem = den_stream_smem<8>(d, a.ring);
    DAMC_CUDA(cudaFuncSetAttribute(den_stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    den_stream_kernel<8><<<ceil_div(B, 8), DS_CONSUMERS + 32, smem, s>>>(a);
  } else {
    a.ring = 3;
    const size_t smem = den_stream_smem<16>(d, a.ring);
    DAMC_CUDA(cudaFuncSetAttribute(den_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    den_stream_kernel<16><<<ceil_div(B, 16), DS_CONSUMERS + 32, smem, s>>>(a);
  }
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

static int run_hoist_and_time(const DenPack* d, const float* xemb, int B, int T, const float* host_logsnr, float* cx,
                              float* ct, float* dlog, cudaStream_t s, bool hoist = true) {
  DAMC_CUDA(cudaMemcpyAsync(dlog, host_logsnr, sizeof(float) * T, cudaMemcpyHostToDevice, s));
  den_time_kernel<<<T, 256, sizeof(float) * 3 * d->ntemb, s>>>(dlog, d->tw1, d->tb1, d->tw2, d->tb2, d->WcT_t,
                                                             d->ntemb, d->csum, ct);
  DAMC_CUDA(cudaGetLastError());
  if (!hoist) return DAMC_OK;   // the caller forms cx on the tensor cores (den_seq_hoist)
  const size_t sm = sizeof(float) * d->nxemb * 8;
  DAMC_CUDA(cudaFuncSetAttribute(den_hoist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  den_hoist_kernel<<<dim3(ceil_div(d->csum, 256), ceil_div(B, 8)), 256, sm, s>>>(xemb, d->WcT_x, d->bc, B, d->nxemb,
                                                                                  d->csum, cx);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

DenWs den_ws(const DenPack* d, int B, int T, int precision, void* base) {
  DenWs w{};
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += align_up(n, 256); return base ? (float*)((char*)base + r) : nullptr; };
  w.cx = take(sizeof(float) * (size_t)B * d->csum);
  w.ct = take(sizeof(float) * (size_t)T * d->csum);
  w.dlog = take(sizeof(float) * (size_t)(T + 1));
  w.coef = take(sizeof(float) * 8 * (size_t)T);
  for (int i = 0; i < DEN_LAYERS; ++i)
    w.A[i] = is_tc_precision(precision) ? (void*)take(2 * (size_t)B * (d->din[i] + d->dout[i])) : nullptr;
  w.zbuf = is_tc_precision(precision) ? take(sizeof(float) * (size_t)B * d->nz) : nullptr;
  w.seed_dev = is_tc_precision(precision) ? (unsigned long long*)take(16) : nullptr;
  if (is_tc_precision(precision) && den_seq_shape_ok(d)) {
    const size_t bp = align_up((size_t)B, 128);
    w.skip[0] = take(2 * bp * d->dout[0]);
    w.skip[1] = take(2 * bp * d->dout[1]);
    w.G = take(4 * bp * d->csum * (size_t)den_seq_window(B, T, d->csum));
    w.nbuf = take(4 * bp * d->nz * (size_t)den_seq_window(B, T, d->csum));
    w.zT = take(4 * bp * d->nz);
    w.xr = take(4 * (size_t)B * d->nxemb);
  }
  w.base = base;
  w.bytes = o;
  return w;
}

DenPack::~DenPack() {
  if (slab) cudaFree(slab);
  for (DenTcPack* t : tc) den_tc_free(t);
}

void DenPack::sources(std::vector<HashSrc>& out) const {
  const damc_denoiser_desc* h = &src;
  auto add = [&](const float* p, size_t n) { out.push_back(HashSrc{p, (unsigned long long)n, 0ull}); };
  add(h->time_w1, (size_t)ntemb * ntemb); add(h->time_b1, ntemb); add(h->time_w2, (size_t)ntemb * ntemb); add(h->time_b2, ntemb);
  add(h->Bproj, (size_t)nz * (nz / 2));
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const size_t di = din[i], dn = dout[i];
    add(h->W[i], dn * di); add(h->b[i], dn); add(h->Wc[i], dn * (ntemb + nxemb)); add(h->bc[i], dn);
    add(h->Wg[i], dn * dn); add(h->bg[i], dn); add(h->Wb[i], dn * dn); add(h->Ws[i], dn * di); add(h->bs[i], dn);
  }
}

int DenPack::refill(cudaStream_t s, const int* dirty) {
  const damc_denoiser_desc* h = &src;
  const int nt = ntemb;
  DAMC_TRY(launch_gated_copy(tw1, h->time_w1, (size_t)nt * nt, dirty, s));
  DAMC_TRY(launch_gated_copy(tb1, h->time_b1, nt, dirty, s));
  DAMC_TRY(launch_gated_copy(tw2, h->time_w2, (size_t)nt * nt, dirty, s));
  DAMC_TRY(launch_gated_copy(tb2, h->time_b2, nt, dirty, s));
  DAMC_TRY(launch_gated_copy(Bp, h->Bproj, (size_t)nz * (nz / 2), dirty, s));
  // (the row padding of the weight stream is zeroed once at pack time; the packing kernels never touch it)
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int di = din[i], dn = dout[i], off = coff[i];
    pack_interleave_T<<<ceil_div(di * dn, 256), 256, 0, s>>>(h->W[i], h->Ws[i], dn, di, Wms[i], dirty);
    pack_interleave_T<<<ceil_div(dn * dn, 256), 256, 0, s>>>(h->Wg[i], h->Wb[i], dn, dn, Wgb[i], dirty);
    pack_ctx_T<<<ceil_div(dn * (nt + nxemb), 256), 256, 0, s>>>(h->Wc[i], dn, nt, nxemb, off, csum, WcT_t, WcT_x, dirty);
    DAMC_TRY(launch_gated_copy(bias3[i], h->b[i], dn, dirty, s));
    DAMC_TRY(launch_gated_copy(bias3[i] + dn, h->bs[i], dn, dirty, s));
    DAMC_TRY(launch_gated_copy(bias3[i] + 2 * dn, h->bg[i], dn, dirty, s));
    DAMC_TRY(launch_gated_copy(bc + off, h->bc[i], dn, dirty, s));
  }
  DAMC_CUDA(cudaGetLastError());
  for (int prec = 0; prec < 3; ++prec)
    if (tc[prec]) DAMC_TRY(den_tc_refill(this, prec, s, dirty));
  return DAMC_OK;
}

static int check_den_precision(int precision) {
  if (precision != DAMC_PREC_FP32 && !is_tc_precision(precision)) DAMC_FAIL(DAMC_ERR_INVALID, "denoiser: unknown precision %d", precision);
  return DAMC_OK;
}

}  // namespace damc

using namespace damc;

extern "C" int damc_pack_denoiser(damc_handle** out, const damc_denoiser_desc* h, void* stream) {
  if (!out || !h) DAMC_FAIL(DAMC_ERR_INVALID, "damc_pack_denoiser: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (h->nz < 2 || h->nz % 2 || h->nz > 256 || h->ntemb % 2 || h->ntemb < 4 || h->ntemb > 1024)
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser: need even nz <= 256 and even ntemb (nz=%d ntemb=%d)", h->nz, h->ntemb);
  int csum = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    if (h->dim_out[i] > DEN_THREADS || h->dim_in[i] > DEN_MAXW || h->dim_out[i] < 1)
      DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser layer %d: %d -> %d exceeds the fused kernel's limits (in<=%d, out<=%d)", i, h->dim_in[i], h->dim_out[i], DEN_MAXW, DEN_THREADS);
    csum += h->dim_out[i];
  }
  // wiring of Diffusion_UnetA (diffusion_net.py:481-495)
  const bool ok = h->dim_in[0] == 2 * h->nz && h->dim_in[1] == h->dim_out[0] && h->dim_in[2] == h->dim_out[1] &&
                  h->dim_in[3] == h->dim_out[2] && h->dim_in[4] == h->dim_out[3] + h->dim_out[2] &&
                  h->dim_in[5] == h->dim_out[4] + h->dim_out[1] && h->dim_in[6] == h->dim_out[5] + h->dim_out[0] &&
                  h->dim_out[6] == h->nz;
  if (!ok) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser: layer widths do not match Diffusion_UnetA's skip wiring");
  DenPack* d = new DenPack();
  d->kind = H_DEN; d->nz = h->nz; d->nxemb = h->nxemb; d->ntemb = h->ntemb; d->residual = h->residual; d->csum = csum;
  size_t nstream = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int dn = h->dim_out[i];
    d->R[i] = std::max(1, std::min(64, 4096 / dn));
    d->rows_gb[i] = ceil_div(dn, d->R[i]) * d->R[i];
    d->rows_ms[i] = ceil_div(h->dim_in[i], d->R[i]) * d->R[i];
    if (d->rows_ms[i] > DEN_MAXW || d->rows_gb[i] > 256) { delete d; DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser layer %d: padded widths exceed the kernel's buffers", i); }
    nstream += 2 * (size_t)(d->rows_gb[i] + d->rows_ms[i]) * dn;
  }
  d->stream_floats = nstream;
  size_t total = 2 * ((size_t)h->ntemb * h->ntemb + h->ntemb) + (size_t)h->nz * (h->nz / 2) +
                 (size_t)(h->ntemb + h->nxemb) * csum + csum + 64 + nstream;
  for (int i = 0; i < DEN_LAYERS; ++i) total += 3 * (size_t)h->dim_out[i];
  if (cudaMalloc(&d->slab, total * sizeof(float)) != cudaSuccess) { delete d; DAMC_FAIL(DAMC_ERR_CUDA, "damc_pack_denoiser: cudaMalloc failed"); }
  float* p = d->slab;
  auto take = [&](size_t n) { float* r = p; p += n; return r; };
  const int nt = h->ntemb;
  d->tw1 = take((size_t)nt * nt); d->tb1 = take(nt); d->tw2 = take((size_t)nt * nt); d->tb2 = take(nt);
  d->Bp = take((size_t)h->nz * (h->nz / 2));
  d->WcT_t = take((size_t)nt * csum); d->WcT_x = take((size_t)h->nxemb * csum); d->bc = take(csum);
  int off = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int di = h->dim_in[i], dn = h->dim_out[i];
    d->din[i] = di; d->dout[i] = dn; d->coff[i] = off;
    d->bias3[i] = take(3 * (size_t)dn);
    off += dn;
  }
  p = d->slab + align_up((size_t)(p - d->slab), 32);  // bulk copies need 16-byte aligned sources
  d->wstream = p;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    d->Wgb[i] = take(2 * (size_t)d->rows_gb[i] * d->dout[i]);
    d->Wms[i] = take(2 * (size_t)d->rows_ms[i] * d->dout[i]);
  }
  d->src = *h;
  if (cudaMemsetAsync(d->wstream, 0, sizeof(float) * d->stream_floats, s) != cudaSuccess) { delete d; DAMC_FAIL(DAMC_ERR_CUDA, "damc_pack_denoiser: memset failed"); }
  int rr = d->refill(s, nullptr);
  if (rr == DAMC_OK) rr = handle_hash_init(d, s);
  if (rr != DAMC_OK) { delete d; return rr; }
  if (den_stream_smem<16>(d, 3) > 227 * 1024) { delete d; DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser: kernel needs %zu B shared memory", den_stream_smem<16>(d, 3)); }
  *out = d;
  return DAMC_OK;
}

extern "C" size_t damc_denoise_workspace_bytes(const damc_handle* den, int B, int T, int precision) {
  if (!den || den->kind != H_DEN || B <= 0 || T <= 0) return 0;
  return den_ws(static_cast<const DenPack*>(den), B, T, precision, nullptr).bytes;
}

extern "C" int damc_denoise(const damc_handle* den, float* z, const float* xemb, int B, int T, const float* host_logsnr,
                            int var_type, int with_noise, const float* noise, uint64_t seed, uint64_t chain0,
                            int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!den || den->kind != H_DEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_denoise: not a denoiser handle");
  if (!z || !xemb || !host_logsnr || B <= 0 || T < 2) DAMC_FAIL(DAMC_ERR_INVALID, "damc_denoise: bad arguments (need T >= 2)");
  if (var_type != 0 && var_type != 1) DAMC_FAIL(DAMC_ERR_INVALID, "damc_denoise: var_type must be 0 ('small') or 1 ('large')");
  DAMC_TRY(check_den_precision(precision));
  const DenPack* d = static_cast<const DenPack*>(den);
  cudaStream_t s = (cudaStream_t)stream;
  const DenWs w = den_ws(d, B, T, precision, workspace);
  if (!workspace || workspace_bytes < w.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  const bool tc_hoist = den_seq_hoist_usable(d, precision, B);
  DAMC_TRY(run_hoist_and_time(d, xemb, B, T, host_logsnr, w.cx, w.ct, w.dlog, s, !tc_hoist));
  if (tc_hoist) DAMC_TRY(den_seq_hoist(d, precision, w, xemb, B, s));
  std::vector<float> coef(8 * (size_t)T, 0.f);
  for (int k = 0, i = T - 1; i >= 0; --i, ++k) {
    float* c = &coef[8 * (size_t)k];
    const double lt = host_logsnr[i], ls = host_logsnr[i > 0 ? i - 1 : 0];
    reverse_coeffs(lt, ls, var_type, &c[0], &c[1], &c[2], &c[3], &c[4]);
    if (!with_noise) c[4] = 0.f;
    c[5] = i == 0 ? 1.f : 0.f;
  }
  if (is_tc_precision(precision))
    return den_tc_run(d, precision, w, z, nullptr, B, T, T, coef.data(), noise, noise == nullptr, seed, chain0, s);
  DAMC_CUDA(cudaMemcpyAsync(w.coef, coef.data(), sizeof(float) * coef.size(), cudaMemcpyHostToDevice, s));
  DenStreamArgs a{};
  fill_stream_args(d, &a);
  a.cx = w.cx; a.ct = w.ct; a.coef = w.coef; a.B = B; a.T = T; a.z = z; a.eps_out = nullptr;
  a.noise = noise; a.use_philox = noise == nullptr; a.seed = seed; a.chain0 = chain0;
  DAMC_TRY(launch_den_stream(d, a, B, s));
  count_launch(3);
  return DAMC_OK;
}

extern "C" int damc_denoiser_eps(const damc_handle* den, const float* z, const float* xemb, float logsnr, float* eps_out,
                                 int B, int precision, void* workspace, size_t workspace_bytes, void* stream) {
  if (!den || den->kind != H_DEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_denoiser_eps: not a denoiser handle");
  if (!z || !xemb || !eps_out || B <= 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_denoiser_eps: bad arguments");
  DAMC_TRY(check_den_precision(precision));
  const DenPack* d = static_cast<const DenPack*>(den);
  cudaStream_t s = (cudaStream_t)stream;
  const DenWs w = den_ws(d, B, 1, precision, workspace);
  if (!workspace || workspace_bytes < w.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", w.bytes, workspace_bytes);
  DAMC_TRY(run_hoist_and_time(d, xemb, B, 1, &logsnr, w.cx, w.ct, w.dlog, s));
  if (is_tc_precision(precision)) {
    const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    return den_tc_run(d, precision, w, const_cast<float*>(z), eps_out, B, 1, 1, zero8, nullptr, 0, 0, 0, s);
  }
  DAMC_CUDA(cudaMemsetAsync(w.coef, 0, sizeof(float) * 8, s));
  DenStreamArgs a{};
  fill_stream_args(d, &a);
  a.cx = w.cx; a.ct = w.ct; a.coef = w.coef; a.B = B; a.T = 1; a.z = const_cast<float*>(z); a.eps_out = eps_out;
  DAMC_TRY(launch_den_stream(d, a, B, s));
  count_launch(3);
  return DAMC_OK;
}
