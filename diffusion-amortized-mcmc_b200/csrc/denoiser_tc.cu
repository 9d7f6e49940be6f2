// denoiser_tc.cu -- DAMC reverse steps on the tcgen05 engine (DAMC_PREC_BF16 / DAMC_PREC_FP16).
//
// Replaces the 37 cuBLAS GEMMs + ~40 elementwise launches PyTorch issues per reverse step of _netQ_U.forward (reference
// workspace/src/diffusion_net.py:597-620; Q.p = Diffusion_UnetA :463-533, ConcatSquashLinearSkipCtx :417-445) by
// 1 + 7 launches:
//   * den_prep_kernel (fp32 CUDA cores): the input embedding [sin(2 pi zB), cos(2 pi zB), z] (:497-499; the phase z.B is
//     O(10) radians, so it is formed in fp32 and only the result is rounded to the operand type) and the ctx activations
//     c_L = SiLU(cx + ct[t]) of all 7 layers, written straight into the layers' operand rows;
//   * one convgemm_tc_kernel launch per ConcatSquashLinearSkipCtx layer.  The four Linears of a layer become ONE GEMM:
//     operand row = [h (din) | c_L (dout)]; every N tile holds the four terms of BN/4 output features as column blocks
//     [gate | hyper-bias | main | skip].  An h k-block multiplies only the (main, skip) weight rows, a c k-block only the
//     (gate, hyper-bias) rows (half-width MMAs into the matching half of the TMEM accumulator), and the epilogue forms
//         out = (main + b) * sigmoid(gate + bg) + hyper_bias + skip + bs                             (:439-445)
//     writes leaky_relu(out, 0.01) as the next layer's operand (and the U-net skip copy), or -- last layer -- applies
//     eps = z + out and the reverse update of z (:610-620) in fp32.
// z, the schedule coefficients, the hoisted ctx projections and the update stay fp32; only GEMM operands are 16-bit.
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

// ---- packing: [4*dout][din+dout]; rows of N tile t (bn rows): [gate Q | hyper-bias Q | main Q | skip Q], Q = bn/4 -------
template <typename T>
__global__ void pack_den_blocks(const float* __restrict__ W, const float* __restrict__ Ws, const float* __restrict__ Wg,
                                const float* __restrict__ Wb, int din, int dout, int bn, T* __restrict__ dst,
                                const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const int Kt = din + dout, Q = bn >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 4ll * dout * Kt) return;
  const int row = (int)(i / Kt), k = (int)(i - (long long)row * Kt);
  const int tile = row / bn, r = row - tile * bn;
  const int q = r / Q, n = tile * Q + (r - q * Q);   // q: 0 gate, 1 hyper-bias, 2 main, 3 skip ; n: output feature
  float v = 0.f;   // the (gate, hyper-bias) x h and (main, skip) x c blocks are never loaded; keep them zero anyway
  if (q == 0) { if (k >= din) v = Wg[(size_t)n * dout + (k - din)]; }
  else if (q == 1) { if (k >= din) v = Wb[(size_t)n * dout + (k - din)]; }
  else if (q == 2) { if (k < din) v = W[(size_t)n * din + k]; }
  else { if (k < din) v = Ws[(size_t)n * din + k]; }
  dst[i] = T(v);
}
__global__ void pack_den_bias4(const float* __restrict__ b, const float* __restrict__ bs, const float* __restrict__ bg,
                               int dout, float* __restrict__ dst, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= dout) return;
  reinterpret_cast<float4*>(dst)[n] = make_float4(bg[n], 0.f, b[n], bs[n]);
}

void den_tc_free(DenTcPack* t) {
  if (!t) return;
  if (t->gexec) cudaGraphExecDestroy(t->gexec);
  if (t->cap_stream) cudaStreamDestroy(t->cap_stream);
  if (t->slab) cudaFree(t->slab);
  den_seq_free(t->seq);
  delete t;
}

static const int kDenBn[DEN_NBN] = {256, 128, 64};
int den_tc_ensure(const DenPack* d, int precision, cudaStream_t s);
static int den_bn_index(int bn) { return bn == 256 ? 0 : bn == 128 ? 1 : 2; }

static int den_tc_pack_variant(const DenPack* d, int precision, int v, cudaStream_t s, const int* dirty = nullptr) {
  DenTcPack* t = d->tc[precision];
  const damc_denoiser_desc* h = &d->src;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int di = d->din[i], dn = d->dout[i];
    const long long n = 4ll * dn * (di + dn);
    const int blocks = (int)((n + 255) / 256);
    if (precision == DAMC_PREC_FP16)
      pack_den_blocks<__half><<<blocks, 256, 0, s>>>(h->W[i], h->Ws[i], h->Wg[i], h->Wb[i], di, dn, kDenBn[v], (__half*)t->Wq[v][i], dirty);
    else
      pack_den_blocks<__nv_bfloat16><<<blocks, 256, 0, s>>>(h->W[i], h->Ws[i], h->Wg[i], h->Wb[i], di, dn, kDenBn[v], (__nv_bfloat16*)t->Wq[v][i], dirty);
  }
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

int den_tc_pack_bn(const DenPack* d, int precision, int v, cudaStream_t s) {
  DAMC_TRY(den_tc_ensure(d, precision, s));
  DenTcPack* t = d->tc[precision];
  if (!t->live[v]) { t->live[v] = true; DAMC_TRY(den_tc_pack_variant(d, precision, v, s)); }
  return DAMC_OK;
}

int den_tc_refill(const DenPack* d, int precision, cudaStream_t s, const int* dirty) {
  DenTcPack* t = d->tc[precision];
  if (!t) return DAMC_OK;
  const damc_denoiser_desc* h = &d->src;
  for (int v = 0; v < DEN_NBN; ++v)
    if (t->live[v]) DAMC_TRY(den_tc_pack_variant(d, precision, v, s, dirty));
  for (int i = 0; i < DEN_LAYERS; ++i)
    pack_den_bias4<<<ceil_div(d->dout[i], 128), 128, 0, s>>>(h->b[i], h->bs[i], h->bg[i], d->dout[i], t->bias4[i], dirty);
  DAMC_CUDA(cudaGetLastError());
  return den_seq_refill(d, precision, s, dirty);
}

int den_tc_ensure(const DenPack* d, int precision, cudaStream_t s) {
  if (!is_tc_precision(precision)) DAMC_FAIL(DAMC_ERR_INVALID, "denoiser: precision %d is not a tensor-core mode", precision);
  if (d->tc[precision]) return DAMC_OK;
  if (!tc_available()) DAMC_FAIL(DAMC_ERR_CUDA, "denoiser: the tcgen05 engine needs cuTensorMapEncodeTiled from the driver");
  if (d->nz % 4) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser (tensor-core mode): nz %d must be a multiple of 4", d->nz);
  size_t wbytes = 0, bbytes = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    if (d->din[i] % 64 || d->dout[i] % 64)
      DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser (tensor-core mode): layer %d widths %d -> %d must be multiples of 64; use precision fp32",
                i, d->din[i], d->dout[i]);
    wbytes += align_up(2 * 4 * (size_t)d->dout[i] * (d->din[i] + d->dout[i]), 256);
    bbytes += align_up(16 * (size_t)d->dout[i], 256);
  }
  DenTcPack* t = new DenTcPack();
  if (cudaMalloc(&t->slab, DEN_NBN * wbytes + bbytes) != cudaSuccess) { delete t; DAMC_FAIL(DAMC_ERR_CUDA, "denoiser: cudaMalloc(%zu) failed", DEN_NBN * wbytes + bbytes); }
  char* p = (char*)t->slab;
  for (int v = 0; v < DEN_NBN; ++v)
    for (int i = 0; i < DEN_LAYERS; ++i) {
      t->Wq[v][i] = p;
      p += align_up(2 * 4 * (size_t)d->dout[i] * (d->din[i] + d->dout[i]), 256);
    }
  for (int i = 0; i < DEN_LAYERS; ++i) {
    t->bias4[i] = (float*)p;
    p += align_up(16 * (size_t)d->dout[i], 256);
  }
  d->tc[precision] = t;
  return den_tc_refill(d, precision, s);
}

// ---- per-step operand preparation -------------------------------------------------------------------------------------
struct DenPrepArgs {
  const float* z;       // [B][nz]
  const float* Bp;      // [nz][nz/2]
  const float* cx;      // [B][csum]
  const float* ctrow;   // [csum]   (this step's temb half of the ctx pre-activations)
  void* A[DEN_LAYERS];
  int ld[DEN_LAYERS], din[DEN_LAYERS], dout[DEN_LAYERS], coff[DEN_LAYERS];
  int B, nz, csum, fp16;
};

__device__ __forceinline__ float den_silu(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ uint16_t den_cvt(bool fp16, float v) {
  if (fp16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

// One launch, two block roles:
//   blocks [0, nb_embed)  : input embedding of EMB_CHAINS chains -- p.B staged in smem, 8 chains per thread so each weight read
//                           feeds 8 FMAs; phase, sin and cos in fp32                                (diffusion_net.py:497-499)
//   blocks [nb_embed, ..) : ctx activations c_L = SiLU(cx + ct) of CTX_CHAINS chains, float4 in / 4 x 16-bit out, written into
//                           the [din, din+dout) slice of each layer's operand rows                  (:426-433, hoisted halves)
constexpr int EMB_CHAINS = 32, EMB_PITCH = 36, CTX_CHAINS = 8;  // pitch 36: 16-byte aligned rows, 4-way (not 32-way) conflicts on the transposing store
__global__ void __launch_bounds__(256) den_prep_kernel(const DenPrepArgs a, int nb_embed) {
  extern __shared__ __align__(16) float sm[];
  const int nz = a.nz, half = nz >> 1;
  const bool fp16 = a.fp16 != 0;
  // programmatic dependent launch: the first layer's GEMM may set up now; this kernel waits for the previous step's
  // last layer (which wrote z) before reading anything
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if ((int)blockIdx.x < nb_embed) {
    float* Bs = sm;               // [nz][half]
    float* zs = sm + nz * half;   // [nz][EMB_PITCH]  (transposed: the 8 chains of a thread are two float4 reads)
    const int b0 = blockIdx.x * EMB_CHAINS;
    const int nb = min(EMB_CHAINS, a.B - b0);
    for (int i = threadIdx.x; i < nz * half / 4; i += blockDim.x)
      reinterpret_cast<float4*>(Bs)[i] = __ldg(reinterpret_cast<const float4*>(a.Bp) + i);
    for (int i = threadIdx.x; i < EMB_CHAINS * nz; i += blockDim.x) {
      const int c = i / nz, k = i - c * nz;   // coalesced over k
      zs[k * EMB_PITCH + c] = c < nb ? a.z[(size_t)(b0 + c) * nz + k] : 0.f;
    }
    __syncthreads();
    uint16_t* A0 = reinterpret_cast<uint16_t*>(a.A[0]);
    for (int i = threadIdx.x; i < (EMB_CHAINS / 8) * half; i += blockDim.x) {
      const int cg = i / half, j = i - cg * half;
      const float* zr = zs + cg * 8;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
      for (int k = 0; k < nz; ++k) {
        const float w = Bs[k * half + j];
        const float4 z0 = *reinterpret_cast<const float4*>(zr + k * EMB_PITCH);
        const float4 z1 = *reinterpret_cast<const float4*>(zr + k * EMB_PITCH + 4);
        acc[0] = fmaf(z0.x, w, acc[0]); acc[1] = fmaf(z0.y, w, acc[1]); acc[2] = fmaf(z0.z, w, acc[2]); acc[3] = fmaf(z0.w, w, acc[3]);
        acc[4] = fmaf(z1.x, w, acc[4]); acc[5] = fmaf(z1.y, w, acc[5]); acc[6] = fmaf(z1.z, w, acc[6]); acc[7] = fmaf(z1.w, w, acc[7]);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int b = b0 + cg * 8 + c;
        if (b < a.B) {
          float sn, cs;
          sincosf(6.283185307179586f * acc[c], &sn, &cs);
          uint16_t* row = A0 + (size_t)b * a.ld[0];
          row[j] = den_cvt(fp16, sn);
          row[half + j] = den_cvt(fp16, cs);
        }
      }
    }
    for (int i = threadIdx.x; i < nb * (nz / 2); i += blockDim.x) {  // the raw z slice, two elements per thread
      const int c = i / (nz / 2), k2 = (i - c * (nz / 2)) * 2;
      const uint32_t o = (uint32_t)den_cvt(fp16, zs[k2 * EMB_PITCH + c]) | ((uint32_t)den_cvt(fp16, zs[(k2 + 1) * EMB_PITCH + c]) << 16);
      *reinterpret_cast<uint32_t*>(A0 + (size_t)(b0 + c) * a.ld[0] + 2 * half + k2) = o;
    }
    return;
  }
  const int b0 = ((int)blockIdx.x - nb_embed) * CTX_CHAINS;
  const int nb = min(CTX_CHAINS, a.B - b0);
  const int g4 = a.csum >> 2;
  for (int g = threadIdx.x; g < g4; g += blockDim.x) {   // this thread's 4 ctx columns: same layer / offset for every chain
    const int col = g << 2;
    int L = 0;
#pragma unroll
    for (int l = 1; l < DEN_LAYERS; ++l) L += col >= a.coff[l];
    const float4 t4 = __ldg(reinterpret_cast<const float4*>(a.ctrow + col));
    uint16_t* dst = reinterpret_cast<uint16_t*>(a.A[L]) + a.din[L] + (col - a.coff[L]);
    const int ld = a.ld[L];
    float4 x4[CTX_CHAINS];
#pragma unroll
    for (int c = 0; c < CTX_CHAINS; ++c)   // all loads in flight before the first use
      x4[c] = c < nb ? *reinterpret_cast<const float4*>(a.cx + (size_t)(b0 + c) * a.csum + col) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < CTX_CHAINS; ++c) {
      if (c >= nb) break;
      const uint2 o = make_uint2(
          (uint32_t)den_cvt(fp16, den_silu(x4[c].x + t4.x)) | ((uint32_t)den_cvt(fp16, den_silu(x4[c].y + t4.y)) << 16),
          (uint32_t)den_cvt(fp16, den_silu(x4[c].z + t4.z)) | ((uint32_t)den_cvt(fp16, den_silu(x4[c].w + t4.w)) << 16));
      *reinterpret_cast<uint2*>(dst + (size_t)(b0 + c) * ld) = o;
    }
  }
}

// ---- driver -----------------------------------------------------------------------------------------------------------
// issue the nsteps x (1 + 7) launches on stream s (directly, or into a stream capture)
static int den_tc_issue(const DenPack* d, int precision, const DenWs& w, float* z, float* eps_out, int B, int T, int nsteps,
                        const float* host_coef, const float* noise, int use_philox, uint64_t seed,
                        const unsigned long long* seed_ptr, uint64_t chain0, cudaStream_t s) {
  const DenTcPack* t = d->tc[precision];
  // prepared launches: tensor maps are encoded once per call, the epilogue scalars of the last layer change per step
  TcLaunch* L[DEN_LAYERS] = {nullptr};
  struct Guard { TcLaunch** l; ~Guard() { for (int i = 0; i < DEN_LAYERS; ++i) if (l[i]) tc_free(l[i]); } } guard{L};
  // destinations of leaky_relu(out_L): next layer's input slice, and the U-net skip slice of the matching out layer
  const int skip_to[3] = {6, 5, 4};
  for (int i = 0; i < DEN_LAYERS; ++i) {
    GemmPlan p{};
    p.A = w.A[i];
    p.B = B; p.Hm = 1; p.Wm = 1; p.Cs = d->din[i] + d->dout[i];
    p.ntaps = 1;
    p.taps[0] = Tap{0, 0, 0, 0};
    p.N = p.Np = 4 * d->dout[i];
    const int bn = tc_den_tile_width(B, p.Np), v = den_bn_index(bn);
    p.Wtc = t->Wq[v][i];
    p.ksplit = 1;
    p.epi.kind = i == DEN_LAYERS - 1 ? EPI_DEN_FINAL : EPI_DEN_LAYER;
    DenEpi& e = p.epi.den;
    e.din = d->din[i];
    e.bn = bn;
    e.bias4 = t->bias4[i];
    if (i < DEN_LAYERS - 1) {
      e.dst1 = w.A[i + 1]; e.ld1 = d->din[i + 1] + d->dout[i + 1]; e.off1 = 0;
      if (i < 3) {
        const int j = skip_to[i];
        e.dst2 = w.A[j]; e.ld2 = d->din[j] + d->dout[j]; e.off2 = d->dout[j - 1];
      }
    } else {
      e.z = z; e.eps_out = eps_out; e.nz = d->nz; e.residual = d->residual;
      e.use_philox = use_philox; e.seed = seed; e.seed_ptr = seed_ptr; e.chain0 = chain0;
    }
    DAMC_TRY(tc_prepare(p, precision, &L[i]));
  }
  DenPrepArgs pa{};
  pa.z = z; pa.Bp = d->Bp; pa.cx = w.cx; pa.B = B; pa.nz = d->nz; pa.csum = d->csum;
  pa.fp16 = precision == DAMC_PREC_FP16;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    pa.A[i] = w.A[i]; pa.ld[i] = d->din[i] + d->dout[i]; pa.din[i] = d->din[i]; pa.dout[i] = d->dout[i]; pa.coff[i] = d->coff[i];
  }
  const size_t prep_smem = sizeof(float) * ((size_t)d->nz * (d->nz / 2) + (size_t)EMB_PITCH * d->nz);
  const int nb_embed = ceil_div(B, EMB_CHAINS), nb_ctx = ceil_div(B, CTX_CHAINS);
  for (int st = 0; st < nsteps; ++st) {
    const int irev = eps_out ? 0 : T - 1 - st;
    pa.ctrow = w.ct + (size_t)irev * d->csum;
    {
      cudaLaunchConfig_t cfg{};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.gridDim = dim3(nb_embed + nb_ctx);
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = prep_smem;
      cfg.stream = s;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      DAMC_CUDA(cudaLaunchKernelEx(&cfg, den_prep_kernel, pa, nb_embed));
    }
    const float* cf = host_coef + 8 * (size_t)st;
    DenEpi& e = tc_plan(L[DEN_LAYERS - 1])->epi.den;
    e.c_pred = cf[0]; e.c_eps = cf[1]; e.c_zt = cf[2]; e.c_x = cf[3]; e.c_std = cf[4];
    e.last = cf[5] != 0.f;
    e.step = (unsigned long long)st;
    e.noise = noise ? noise + (size_t)st * B * d->nz : nullptr;
    for (int i = 0; i < DEN_LAYERS; ++i) {
      profile_mark(s, true);
      const int r = tc_launch(L[i], s);
      profile_mark(s, false);
      if (r != DAMC_OK) return r;
    }
  }
  return DAMC_OK;
}

int den_tc_run(const DenPack* d, int precision, const DenWs& w, float* z, float* eps_out, int B, int T, int nsteps,
               const float* host_coef, const float* noise, int use_philox, uint64_t seed, uint64_t chain0,
               cudaStream_t s) {
  DAMC_TRY(den_tc_ensure(d, precision, s));
  if (eps_out == nullptr && w.G && den_seq_supported(d))   // sampler loop: hoisted-context schedule (denoiser_seq.cu)
    return den_seq_run(d, precision, w, z, B, T, nsteps, host_coef, noise, use_philox, seed, chain0, s);
  if (den_cluster_supported(d, B))   // one launch for all T steps: 4-CTA clusters own 128 chains each (denoiser_cluster.cu)
    return den_cluster_run(d, precision, w, z, eps_out, B, T, nsteps, host_coef, noise, use_philox, seed, chain0, s);
  DenTcPack* t = d->tc[precision];
  // weight row orders for the tile widths this batch size uses (packed on first use, refreshed by refill())
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int v = den_bn_index(tc_den_tile_width(B, 4 * d->dout[i]));
    if (!t->live[v]) { t->live[v] = true; DAMC_TRY(den_tc_pack_variant(d, precision, v, s)); }
  }
  const size_t prep_smem = sizeof(float) * ((size_t)d->nz * (d->nz / 2) + (size_t)EMB_PITCH * d->nz);
  DAMC_CUDA(cudaFuncSetAttribute(den_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
  count_launch(2 + nsteps * (1 + DEN_LAYERS));

  // ---- CUDA graph replay of the launch-bound T-step loop (Philox noise, whole sampler) ------------------------------
  static const bool use_graph = []{ const char* e = getenv("DAMC_GRAPH"); return !(e && e[0] == '0'); }();
  const bool graphable = use_graph && !profiling() && noise == nullptr && eps_out == nullptr && nsteps > 1 && w.zbuf && w.seed_dev;
  if (!graphable) {
    DAMC_TRY(den_tc_issue(d, precision, w, z, eps_out, B, T, nsteps, host_coef, noise, use_philox, seed, nullptr, chain0, s));
    DAMC_CUDA(cudaGetLastError());
    return DAMC_OK;
  }
  unsigned long long h = 1469598103934665603ull;   // FNV-1a over the step coefficients (schedule, var_type, with_noise)
  for (size_t i = 0; i < 8 * (size_t)nsteps * sizeof(float); ++i) { h ^= reinterpret_cast<const unsigned char*>(host_coef)[i]; h *= 1099511628211ull; }
  const DenTcPack::GraphKey key = {B, T, use_philox, w.base, (unsigned long long)chain0, h};
  const bool same = t->gkey.B == key.B && t->gkey.T == key.T && t->gkey.use_philox == key.use_philox &&
                    t->gkey.ws_base == key.ws_base && t->gkey.chain0 == key.chain0 && t->gkey.coef_hash == key.coef_hash;
  if (!same) {   // new configuration: run it directly once; capture if it comes back
    if (t->gexec) { cudaGraphExecDestroy(t->gexec); t->gexec = nullptr; }
    t->gkey = key;
    t->gkey_seen = 1;
    DAMC_TRY(den_tc_issue(d, precision, w, z, eps_out, B, T, nsteps, host_coef, noise, use_philox, seed, nullptr, chain0, s));
    DAMC_CUDA(cudaGetLastError());
    return DAMC_OK;
  }
  ++t->gkey_seen;
  if (!t->gexec) {
    cudaGraph_t graph = nullptr;
    if (!t->cap_stream) DAMC_CUDA(cudaStreamCreateWithFlags(&t->cap_stream, cudaStreamNonBlocking));
    cudaStream_t cs = t->cap_stream;   // nothing executes during capture; the instantiated graph is launched on s
    DAMC_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    const int r = den_tc_issue(d, precision, w, w.zbuf, nullptr, B, T, nsteps, host_coef, nullptr, use_philox, 0, w.seed_dev, chain0, cs);
    const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (r != DAMC_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess || !graph) DAMC_FAIL(DAMC_ERR_CUDA, "denoiser: stream capture failed: %s", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&t->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { t->gexec = nullptr; DAMC_FAIL(DAMC_ERR_CUDA, "denoiser: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
  }
  const unsigned long long seed_host = seed;
  DAMC_CUDA(cudaMemcpyAsync(w.seed_dev, &seed_host, sizeof(seed_host), cudaMemcpyHostToDevice, s));
  DAMC_CUDA(cudaMemcpyAsync(w.zbuf, z, sizeof(float) * (size_t)B * d->nz, cudaMemcpyDeviceToDevice, s));
  DAMC_CUDA(cudaGraphLaunch(t->gexec, s));
  DAMC_CUDA(cudaMemcpyAsync(z, w.zbuf, sizeof(float) * (size_t)B * d->nz, cudaMemcpyDeviceToDevice, s));
  return DAMC_OK;
}

}  // namespace damc
