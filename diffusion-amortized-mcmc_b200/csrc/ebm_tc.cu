// ebm_tc.cu -- the posterior sampler's per-step tail on the tensor cores (16-bit generator modes): EBM forward, analytic backward
// with respect to z, generator-gradient partial sums, prior term, noise and the Langevin update, one launch, one CTA per 128 chains.
//
// Replaces, per Langevin step, netE(z) + autograd.grad through netE + the update of sample_langevin_post_z_with_prior
// (reference workspace/src/MCMC.py:57-64; netE = Linear(nz,ndf) LReLU Linear(ndf,ndf) LReLU Linear(ndf,1), diffusion_net.py:212-223).
// The CUDA-core form (ebm_step_kernel, ebm_langevin.cu) streams the fp32 MLP from L2 once per 4-16 chains: 263 us at 16 384
// SVHN chains (5 % of a step), 40 us at 128 CIFAR-10 chains (6 %).  Here the four mat-mat products of a 128-chain tile
//     h1 = z W1^T            a1 = lrelu(h1 + b1)            (m1 = h1 + b1 > 0, kept as bits in the row's thread)
//     h2 = a1 W2^T           d2 = (h2 + b2 > 0 ? 1 : slope) * w3
//     t  = d2 W2             d1 = (m1 ? 1 : slope) * t
//     gE = d1 W1
// run as tcgen05 GEMMs (M = 128 chains, N = 256 | nz, K = nz | 256, fp32 accumulators in TMEM): the activations never leave
// the SM (each epilogue writes the next GEMM's K-major SWIZZLE_128B operand tile into shared memory), the 16-bit weights
// (hidden width zero-padded to 256) stream through a TMA ring from L2.  z, the gradient sum and the update stay fp32; the
// Philox draw is the one of the CUDA-core kernel (same bits).  Used in the 16-bit modes when no trace is requested; the fp32 /
// tf32 modes, traces and targets without an EBM keep ebm_step_kernel.
// Operand type: fp16 in BOTH 16-bit generator modes.  dE/dz is discontinuous in the pre-activations (a LeakyReLU sign flip of one
// of the 2 ndf hidden units changes it by a few per cent); with bf16 operands (2^-9) about one unit per chain and evaluation
// sits inside the rounding band, with fp16 (2^-12; every EBM quantity is far inside the fp16 range) about 0.15 -- measured:
// median per-chain error of dE/dz 5e-4, against ~1e-5 for the fp32 CUDA-core kernel and ~1e-2 with bf16 operands.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <atomic>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "tc_ptx.cuh"

namespace damc {

constexpr int ET_THREADS = 320;          // warp 0: weight TMA, warp 1: MMA issuer + TMEM, warps 2-9: workers (two per TMEM lane quarter)
constexpr int ET_H = 256;                // hidden width, zero-padded
constexpr int ET_TILE = 128 * 128;       // one K-major k-block of a 128-row operand tile (16 KB)
constexpr int ET_WSTAGE = ET_H * 128;    // one weight k-block: 256 rows x 128 B
constexpr int ET_STAGES = 2;

struct EbmTcPack {
  void* slab = nullptr;
  void *W1p = nullptr, *W2p = nullptr, *W2Tp = nullptr, *W1Tp = nullptr;   // [256][nzp], [256][256], [256][256], [nzp][256]
  float* vec = nullptr;                                                    // b1, b2, w3 zero-padded to 256 each (fp32)
  CUtensorMap tm[4];
  int nzp = 0;
};

template <typename T>
__global__ void pack_ebm_tc_kernel(const float* __restrict__ W1, const float* __restrict__ W2, const float* __restrict__ b1,
                                   const float* __restrict__ b2, const float* __restrict__ w3, int nz, int ndf, int nzp,
                                   T* __restrict__ W1p, T* __restrict__ W2p, T* __restrict__ W2Tp, T* __restrict__ W1Tp,
                                   float* __restrict__ vec, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ET_H * nzp) {   // W1p[j][k] and its transpose W1Tp[k][j]
    const int j = i / nzp, k = i - j * nzp;
    const float v = (j < ndf && k < nz) ? W1[(size_t)j * nz + k] : 0.f;
    W1p[i] = T(v);
    W1Tp[(size_t)k * ET_H + j] = T(v);
  }
  if (i < ET_H * ET_H) {   // W2p[j2][j1] and W2Tp[j1][j2]
    const int j2 = i / ET_H, j1 = i - j2 * ET_H;
    const float v = (j2 < ndf && j1 < ndf) ? W2[(size_t)j2 * ndf + j1] : 0.f;
    W2p[i] = T(v);
    W2Tp[(size_t)j1 * ET_H + j2] = T(v);
  }
  if (i < ET_H) {
    vec[i] = i < ndf ? b1[i] : 0.f;
    vec[ET_H + i] = i < ndf ? b2[i] : 0.f;
    vec[2 * ET_H + i] = i < ndf ? w3[i] : 0.f;
  }
}

// sum of the generator-gradient split-K partials, ascending split order (the order ebm_step_kernel uses), in place into split 0.
// One 128-chain CTA would otherwise pull all S x 128 rows through one SM (2 MB at 128 CIFAR-10 chains, S = 32).
__global__ void __launch_bounds__(256) dz_reduce_kernel(float* __restrict__ part, int S, size_t n4, size_t stride4) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  float4* p = reinterpret_cast<float4*>(part);
  float4 acc = p[i];
  for (int s0 = 1; s0 < S; s0 += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = s0 + u < S ? p[(size_t)(s0 + u) * stride4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  p[i] = acc;
}

struct EbmTcArgs {
  float* z;
  int B, nz, nzp;
  float step, slope;
  int with_noise;
  const float* noise;
  unsigned long long seed, chain0, step_index;
  const unsigned long long* seed_ptr;
  const float* gpart;
  int nsplit, gstride;
  float gpart_scale;
  const float* vec;   // b1 | b2 | w3, 256 floats each
  int op_fp16;
  uint32_t idesc_h, idesc_z;
  int K;                      // Langevin steps in this launch (posterior tail: 1; prior sampler: all K, weights re-streamed per step)
  long long noise_stride;     // elements between the injected-noise slabs of consecutive steps
};

__global__ void __launch_bounds__(ET_THREADS, 1)
ebm_tc_step_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                   const __grid_constant__ CUtensorMap tmW2T, const __grid_constant__ CUtensorMap tmW1T,
                   const __grid_constant__ EbmTcArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bufB = base, bufA = base + 4u * ET_TILE, ring = base + 8u * ET_TILE;   // 64 KB | 64 KB | 2 x 32 KB
  const uint32_t off_vec = 8u * ET_TILE + ET_STAGES * ET_WSTAGE;
  float* const svec = reinterpret_cast<float*>(gen_base + off_vec);
  const uint32_t bars = base + off_vec + 3u * ET_H * 4u;
  auto bar_wfull = [&](int s) { return bars + 8u * s; };
  auto bar_wempty = [&](int s) { return bars + 8u * (ET_STAGES + s); };
  const uint32_t bar_acc = bars + 8u * (2 * ET_STAGES), bar_ready = bars + 8u * (2 * ET_STAGES + 1);
  const uint32_t tmem_slot = bars + 8u * (2 * ET_STAGES + 2);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + off_vec + 3u * ET_H * 4u + 8u * (2 * ET_STAGES + 2));

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); prefetch_tmap(&tmW2T); prefetch_tmap(&tmW1T);
    for (int s = 0; s < ET_STAGES; ++s) { mbar_init(bar_wfull(s), 1); mbar_init(bar_wempty(s), 1); }
    mbar_init(bar_acc, 1);
    mbar_init(bar_ready, 8);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  for (int i = threadIdx.x; i < 3 * ET_H; i += ET_THREADS) svec[i] = __ldg(a.vec + i);   // packed by an earlier, completed launch
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");   // z and the generator's gradient partials come from the previous kernels
  const uint32_t tmem_base = *tmem_slot_ptr;
  const int kz = a.nzp >> 6;              // k-blocks of the first GEMM
  const int nkb[4] = {kz, 4, 4, 4};

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const CUtensorMap* tms[4] = {&tmW1, &tmW2, &tmW2T, &tmW1T};
      for (int it = 0; it < a.K; ++it)
      for (int g = 0; g < 4; ++g)
        for (int kb = 0; kb < nkb[g]; ++kb) {
          mbar_wait(bar_wempty(stage), phase ^ 1u);
          mbar_expect_tx(bar_wfull(stage), (uint32_t)(g == 3 ? a.nzp : ET_H) * 128u);
          tma_load_2d(ring + (uint32_t)stage * ET_WSTAGE, tms[g], bar_wfull(stage), kb * 64, 0);
          if (++stage == ET_STAGES) { stage = 0; phase ^= 1u; }
        }
    }
  } else if (warp == 1) {
    {   // the whole warp runs the issue loop (converged); one elected lane issues (umma_elect, tc_ptx.cuh)
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < a.K; ++it)
      for (int g = 0; g < 4; ++g) {
        mbar_wait(bar_ready, (uint32_t)g & 1u);   // the operand tile of GEMM g is in shared memory (4 phases per step: same parities)
        tc_fence_after();
        const uint32_t abuf = (g & 1) ? bufA : bufB;
        for (int kb = 0; kb < nkb[g]; ++kb) {
          mbar_wait(bar_wfull(stage), phase);
          tc_fence_after();
          const uint64_t adesc = make_sdesc(abuf + (uint32_t)kb * ET_TILE), bdesc = make_sdesc(ring + (uint32_t)stage * ET_WSTAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_elect<false, 1>(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), g == 3 ? a.idesc_z : a.idesc_h,
                                 (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_elect<1>(bar_wempty(stage));
          if (++stage == ET_STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit_elect<1>(bar_acc);
      }
    }
  } else {
    // ===================== 8 worker warps: two per TMEM lane quarter, each thread owns half of a chain row's columns ==========
    const int ew = warp - 2, q = warp & 3, half = ew >> 2, r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool fp16 = a.op_fp16 != 0;
    auto ready = [&]() {   // this warp's part of the next operand tile is written: hand it to the tensor core
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_ready);
    };
    // 8 consecutive columns [col, col + 8) of row r -> one 16-byte chunk of the K-major SWIZZLE_128B tile
    auto put8 = [&](uint32_t buf, int col, const float (&v)[8]) {
      const uint32_t dst = buf + (uint32_t)(col >> 6) * ET_TILE + (uint32_t)r * 128u + (uint32_t)((((col & 63) >> 3) ^ (r & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack2(fp16, v[0], v[1])), "r"(pack2(fp16, v[2], v[3])),
                   "r"(pack2(fp16, v[4], v[5])), "r"(pack2(fp16, v[6], v[7])) : "memory");
    };
    // ---- operand of GEMM 0: the z tile.  A warp takes 16 rows; a row is one coalesced 512-byte read (lane = 4 columns); four
    // rows' loads are in flight before the first is converted ----
    const int c4 = lane * 4;
    for (int it = 0; it < a.K; ++it) {
#pragma unroll 1
    for (int r0 = 0; r0 < 16; r0 += 4) {
      float4 zq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long bb = (long long)blockIdx.x * 128 + ew * 16 + r0 + u;
        zq[u] = (bb < a.B && c4 < a.nz) ? *reinterpret_cast<const float4*>(a.z + bb * a.nz + c4) : make_float4(0.f, 0.f, 0.f, 0.f);   // nz % 4 == 0
      }
      if (c4 < a.nzp) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int R = ew * 16 + r0 + u;
          const uint32_t dst = bufB + (uint32_t)(c4 >> 6) * ET_TILE + (uint32_t)R * 128u + (uint32_t)((((c4 & 63) >> 3) ^ (R & 7)) << 4) +
                               (uint32_t)((c4 & 4) << 1);
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(pack2(fp16, zq[u].x, zq[u].y)), "r"(pack2(fp16, zq[u].z, zq[u].w)) : "memory");
        }
      }
    }
    ready();
    uint32_t m1[4];
    // ---- epilogue 1: a1 = lrelu(h1 + b1) -> bufA ; epilogue 2: d2 = (h2 + b2 > 0 ? 1 : slope) w3 -> bufB ; epilogue 3: d1 -> bufA ----
#pragma unroll 1
    for (int g = 0; g < 3; ++g) {
      mbar_wait(bar_acc, (uint32_t)g & 1u);
      tc_fence_after();
      const uint32_t obuf = (g & 1) ? bufB : bufA;
#pragma unroll
      for (int i32 = 0; i32 < 4; ++i32) {
        const int c32 = half * 4 + i32;
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)(c32 * 32), v);
        tmem_ld_wait();
        uint32_t bits = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = c32 * 32 + j * 8 + e;
            const float acc = __uint_as_float(v[j * 8 + e]);
            if (g == 0) {
              const float h = acc + svec[c];
              const bool p = h > 0.f;
              bits |= (p ? 1u : 0u) << (j * 8 + e);
              o[e] = p ? h : a.slope * h;
            } else if (g == 1) {
              const float h = acc + svec[ET_H + c];
              o[e] = (h > 0.f ? 1.f : a.slope) * svec[2 * ET_H + c];
            } else {
              o[e] = ((m1[i32] >> (j * 8 + e)) & 1u) ? acc : a.slope * acc;
            }
          }
          put8(obuf, c32 * 32 + j * 8, o);
        }
        if (g == 0) m1[i32] = bits;
      }
      ready();
    }
    // ---- epilogue 4: gE (TMEM, one row per thread) -> fp32 tile in bufB (free since GEMM 2 finished), 16-byte chunks XOR-swizzled
    // by the row so that both the row-wise writes here and the warp-coalesced reads below are conflict-free ----
    mbar_wait(bar_acc, 1u);
    tc_fence_after();
    const int qpr = a.nzp >> 2;                      // float4 chunks per row (16 | 32)
    float4* const gbuf = reinterpret_cast<float4*>(gen_base);   // = bufB
    {
      const int ncol = a.nzp >> 1;                   // this thread's half of the row
#pragma unroll 1
      for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(t_lane + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int ch = (c0 >> 2) + j;
          gbuf[r * qpr + (ch ^ (r & (qpr - 1)))] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
        }
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 worker warps
    // ---- update: a warp takes 16 rows; lane = one float4 of the row: coalesced z / gradient / noise reads, one Philox quad per
    // lane, coalesced z writes ----
    const float half_s2 = 0.5f * a.step * a.step;
    const unsigned long long seed = a.seed_ptr ? a.seed_ptr[0] : a.seed;
    const unsigned long long chain_base = (a.seed_ptr ? a.seed_ptr[1] : a.chain0) + (unsigned long long)blockIdx.x * 128ull;
    const unsigned long long stp = (a.seed_ptr ? a.seed_ptr[2] : 0ull) + a.step_index + (unsigned long long)it;
    const float* noise_it = a.noise ? a.noise + (long long)it * a.noise_stride : nullptr;
    const bool live = c4 < a.nz;                     // nz % 4 == 0: a quad is either live or padding
#pragma unroll 1
    for (int r0 = 0; r0 < 16; r0 += 4) {
      float4 zq[4], gq[4], nq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int R = ew * 16 + r0 + u;
        const long long bb = (long long)blockIdx.x * 128 + R;
        const bool okr = live && bb < a.B;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        zq[u] = okr ? *reinterpret_cast<const float4*>(a.z + bb * a.nz + c4) : zero;
        gq[u] = (okr && a.gpart != nullptr) ? *reinterpret_cast<const float4*>(a.gpart + (size_t)bb * a.gstride + c4) : zero;   // split 0 = the reduced sum
        nq[u] = (okr && a.with_noise && noise_it) ? *reinterpret_cast<const float4*>(noise_it + bb * a.nz + c4) : zero;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int R = ew * 16 + r0 + u;
        const long long bb = (long long)blockIdx.x * 128 + R;
        if (!(live && bb < a.B)) continue;
        const float4 ge = gbuf[R * qpr + (lane ^ (R & (qpr - 1)))];
        float nrm[4] = {nq[u].x, nq[u].y, nq[u].z, nq[u].w};
        if (a.with_noise && !noise_it) philox_normal4(seed, chain_base + (unsigned long long)R, stp, (uint32_t)lane, nrm);
        const float zv[4] = {zq[u].x, zq[u].y, zq[u].z, zq[u].w}, gG[4] = {gq[u].x, gq[u].y, gq[u].z, gq[u].w};
        const float gE[4] = {ge.x, ge.y, ge.z, ge.w};
        float zn[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float g = fmaf(gG[e], a.gpart_scale, gE[e]);
          zn[e] = zv[e] - half_s2 * (g + zv[e]) + a.step * nrm[e];
        }
        *reinterpret_cast<float4*>(a.z + bb * a.nz + c4) = make_float4(zn[0], zn[1], zn[2], zn[3]);
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // every warp is done with the fp32 tile in bufB before the next step's z tile lands there
    }   // step loop

  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
static int tc_slot(int precision) { return precision == DAMC_PREC_FP16 ? 1 : 0; }
constexpr int ET_PREC = DAMC_PREC_FP16;   // operand type of the EBM GEMMs in every 16-bit generator mode (see the header)

void ebm_tc_free(EbmTcPack* t) {
  if (!t) return;
  if (t->slab) cudaFree(t->slab);
  delete t;
}

int ebm_tc_refill(const MlpPack* m, int precision, cudaStream_t s, const int* dirty) {
  EbmTcPack* t = m->tcp[tc_slot(precision)];
  if (!t) return DAMC_OK;
  const int n = ET_H * ET_H, blocks = ceil_div(n, 256);
  if (precision == DAMC_PREC_FP16)
    pack_ebm_tc_kernel<__half><<<blocks, 256, 0, s>>>(m->src[0], m->src[2], m->src[1], m->src[3], m->src[4], m->nz, m->ndf, t->nzp,
                                                      (__half*)t->W1p, (__half*)t->W2p, (__half*)t->W2Tp, (__half*)t->W1Tp, t->vec, dirty);
  else
    pack_ebm_tc_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(m->src[0], m->src[2], m->src[1], m->src[3], m->src[4], m->nz, m->ndf, t->nzp,
                                                             (__nv_bfloat16*)t->W1p, (__nv_bfloat16*)t->W2p, (__nv_bfloat16*)t->W2Tp,
                                                             (__nv_bfloat16*)t->W1Tp, t->vec, dirty);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

bool ebm_tc_usable(const MlpPack* m, int precision, const float* trace) {
  const char* e = getenv("DAMC_EBM_TC");   // read per call: the tests compare both forms in one process
  return !(e && e[0] == '0') && m != nullptr && is_tc_precision(precision) && trace == nullptr && m->nz >= 4 && m->nz <= 128 && (m->nz % 4) == 0 &&
         m->ndf <= ET_H && tc_available();
}

static int ebm_tc_ensure(const MlpPack* m, int precision, cudaStream_t s) {
  EbmTcPack*& t = m->tcp[tc_slot(precision)];
  if (t) return DAMC_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone)
    DAMC_FAIL(DAMC_ERR_INVALID, "EBM tensor-core pack must be built before a stream capture");
  EbmTcPack* n = new EbmTcPack();
  n->nzp = (int)align_up(m->nz, 64);
  const size_t e1 = (size_t)ET_H * n->nzp * 2, e2 = (size_t)ET_H * ET_H * 2;
  const size_t bytes = 2 * e1 + 2 * e2 + 3 * ET_H * sizeof(float);
  if (cudaMalloc(&n->slab, bytes) != cudaSuccess) { delete n; DAMC_FAIL(DAMC_ERR_CUDA, "EBM tensor-core pack: cudaMalloc failed"); }
  char* p = (char*)n->slab;
  n->W1p = p; p += e1;
  n->W1Tp = p; p += e1;
  n->W2p = p; p += e2;
  n->W2Tp = p; p += e2;
  n->vec = (float*)p;
  const int fp16 = precision == DAMC_PREC_FP16 ? 1 : 0;
  int r = tc_encode_2d(&n->tm[0], fp16, n->W1p, n->nzp, ET_H, ET_H);
  if (r == DAMC_OK) r = tc_encode_2d(&n->tm[1], fp16, n->W2p, ET_H, ET_H, ET_H);
  if (r == DAMC_OK) r = tc_encode_2d(&n->tm[2], fp16, n->W2Tp, ET_H, ET_H, ET_H);
  if (r == DAMC_OK) r = tc_encode_2d(&n->tm[3], fp16, n->W1Tp, ET_H, n->nzp, n->nzp);
  if (r != DAMC_OK) { ebm_tc_free(n); return r; }
  t = n;
  return ebm_tc_refill(m, precision, s, nullptr);
}

int launch_ebm_step_tc(const MlpPack* m, int precision, float* z, int B, float step, int with_noise, const float* noise,
                       uint64_t seed, uint64_t chain0, uint64_t step_index, const float* gpart, int nsplit, int gstride,
                       float gpart_scale, cudaStream_t stream, const unsigned long long* seed_ptr, int K) {
  (void)precision;
  if (K < 1) return DAMC_OK;
  DAMC_TRY(ebm_tc_ensure(m, ET_PREC, stream));
  const EbmTcPack* t = m->tcp[tc_slot(ET_PREC)];
  EbmTcArgs a{};
  a.z = z; a.B = B; a.nz = m->nz; a.nzp = t->nzp; a.step = step; a.slope = m->slope; a.with_noise = with_noise; a.noise = noise;
  a.seed = seed; a.chain0 = chain0; a.step_index = step_index; a.seed_ptr = seed_ptr;
  a.gpart = gpart; a.nsplit = nsplit; a.gstride = gstride; a.gpart_scale = gpart_scale;
  a.vec = t->vec;
  a.K = K;
  a.noise_stride = (long long)B * m->nz;
  a.op_fp16 = ET_PREC == DAMC_PREC_FP16 ? 1 : 0;
  const uint32_t opfmt = a.op_fp16 ? 0u : 1u;
  a.idesc_h = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(ET_H >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_z = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(t->nzp >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const size_t smem = 8 * ET_TILE + ET_STAGES * ET_WSTAGE + 3 * ET_H * 4 + 8 * (2 * ET_STAGES + 2) + 16 + 1024;
  static std::atomic<bool> attr_done[64];   // the dynamic-smem opt-in is a per-device function attribute (idempotent if two threads race)
  int dev = 0;
  DAMC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM tensor-core step: device ordinal %d out of range", dev);
  if (!attr_done[dev].load(std::memory_order_acquire)) {
    DAMC_CUDA(cudaFuncSetAttribute(ebm_tc_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done[dev].store(true, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (gpart != nullptr && nsplit > 1) {   // split-K partials of the first layer's dgrad -> one sum per chain, with the whole GPU
    const size_t n4 = (size_t)B * gstride / 4;
    cfg.gridDim = dim3((unsigned)((n4 + 255) / 256));
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    DAMC_CUDA(cudaLaunchKernelEx(&cfg, dz_reduce_kernel, const_cast<float*>(gpart), nsplit, n4, n4));
    count_launch();
  }
  cfg.gridDim = dim3(ceil_div(B, 128));
  cfg.blockDim = dim3(ET_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DAMC_CUDA(cudaLaunchKernelEx(&cfg, ebm_tc_step_kernel, t->tm[0], t->tm[1], t->tm[2], t->tm[3], a));
  count_launch();
  return DAMC_OK;
}

}  // namespace damc
