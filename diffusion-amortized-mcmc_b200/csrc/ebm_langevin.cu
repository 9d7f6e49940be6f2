// ebm_langevin.cu -- persistent on-device Langevin chain loop with the EBM-prior MLP resident in shared memory.
//
// Replaces sample_langevin_prior_z (reference workspace/src/MCMC.py:27-46) and the per-step "E + prior + update"
// tail of sample_langevin_post_z_with_prior (:57-64).  One launch runs all K steps:
//     h1 = W1 z + b1 ; h2 = W2 lrelu(h1) + b2 ; E = w3.lrelu(h2) + b3                  (diffusion_net.py:212-223)
//     gE = W1^T( n1 * W2^T( n2 * w3 ) ),  n = 1 | slope                                  (analytic backward)
//     z <- z - s^2/2 (gE [+ gG] + z) + s * eps                                           (MCMC.py:36-38 / :62-64)
//
// B200 layout: fp32 weights are 4*(nz*ndf + ndf*ndf) = 262 KB at nz=128, ndf=200 -- more than one SM's 227 KB -- so a
// 2-CTA thread-block cluster owns a tile of CH chains and splits the HIDDEN units: CTA r keeps rows J_r of W1 and
// columns J_r of W2 (131 KB) for the whole kernel.  Two DSMEM exchanges per step (layer-2 partial sums, dz partial
// sums); both CTAs then hold identical z (fp add is commutative, so the two sums are bit-identical).  No HBM traffic
// inside the loop except optional injected noise; z is read once and written once.
#include <cooperative_groups.h>

#include <stdlib.h>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace cg = cooperative_groups;

namespace damc {

static __host__ __device__ inline int pad_stride(int n) {  // >= n, == 4 (mod 32): conflict-free LDS.128 across rows
  int s = (n + 3) & ~3;
  while ((s & 31) != 4) s += 4;
  return s;
}

struct EbmSmemPlan {
  int HJ, HJP, NZP, S1, S2;
  size_t oW1, oW2, ob1, ob2, ow3, oz, oa1, op, opp, od2, ot, od1, og, ogp, ored, total;
};

template <int CH>
static __host__ __device__ inline EbmSmemPlan ebm_plan(int nz, int ndf) {
  EbmSmemPlan p;
  p.HJ = (ndf + 1) / 2;
  p.HJP = (p.HJ + 3) & ~3;
  p.NZP = (nz + 3) & ~3;
  p.S1 = pad_stride(p.NZP);
  p.S2 = pad_stride(p.HJP);
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 3) & ~(size_t)3; return r; };
  p.oW1 = take((size_t)p.HJ * p.S1);
  p.oW2 = take((size_t)ndf * p.S2);
  p.ob1 = take(p.HJ);
  p.ob2 = take(ndf);
  p.ow3 = take(ndf);
  p.oz = take((size_t)CH * p.NZP);
  p.oa1 = take((size_t)CH * p.HJP);
  p.op = take((size_t)CH * ndf);
  p.opp = take((size_t)CH * ndf);
  p.od2 = take((size_t)ndf * CH);
  p.ot = take((size_t)2 * p.HJ * CH);
  p.od1 = take((size_t)p.HJ * CH);
  p.og = take((size_t)2 * CH * p.NZP);
  p.ogp = take((size_t)CH * p.NZP);
  p.ored = take(4 * CH);
  p.total = o * sizeof(float);
  return p;
}

struct EbmArgs {
  const float *W1, *b1, *W2, *b2, *w3, *b3;  // PyTorch Linear layouts, fp32, device
  int nz, ndf;
  float slope;
  float* z;  // [B,nz] in/out
  int B, K;
  float step;
  int with_noise;
  const float* noise;  // [K,B,nz] or null
  uint64_t seed, chain0, step0;
  float* trace;      // null or [K, trace_stride]
  int trace_stride;  // 2 (prior: en, z_norm) or 4 (posterior: en, llhd(other kernel), z_n, mean grad)
  const float* gpart;  // null or [nsplit][B][gstride] generator dz partial sums
  int nsplit, gstride;
  float inv_count;  // 1/(B*nz) for mean(grad)
  int use_ebm;      // 0: skip the MLP (toy-style target), update only
};

template <int CH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) ebm_langevin_kernel(const EbmArgs a) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned rank = cluster.block_rank();
  const int tid = threadIdx.x;
  const int nz = a.nz, ndf = a.ndf;
  const EbmSmemPlan P = ebm_plan<CH>(nz, ndf);
  const int HJ = P.HJ, HJP = P.HJP, NZP = P.NZP, S1 = P.S1, S2 = P.S2;
  const int j0 = rank * HJ;                      // first hidden unit owned by this CTA
  const int hj = max(0, min(HJ, ndf - j0));      // number owned (rank 1 may own fewer when ndf is odd)
  float* sW1 = smem + P.oW1;  // [HJ][S1]   rows j0.. of W1
  float* sW2 = smem + P.oW2;  // [ndf][S2]  columns j0.. of W2
  float* sb1 = smem + P.ob1;
  float* sb2 = smem + P.ob2;
  float* sw3 = smem + P.ow3;
  float* sz = smem + P.oz;    // [CH][NZP]
  float* sa1 = smem + P.oa1;  // [CH][HJP]
  float* sp = smem + P.op;    // [CH][ndf]  own layer-2 partial
  float* spp = smem + P.opp;  // [CH][ndf]  peer's partial (written by the peer through DSMEM)
  float* sd2 = smem + P.od2;  // [ndf][CH]
  float* st = smem + P.ot;    // [2][HJ][CH]
  float* sd1 = smem + P.od1;  // [HJ][CH]
  float* sg = smem + P.og;    // [2][CH][NZP]
  float* sgp = smem + P.ogp;  // [CH][NZP]  peer's dz partial
  float* sred = smem + P.ored;
  float* peer_spp = cluster.map_shared_rank(spp, rank ^ 1);
  float* peer_sgp = cluster.map_shared_rank(sgp, rank ^ 1);

  const int cl = blockIdx.x >> 1;
  const int c0 = cl * CH;
  const int nvalid = min(CH, a.B - c0);

  // ---- one-time: weights -> smem (zero padded), z tile -> smem ---------------------------------------------------
  if (a.use_ebm) {
    {
      const int wid = tid >> 5, ln = tid & 31, nw = blockDim.x >> 5;
      for (int r = wid; r < HJ; r += nw) {  // one warp per row; the column loop unrolls into independent loads
        const float* src = a.W1 + (size_t)(j0 + r) * nz;
#pragma unroll 4
        for (int c = ln; c < S1; c += 32) sW1[r * S1 + c] = (r < hj && c < nz) ? __ldg(src + c) : 0.f;
      }
      for (int r = wid; r < ndf; r += nw) {
        const float* src = a.W2 + (size_t)r * ndf + j0;
#pragma unroll 4
        for (int c = ln; c < S2; c += 32) sW2[r * S2 + c] = (c < hj) ? __ldg(src + c) : 0.f;
      }
    }
    for (int i = tid; i < HJ; i += blockDim.x) sb1[i] = i < hj ? a.b1[j0 + i] : 0.f;
    for (int i = tid; i < ndf; i += blockDim.x) { sb2[i] = a.b2[i]; sw3[i] = a.w3[i]; }
    for (int i = tid; i < CH * HJP; i += blockDim.x) sa1[i] = 0.f;
  }
  for (int i = tid; i < CH * NZP; i += blockDim.x) {
    const int c = i / NZP, k = i - c * NZP;
    sz[i] = (c < nvalid && k < nz) ? a.z[(size_t)(c0 + c) * nz + k] : 0.f;
  }
  const float b3 = a.use_ebm ? a.b3[0] : 0.f;
  const float slope = a.slope;
  const float half_s2 = 0.5f * a.step * a.step;
  cluster.sync();

  for (int it = 0; it < a.K; ++it) {
    if (a.trace != nullptr) {
      if (tid < 4 * CH) sred[tid] = 0.f;
      __syncthreads();
    }
    if (a.use_ebm) {
      // ---- phase 1: h1 / a1 for owned hidden units (thread = unit x half of the chains) --------------------------
      if (tid < 2 * hj) {
        const int jl = tid % hj, cg0 = (tid / hj) * (CH / 2);
        float acc[CH / 2];
#pragma unroll
        for (int c = 0; c < CH / 2; ++c) acc[c] = 0.f;
        const float4* w = reinterpret_cast<const float4*>(sW1 + (size_t)jl * S1);
        for (int i = 0; i < NZP / 4; ++i) {
          const float4 wv = w[i];
#pragma unroll
          for (int c = 0; c < CH / 2; ++c) {
            const float4 zv = reinterpret_cast<const float4*>(sz + (size_t)(cg0 + c) * NZP)[i];
            acc[c] = fmaf(wv.x, zv.x, acc[c]);
            acc[c] = fmaf(wv.y, zv.y, acc[c]);
            acc[c] = fmaf(wv.z, zv.z, acc[c]);
            acc[c] = fmaf(wv.w, zv.w, acc[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < CH / 2; ++c) {
          const float h = acc[c] + sb1[jl];
          sa1[(size_t)(cg0 + c) * HJP + jl] = h > 0.f ? h : slope * h;
        }
      }
      __syncthreads();
      // ---- phase 2: layer-2 partial sums over owned units, for every output unit (thread = output unit) ---------
      if (tid < ndf) {
        float acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = 0.f;
        const float4* w = reinterpret_cast<const float4*>(sW2 + (size_t)tid * S2);
        for (int i = 0; i < HJP / 4; ++i) {
          const float4 wv = w[i];
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            const float4 av = reinterpret_cast<const float4*>(sa1 + (size_t)c * HJP)[i];
            acc[c] = fmaf(wv.x, av.x, acc[c]);
            acc[c] = fmaf(wv.y, av.y, acc[c]);
            acc[c] = fmaf(wv.z, av.z, acc[c]);
            acc[c] = fmaf(wv.w, av.w, acc[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          sp[c * ndf + tid] = acc[c];
          peer_spp[c * ndf + tid] = acc[c];  // DSMEM store
        }
      }
      cluster.sync();
      // ---- h2, E, d2 = n2 * w3 (thread = output unit; both CTAs compute the same values) --------------------------
      if (tid < ndf) {
        const float w3 = sw3[tid], bb = sb2[tid];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float h = (sp[c * ndf + tid] + spp[c * ndf + tid]) + bb;
          sd2[tid * CH + c] = (h > 0.f ? 1.f : slope) * w3;
          if (a.trace != nullptr && rank == 0 && c < nvalid) atomicAdd(&sred[c], w3 * (h > 0.f ? h : slope * h));
        }
      }
      __syncthreads();
      // ---- phase 3: t = W2[:,J]^T d2 over two halves of the output units (thread = owned unit x half) ------------
      if (tid < 2 * hj) {
        const int il = tid % hj, hf = tid / hj;
        const int jb = hf * ((ndf + 1) / 2), je = min(ndf, jb + (ndf + 1) / 2);
        float acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = 0.f;
        for (int j = jb; j < je; ++j) {
          const float wv = sW2[(size_t)j * S2 + il];
#pragma unroll
          for (int c4 = 0; c4 < CH / 4; ++c4) {
            const float4 dv = reinterpret_cast<const float4*>(sd2 + j * CH)[c4];
            acc[c4 * 4 + 0] = fmaf(wv, dv.x, acc[c4 * 4 + 0]);
            acc[c4 * 4 + 1] = fmaf(wv, dv.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(wv, dv.z, acc[c4 * 4 + 2]);
            acc[c4 * 4 + 3] = fmaf(wv, dv.w, acc[c4 * 4 + 3]);
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) st[((size_t)hf * HJ + il) * CH + c] = acc[c];
      }
      __syncthreads();
      for (int i = tid; i < hj * CH; i += blockDim.x) {
        const int il = i / CH, c = i - il * CH;
        const float tv = st[i] + st[(size_t)HJ * CH + i];
        sd1[i] = (sa1[(size_t)c * HJP + il] > 0.f ? 1.f : slope) * tv;
      }
      __syncthreads();
      // ---- phase 4: dz partial = W1[J,:]^T d1 (thread = latent dim x half of the owned units) ---------------------
      if (tid < 2 * NZP) {
        const int k = tid % NZP, hf = tid / NZP;
        const int ib = hf * ((hj + 1) / 2), ie = min(hj, ib + (hj + 1) / 2);
        float acc[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = 0.f;
        for (int i = ib; i < ie; ++i) {
          const float wv = sW1[(size_t)i * S1 + k];
#pragma unroll
          for (int c4 = 0; c4 < CH / 4; ++c4) {
            const float4 dv = reinterpret_cast<const float4*>(sd1 + i * CH)[c4];
            acc[c4 * 4 + 0] = fmaf(wv, dv.x, acc[c4 * 4 + 0]);
            acc[c4 * 4 + 1] = fmaf(wv, dv.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(wv, dv.z, acc[c4 * 4 + 2]);
            acc[c4 * 4 + 3] = fmaf(wv, dv.w, acc[c4 * 4 + 3]);
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) sg[((size_t)hf * CH + c) * NZP + k] = acc[c];
      }
      __syncthreads();
      for (int i = tid; i < CH * NZP; i += blockDim.x) {
        const float v = sg[i] + sg[(size_t)CH * NZP + i];
        sg[i] = v;
        peer_sgp[i] = v;  // DSMEM store
      }
      cluster.sync();
    }
    // ---- phase 5: fused update  z <- z - s^2/2 (gE + gG + z) + s eps   (thread = quad of latent dims) -------------
    for (int qb = 0; qb < CH * NZP / 4; qb += blockDim.x) {  // uniform trip count: warp_sum below needs all lanes
      const int q = min(qb + tid, CH * NZP / 4 - 1);
      const bool mine = qb + tid < CH * NZP / 4;
      const int c = q / (NZP / 4), k4 = (q - c * (NZP / 4)) * 4;
      const bool live = mine && c < nvalid;
      const uint64_t chain = a.chain0 + (uint64_t)(c0 + c);
      float4 zv = reinterpret_cast<float4*>(sz + (size_t)c * NZP)[k4 >> 2];
      float g[4] = {0.f, 0.f, 0.f, 0.f};
      if (a.use_ebm) {
        const float4 g0 = reinterpret_cast<const float4*>(sg + (size_t)c * NZP)[k4 >> 2];
        const float4 g1 = reinterpret_cast<const float4*>(sgp + (size_t)c * NZP)[k4 >> 2];
        g[0] = g0.x + g1.x; g[1] = g0.y + g1.y; g[2] = g0.z + g1.z; g[3] = g0.w + g1.w;
      }
      float nrm[4] = {0.f, 0.f, 0.f, 0.f};
      if (live) {
        if (a.gpart != nullptr) {
          for (int s = 0; s < a.nsplit; ++s) {
            const float* gp = a.gpart + ((size_t)s * a.B + (c0 + c)) * a.gstride + k4;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k4 + e < nz) g[e] += gp[e];
          }
        }
        if (a.with_noise) {
          if (a.noise != nullptr) {
            const float* np = a.noise + ((size_t)it * a.B + (c0 + c)) * nz + k4;
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (k4 + e < nz) nrm[e] = np[e];
          } else {
            philox_normal4(a.seed, chain, a.step0 + (uint64_t)it, (uint32_t)(k4 >> 2), nrm);
          }
        }
      }
      float zz[4] = {zv.x, zv.y, zv.z, zv.w};
      float zsq = 0.f, gsum = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool ok = live && (k4 + e < nz);
        const float grad = g[e] + zz[e];
        zsq += ok ? zz[e] * zz[e] : 0.f;
        gsum += ok ? grad : 0.f;
        zz[e] = ok ? (zz[e] - half_s2 * grad + a.step * nrm[e]) : 0.f;
      }
      if (mine) reinterpret_cast<float4*>(sz + (size_t)c * NZP)[k4 >> 2] = make_float4(zz[0], zz[1], zz[2], zz[3]);
      if (a.trace != nullptr && rank == 0) {
        zsq = warp_sum(zsq);
        gsum = warp_sum(gsum);
        if ((tid & 31) == 0) { atomicAdd(&sred[CH], zsq); atomicAdd(&sred[CH + 1], gsum); }
      }
    }
    __syncthreads();
    if (a.trace != nullptr && rank == 0 && tid == 0) {
      float en = 0.f;
      for (int c = 0; c < nvalid; ++c) en += sred[c] + b3;
      float* tr = a.trace + (size_t)it * a.trace_stride;
      if (a.use_ebm) atomicAdd(&tr[0], en);
      if (a.trace_stride == 2) {
        atomicAdd(&tr[1], 0.5f * sred[CH]);
      } else {
        atomicAdd(&tr[2], 0.5f * sred[CH]);
        atomicAdd(&tr[3], sred[CH + 1] * a.inv_count);
      }
    }
    if (a.trace != nullptr) __syncthreads();  // thread 0 has consumed sred before the next iteration re-zeroes it
  }
  if (rank == 0) {
    for (int i = tid; i < CH * NZP; i += blockDim.x) {
      const int c = i / NZP, k = i - c * NZP;
      if (c < nvalid && k < nz) a.z[(size_t)(c0 + c) * nz + k] = sz[i];
    }
  }
  cluster.sync();  // keep both CTAs' shared memory alive until all DSMEM traffic has landed
}

// chains per 2-CTA cluster: 8 while the batch is small (more clusters, shorter per-cluster latency: 0.70 ms vs 1.17 ms for
// 256 chains x 60 steps), 16 for large batches (every weight read from smem feeds twice the FMAs: 62.6 vs 52.7 M chain-steps/s
// at 16 384 chains)
static constexpr int kChains = 8, kChainsWide = 16;

int launch_ebm_langevin(const MlpPack* m, float* z, int B, int K, float step, int with_noise, const float* noise,
                        uint64_t seed, uint64_t chain0, uint64_t step0, float* trace, int trace_stride,
                        const float* gpart, int nsplit, int gstride, int nz_if_no_ebm, cudaStream_t stream) {
  EbmArgs a{};
  a.use_ebm = m != nullptr;
  if (m) {
    a.W1 = m->W1; a.b1 = m->b1; a.W2 = m->W2; a.b2 = m->b2; a.w3 = m->w3; a.b3 = m->b3;
    a.nz = m->nz; a.ndf = m->ndf; a.slope = m->slope;
  } else {
    a.nz = nz_if_no_ebm; a.ndf = 2; a.slope = 1.f;
  }
  if (a.nz < 1 || a.nz > 128) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM Langevin kernel supports 1 <= nz <= 128 (got %d)", a.nz);
  if (a.ndf < 2 || a.ndf > 256) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM Langevin kernel supports ndf <= 256 (got %d)", a.ndf);
  if (B <= 0 || K < 0) DAMC_FAIL(DAMC_ERR_INVALID, "B must be > 0 and K >= 0 (B=%d K=%d)", B, K);
  a.z = z; a.B = B; a.K = K; a.step = step; a.with_noise = with_noise; a.noise = noise;
  a.seed = seed; a.chain0 = chain0; a.step0 = step0; a.trace = trace; a.trace_stride = trace_stride;
  a.gpart = gpart; a.nsplit = nsplit; a.gstride = gstride; a.inv_count = 1.0f / ((float)B * (float)a.nz);
  const bool wide = B >= 4096 && ebm_plan<kChainsWide>(a.nz, a.ndf).total <= 227 * 1024;
  const EbmSmemPlan P = wide ? ebm_plan<kChainsWide>(a.nz, a.ndf) : ebm_plan<kChains>(a.nz, a.ndf);
  if (P.total > 227 * 1024) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM weights need %zu B of shared memory per CTA", P.total);
  if (wide) {
    DAMC_CUDA(cudaFuncSetAttribute(ebm_langevin_kernel<kChainsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));
    ebm_langevin_kernel<kChainsWide><<<2 * ceil_div(B, kChainsWide), 256, P.total, stream>>>(a);
  } else {
    DAMC_CUDA(cudaFuncSetAttribute(ebm_langevin_kernel<kChains>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P.total));
    ebm_langevin_kernel<kChains><<<2 * ceil_div(B, kChains), 256, P.total, stream>>>(a);
  }
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}

// ---- single step, many chains: weights streamed from L2 (both orientations), chains tiled CH per CTA -----------------------
// Used as the tail of every posterior Langevin step (reference MCMC.py:57-64): with K = 1 there is nothing to amortise a
// 131 KB shared-memory fill over, so thread j walks column j of the transposed weights with coalesced loads instead.
struct EbmStepArgs {
  const float *W1, *W1T, *b1, *W2, *W2T, *b2, *w3, *b3;
  int nz, ndf, use_ebm;
  float slope;
  float* z;
  int B;
  float step;
  int with_noise;
  const float* noise;  // [B,nz] for this step or null
  uint64_t seed, chain0, step_index;
  const unsigned long long* seed_ptr;   // non-null: {seed, chain0, step0} read from device memory (CUDA-graph replays)
  float* trace;        // null or [4]: sum E, (llhd), |z|^2/2, mean grad
  const float* gpart;
  int nsplit, gstride;
  float gpart_scale;   // multiplies the summed generator partials (undoes the fp16 mode's sigma^2 scaling)
  float inv_count;
};

// acc[c] += sum_i W[i*ld + col] * vec[i*CH + c].  The weight column is fetched in batches of U independent loads (32,
// then 8, then 1): with ~1 us of L2/HBM latency per round trip the number of round trips, not the FMA count, is the cost.
template <int CH, int U>
__device__ __forceinline__ int stream_matvec_batch(const float* __restrict__ Wcol, int ld, int i, int n, const float* vec,
                                                   float (&acc)[CH]) {
  for (; i + U <= n; i += U) {
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) w[u] = __ldg(Wcol + (size_t)(i + u) * ld);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int c4 = 0; c4 < CH / 4; ++c4) {
        const float4 v = *reinterpret_cast<const float4*>(vec + (i + u) * CH + 4 * c4);
        acc[4 * c4] = fmaf(w[u], v.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w[u], v.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w[u], v.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w[u], v.w, acc[4 * c4 + 3]);
      }
    }
  }
  return i;
}
template <int CH>
__device__ __forceinline__ void stream_matvec(const float* __restrict__ Wcol, int ld, int n, const float* vec, float (&acc)[CH]) {
  int i = stream_matvec_batch<CH, 32>(Wcol, ld, 0, n, vec, acc);
  i = stream_matvec_batch<CH, 8>(Wcol, ld, i, n, vec, acc);
  stream_matvec_batch<CH, 1>(Wcol, ld, i, n, vec, acc);
}

template <int CH>
__global__ void __launch_bounds__(256) ebm_step_kernel(const EbmStepArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int nz = a.nz, ndf = a.ndf, tid = threadIdx.x;
  float* zs = sm;                    // [nz][CH]
  float* a1s = zs + nz * CH;         // [ndf][CH]
  float* d2s = a1s + ndf * CH;       // [ndf][CH]
  float* d1s = d2s + ndf * CH;       // [ndf][CH]
  __shared__ float red[3];
  const int c0 = blockIdx.x * CH;
  const int nvalid = min(CH, a.B - c0);
  if (tid < 3) red[tid] = 0.f;
  if (a.use_ebm) {  // pull the whole weight slab (W1 .. W2T are one allocation, ~0.5 MB) into L2 ahead of the mat-vecs
    const char* base = reinterpret_cast<const char*>(a.W1);
    const size_t bytes = (size_t)(a.W2T + (size_t)ndf * ndf - a.W1) * sizeof(float);
    for (size_t off = (size_t)tid * 128; off < bytes; off += (size_t)blockDim.x * 128)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
  }
  for (int i = tid; i < nz * CH; i += blockDim.x) {
    const int c = i / nz, k = i - c * nz;  // coalesced over k
    zs[k * CH + c] = c < nvalid ? a.z[(size_t)(c0 + c) * nz + k] : 0.f;
  }
  __syncthreads();
  float gE[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) gE[c] = 0.f;
  if (a.use_ebm) {
    float acc[CH];
    unsigned m1 = 0u;
    if (tid < ndf) {
      const float bb = a.b1[tid];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = bb;
      stream_matvec<CH>(a.W1T + tid, ndf, nz, zs, acc);
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        if (acc[c] > 0.f) m1 |= 1u << c;
        a1s[tid * CH + c] = acc[c] > 0.f ? acc[c] : a.slope * acc[c];
      }
    }
    __syncthreads();
    if (tid < ndf) {
      const float bb = a.b2[tid], w3 = a.w3[tid];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = bb;
      stream_matvec<CH>(a.W2T + tid, ndf, ndf, a1s, acc);
      float e_part = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        d2s[tid * CH + c] = (acc[c] > 0.f ? 1.f : a.slope) * w3;
        if (c < nvalid) e_part += w3 * (acc[c] > 0.f ? acc[c] : a.slope * acc[c]);
      }
      if (a.trace != nullptr) atomicAdd(&red[0], e_part);
    }
    __syncthreads();
    if (tid < ndf) {
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = 0.f;
      stream_matvec<CH>(a.W2 + tid, ndf, ndf, d2s, acc);
#pragma unroll
      for (int c = 0; c < CH; ++c) d1s[tid * CH + c] = ((m1 >> c) & 1u) ? acc[c] : a.slope * acc[c];
    }
    __syncthreads();
    if (tid < nz) {
      stream_matvec<CH>(a.W1 + tid, nz, ndf, d1s, gE);
    }
  }
  // fused update: thread k owns latent dimension k of the CH chains
  float zsq = 0.f, gsum = 0.f;
  if (tid < nz) {
    const float half_s2 = 0.5f * a.step * a.step;
    // split-K partial sums of the generator gradient: 8 splits x CH chains = 32 independent loads per round trip, summed in
    // ascending split order per chain (deterministic)
    float gG[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) gG[c] = 0.f;
    if (a.gpart != nullptr) {
      constexpr int SU = 32 / CH;   // splits per round trip: 32 independent loads in flight whatever the chain tile
      for (int s0 = 0; s0 < a.nsplit; s0 += SU) {
        float v[SU][CH];
#pragma unroll
        for (int u = 0; u < SU; ++u)
#pragma unroll
          for (int c = 0; c < CH; ++c)
            v[u][c] = (s0 + u < a.nsplit && c < nvalid) ? a.gpart[((size_t)(s0 + u) * a.B + (size_t)(c0 + c)) * a.gstride + tid] : 0.f;
#pragma unroll
        for (int u = 0; u < SU; ++u)
#pragma unroll
          for (int c = 0; c < CH; ++c) gG[c] += v[u][c];
      }
    }
    for (int c = 0; c < nvalid; ++c) {
      const size_t chain = (size_t)(c0 + c);
      float g = gE[c];
      if (a.gpart != nullptr) g = fmaf(gG[c], a.gpart_scale, g);
      const float zv = zs[tid * CH + c];
      const float grad = g + zv;
      float nrm = 0.f;
      if (a.with_noise)
        nrm = a.noise ? a.noise[chain * nz + tid]
                      : (a.seed_ptr   // replayed graphs: (seed, chain0, step0) live in device memory, step_index is relative
                             ? philox_normal1(a.seed_ptr[0], a.seed_ptr[1] + chain, a.seed_ptr[2] + a.step_index, (uint32_t)tid)
                             : philox_normal1(a.seed, a.chain0 + chain, a.step_index, (uint32_t)tid));
      a.z[chain * nz + tid] = zv - half_s2 * grad + a.step * nrm;
      zsq += zv * zv;
      gsum += grad;
    }
  }
  if (a.trace != nullptr) {
    zsq = warp_sum(zsq);
    gsum = warp_sum(gsum);
    if ((tid & 31) == 0) { atomicAdd(&red[1], zsq); atomicAdd(&red[2], gsum); }
    __syncthreads();
    if (tid == 0) {
      if (a.use_ebm) atomicAdd(&a.trace[0], red[0] + (float)nvalid * a.b3[0]);
      atomicAdd(&a.trace[2], 0.5f * red[1]);
      atomicAdd(&a.trace[3], red[2] * a.inv_count);
    }
  }
}

template <int CH>
static int launch_ebm_step_ch(const MlpPack* m, float* z, int B, float step, int with_noise, const float* noise, uint64_t seed,
                              uint64_t chain0, uint64_t step_index, float* trace4, const float* gpart, int nsplit, int gstride,
                              float gpart_scale, int nz_if_no_ebm, cudaStream_t stream, const unsigned long long* seed_ptr) {
  EbmStepArgs a{};
  a.use_ebm = m != nullptr;
  if (m) {
    a.W1 = m->W1; a.W1T = m->W1T; a.b1 = m->b1; a.W2 = m->W2; a.W2T = m->W2T; a.b2 = m->b2; a.w3 = m->w3; a.b3 = m->b3;
    a.nz = m->nz; a.ndf = m->ndf; a.slope = m->slope;
  } else {
    a.nz = nz_if_no_ebm; a.ndf = 0; a.slope = 1.f;
  }
  if (a.nz < 1 || a.nz > 256 || a.ndf > 256) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM step kernel: nz <= 256 and ndf <= 256 (nz=%d ndf=%d)", a.nz, a.ndf);
  a.z = z; a.B = B; a.step = step; a.with_noise = with_noise; a.noise = noise; a.seed = seed; a.chain0 = chain0;
  a.seed_ptr = seed_ptr;
  a.step_index = step_index; a.trace = trace4; a.gpart = gpart; a.nsplit = nsplit; a.gstride = gstride;
  a.gpart_scale = gpart_scale;
  a.inv_count = 1.0f / ((float)B * (float)a.nz);
  const size_t smem = sizeof(float) * CH * ((size_t)a.nz + 3 * (size_t)a.ndf);
  if (smem > 48 * 1024) DAMC_CUDA(cudaFuncSetAttribute(ebm_step_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ebm_step_kernel<CH><<<ceil_div(B, CH), 256, smem, stream>>>(a);
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}

// Chain tile per CTA.  Every CTA streams the whole MLP (0.5 MB) from L2, so the launch moves B / CH x 0.5 MB: small tiles
// (4 chains: two or more CTAs per SM overlap each other's L2 latency) while the grid is a few waves, larger ones once the
// L2 -> SM stream is the bound (16 384 SVHN chains: 2.1 GB and 359 us with 4-chain tiles).  A chain's arithmetic (ascending-k
// FMA chains, ascending split order) does not depend on the tile, so results are bit-identical across tile sizes.
int launch_ebm_step(const MlpPack* m, float* z, int B, float step, int with_noise, const float* noise, uint64_t seed,
                    uint64_t chain0, uint64_t step_index, float* trace4, const float* gpart, int nsplit, int gstride,
                    float gpart_scale, int nz_if_no_ebm, cudaStream_t stream, const unsigned long long* seed_ptr) {
  static const int force = []{ const char* e = getenv("DAMC_EBM_CH"); return e ? atoi(e) : 0; }();
  // measured at 16 384 SVHN chains: 4-chain tiles 356 us, 8-chain 418 us, 16-chain 255 us
  const int ch = force ? force : (B >= 8192 ? 16 : 4);
  if (ch >= 16)
    return launch_ebm_step_ch<16>(m, z, B, step, with_noise, noise, seed, chain0, step_index, trace4, gpart, nsplit, gstride,
                                  gpart_scale, nz_if_no_ebm, stream, seed_ptr);
  if (ch >= 8)
    return launch_ebm_step_ch<8>(m, z, B, step, with_noise, noise, seed, chain0, step_index, trace4, gpart, nsplit, gstride,
                                 gpart_scale, nz_if_no_ebm, stream, seed_ptr);
  return launch_ebm_step_ch<4>(m, z, B, step, with_noise, noise, seed, chain0, step_index, trace4, gpart, nsplit, gstride,
                               gpart_scale, nz_if_no_ebm, stream, seed_ptr);
}

// Eval consumer (reference eval_anomaly_det.py:115-117): score_b = sum (G(z_b) - x_b)^2 + E(z_b) + |z_b|^2 / 2, the squared
// error arriving as `nparts` partial sums per chain from the generator's score-mode forward.  E(z) is the forward half of
// ebm_step_kernel (thread j owns hidden unit j of the 4 chains of the CTA, weights streamed from L2).
__global__ void __launch_bounds__(256) ebm_score_kernel(const EbmStepArgs a, const float* __restrict__ sq_part, int nparts,
                                                        float* __restrict__ score, float* __restrict__ sqerr) {
  constexpr int CH = 4;
  extern __shared__ __align__(16) float sm[];
  const int nz = a.nz, ndf = a.ndf, tid = threadIdx.x;
  float* zs = sm;               // [nz][CH]
  float* a1s = zs + nz * CH;    // [ndf][CH]
  __shared__ float red[8][CH];
  const int c0 = blockIdx.x * CH;
  const int nvalid = min(CH, a.B - c0);
  for (int i = tid; i < nz * CH; i += blockDim.x) {
    const int c = i / nz, k = i - c * nz;
    zs[k * CH + c] = c < nvalid ? a.z[(size_t)(c0 + c) * nz + k] : 0.f;
  }
  __syncthreads();
  float part[CH];   // per thread: its share of E(z_c) + |z_c|^2 / 2
#pragma unroll
  for (int c = 0; c < CH; ++c) part[c] = tid < nz ? 0.5f * zs[tid * CH + c] * zs[tid * CH + c] : 0.f;
  if (a.use_ebm) {
    float acc[CH];
    if (tid < ndf) {
      const float bb = a.b1[tid];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = bb;
      stream_matvec<CH>(a.W1T + tid, ndf, nz, zs, acc);
#pragma unroll
      for (int c = 0; c < CH; ++c) a1s[tid * CH + c] = acc[c] > 0.f ? acc[c] : a.slope * acc[c];
    }
    __syncthreads();
    if (tid < ndf) {
      const float bb = a.b2[tid], w3 = a.w3[tid];
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] = bb;
      stream_matvec<CH>(a.W2T + tid, ndf, ndf, a1s, acc);
#pragma unroll
      for (int c = 0; c < CH; ++c) part[c] = fmaf(w3, acc[c] > 0.f ? acc[c] : a.slope * acc[c], part[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < CH; ++c) {
    const float v = warp_sum(part[c]);
    if ((tid & 31) == 0) red[tid >> 5][c] = v;
  }
  __syncthreads();
  if (tid < nvalid) {
    float e = a.use_ebm ? a.b3[0] : 0.f;
    for (int w = 0; w < 8; ++w) e += red[w][tid];
    float sq = 0.f;
    for (int j = 0; j < nparts; ++j) sq += sq_part[(size_t)(c0 + tid) * nparts + j];   // ascending block order
    if (sqerr) sqerr[c0 + tid] = sq;
    if (score) score[c0 + tid] = sq + e;
  }
}

int launch_ebm_score(const MlpPack* m, const float* z, int B, int nz, const float* sq_part, int nparts, float* score,
                     float* sqerr, cudaStream_t stream) {
  EbmStepArgs a{};
  a.use_ebm = m != nullptr;
  if (m) {
    a.W1T = m->W1T; a.b1 = m->b1; a.W2T = m->W2T; a.b2 = m->b2; a.w3 = m->w3; a.b3 = m->b3;
    a.nz = m->nz; a.ndf = m->ndf; a.slope = m->slope;
  } else {
    a.nz = nz; a.ndf = 0; a.slope = 1.f;
  }
  if (a.nz < 1 || a.nz > 256 || a.ndf > 256) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "EBM score kernel: nz <= 256 and ndf <= 256 (nz=%d ndf=%d)", a.nz, a.ndf);
  a.z = const_cast<float*>(z); a.B = B;
  const size_t smem = sizeof(float) * 4 * ((size_t)a.nz + (size_t)a.ndf);
  ebm_score_kernel<<<ceil_div(B, 4), 256, smem, stream>>>(a, sq_part, nparts, score, sqerr);
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols,
                                 const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  __shared__ float t[32][33];
  const int x = blockIdx.x * 32 + threadIdx.x, y0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (x < cols && y0 + r < rows) t[r][threadIdx.x] = src[(size_t)(y0 + r) * cols + x];
  __syncthreads();
  const int ox = blockIdx.y * 32 + threadIdx.x, oy0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y)
    if (ox < rows && oy0 + r < cols) dst[(size_t)(oy0 + r) * rows + ox] = t[threadIdx.x][r];
}

int launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t stream, const int* dirty) {
  transpose_kernel<<<dim3(ceil_div(cols, 32), ceil_div(rows, 32)), dim3(32, 8), 0, stream>>>(src, dst, rows, cols, dirty);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

}  // namespace damc
