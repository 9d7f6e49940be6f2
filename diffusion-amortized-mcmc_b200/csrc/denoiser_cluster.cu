// denoiser_cluster.cu -- DAMC ancestral sampler, all T reverse steps in ONE launch on the tensor cores.
//
// Replaces the reverse loop of _netQ_U.forward (reference workspace/src/diffusion_net.py:597-620; Q.p = Diffusion_UnetA
// :463-533, ConcatSquashLinearSkipCtx :417-445; helpers diffusion_helper_func.py:36-70).  Same arithmetic as the per-layer
// tcgen05 path of denoiser_tc.cu (16-bit GEMM operands, fp32 accumulation, fp32 z / phase / update), different schedule:
//
//   * the dependency between two layers is per 128-chain M tile, so an 8-CTA thread-block cluster owns one M tile for all T
//     steps and needs no grid-wide ordering: the CTAs split every layer's N tiles (64 columns = 16 output features as
//     [gate | hyper-bias | main | skip] column blocks; 1 or 2 tiles per CTA), write the next layer's 16-bit operand rows
//     to global memory (they stay in L2) and meet at the hardware cluster barrier; the next layer's TMA loads pick the
//     rows up again.  8 barriers per reverse step replace 8 dependent kernel launches.
//   * per CTA: warp 0 = TMA producer of the operand rows (its 16-row slice of every k-block, multicast into all eight
//     smem rings; one load feeds both of the CTA's tiles), warp 2 = TMA producer of the weights (the 32 rows a k-block
//     feeds, per tile), warp 1 = tcgen05.mma issuer (M = 128, N = 32 half-tile MMAs into the matching half of a TMEM
//     accumulator; the two tiles' chains interleaved), warps 3..10 = epilogue (tcgen05.ld, gate / bias / skip algebra,
//     LeakyReLU, 256-bit row stores; last layer: eps = z + out and the reverse update of z with Philox or injected noise)
//     and, between steps, the fp32 operand preparation (input embedding [sin 2 pi zB, cos 2 pi zB, z] with p.B resident
//     in smem, ctx activations SiLU(cx + ct[t])) for the CTA's 16 chains.
//   * measured (profiles/r01_denoiser_cluster_timeline.txt): a k-block costs 0.34 us (1 tile) / 0.49 us (2 tiles) whatever
//     its bytes, cluster size or number of TMA-issuing threads -- i.e. ~85 ns per tcgen05.mma instruction of the in-order
//     issue thread at these tiny N; with 8 barriers (0.5 us) and 6 us of operand preparation a step costs ~60 us.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "tc_ptx.cuh"

namespace damc {

constexpr int DC_CL = 8;                 // CTAs per cluster (the portable maximum)
constexpr int DC_BM = 128, DC_BN = 64;   // M tile (chains per cluster), N tile (accumulator columns of one tile)
constexpr int DC_Q = DC_BN / 4;          // output features per N tile
constexpr int DC_TPC = 2;                // at most 2 N tiles per CTA and layer (dout <= 256), both fed by ONE load of the operand rows
constexpr int DC_BK = 64;
constexpr int DC_A_BYTES = DC_BM * DC_BK * 2;         // 16 KB
constexpr int DC_B_BYTES = (DC_BN / 2) * DC_BK * 2;   //  4 KB: the half of one weight tile a k-block feeds
constexpr int DC_STAGE = DC_A_BYTES + DC_TPC * DC_B_BYTES;   // 24 KB
constexpr int DC_THREADS = 96 + 256;   // warp 0: operand-row TMA, warp 1: MMA issuer, warp 2: weight TMA, warps 3..10: epilogue
constexpr int DC_CHAINS = DC_BM / DC_CL;              // chains prepared per CTA
constexpr int DC_ZP = DC_CHAINS + 4;                  // pitch of the transposed z tile (16-byte aligned rows)

struct DcLayer {
  int kb_h, kb_total, tpc;              // k-blocks of the h part / all, N tiles per CTA (4*dout / 64 / 8 = 1 or 2)
  int boff;                             // offset of this layer's bias quads inside the smem bias table (floats)
  void* dst1; int ld1, off1;            // leaky_relu(out) -> next layer's operand slice
  void* dst2; int ld2, off2;            // U-net skip copy, or null
};

struct DcParams {
  CUtensorMap tmA[DEN_LAYERS], tmB[DEN_LAYERS];
  DcLayer L[DEN_LAYERS];
  const float* bias4[DEN_LAYERS];
  // operand preparation
  const float* Bp; const float* cx; const float* ct; const float* coef;
  void* A[DEN_LAYERS];
  int ld[DEN_LAYERS], din[DEN_LAYERS], dout[DEN_LAYERS], coff[DEN_LAYERS];
  // update
  float* z; float* eps_out; const float* noise;
  unsigned long long seed, chain0;
  int use_philox, residual;
  int B, T, nsteps, nz, csum, fp16, stages, bias_floats;
  unsigned long long* tlog;   // dbg & 16: globaltimer stamps of cluster 0 / CTA 0 during step 1 (see tools/)
  int dbg;   // timing experiments (env DAMC_DC_DBG): 1 no ctx refresh, 2 no __threadfence, 4 no embedding, 8 no layer epilogue math
  uint32_t idesc;
};

__device__ __forceinline__ float dc_silu(float v) { return v / (1.f + __expf(-v)); }
__device__ __forceinline__ uint16_t dc_cvt(bool fp16, float v) {
  if (fp16) { const __half h = __float2half_rn(v); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&h);
}

__device__ __forceinline__ unsigned long long dc_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define DC_STAMP(slot) do { if (P.tlog && blockIdx.x == 0 && st == 1) P.tlog[(slot)] = dc_now(); } while (0)

__global__ void __launch_bounds__(DC_THREADS, 1) den_cluster_kernel(const __grid_constant__ DcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int nz = P.nz, half = nz >> 1;
  // layout: [ring stages x 24 KB][p.B fp32 nz*half][z tile nz x DC_ZP fp32][bias quads of all layers][barriers]
  float* Bs = reinterpret_cast<float*>(smem_al + (size_t)P.stages * DC_STAGE);
  float* zs = Bs + nz * half;
  float* sbias = zs + nz * DC_ZP;
  const uint32_t bars = smem_base + (uint32_t)P.stages * DC_STAGE + 4u * (uint32_t)(nz * half + nz * DC_ZP + P.bias_floats);
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (P.stages + s); };
  auto bar_tfull = [&](int a) { return bars + 8u * (2 * P.stages + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * P.stages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int b0 = (int)(blockIdx.x / DC_CL) * DC_BM;   // first chain of this cluster's M tile
  const bool fp16 = P.fp16 != 0;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < DEN_LAYERS; ++i) { prefetch_tmap(&P.tmA[i]); prefetch_tmap(&P.tmB[i]); }
    // a stage is free when the MMAs of ALL CTAs of the cluster have read it (each peer multicasts operand rows into it)
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), DC_CL); }
    mbar_init(bar_tfull(0), 1);   // one accumulator set per layer; the cluster barrier between layers orders its reuse
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);   // DC_TPC accumulators of 64 columns
  for (int i = tid; i < nz * half; i += DC_THREADS) Bs[i] = __ldg(P.Bp + i);
  for (int l = 0, o = 0; l < DEN_LAYERS; o += 4 * P.dout[l], ++l)
    for (int i = tid; i < 4 * P.dout[l]; i += DC_THREADS) sbias[o + i] = __ldg(P.bias4[l] + i);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // peers' barriers are initialised before any multicast load / commit can reach them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  int stage = 0;           // smem ring position (producer and MMA issuer each track their own copy)
  uint32_t phase = 0;
  int acc_it = 0;          // layer phases done so far: parity of the accumulator-full barrier
  const int et = tid - 96; // epilogue thread index 0..255 (warps 3..10)

  // c_L = SiLU(cx + ct[irev]) for this CTA's quarter of the chains, written into the [din, din+dout) slice of layer L's
  // operand rows.  Thread = 4 consecutive columns of up to 8 chains per pass: all loads of a pass are issued before use.
  auto ctx_slice = [&](int L, int irev) {
    const int c0g = b0 + rank * DC_CHAINS;
    const int g4 = P.dout[L] >> 2, items = DC_CHAINS * g4;
    const float* ctrow = P.ct + (size_t)irev * P.csum + P.coff[L];
    const float* cxb = P.cx + P.coff[L];
    uint16_t* dst = reinterpret_cast<uint16_t*>(P.A[L]) + P.din[L];
    const int ld = P.ld[L];
    for (int i0 = et; i0 < items; i0 += 8 * 256) {
      float4 x4[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256, c = i / g4, col = (i - c * g4) << 2;
        x4[u] = (i < items && c0g + c < P.B) ? *reinterpret_cast<const float4*>(cxb + (size_t)(c0g + c) * P.csum + col)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * 256, c = i / g4, col = (i - c * g4) << 2;
        if (i < items && c0g + c < P.B) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(ctrow + col));
          const uint2 o = make_uint2(
              (uint32_t)dc_cvt(fp16, dc_silu(x4[u].x + t4.x)) | ((uint32_t)dc_cvt(fp16, dc_silu(x4[u].y + t4.y)) << 16),
              (uint32_t)dc_cvt(fp16, dc_silu(x4[u].z + t4.z)) | ((uint32_t)dc_cvt(fp16, dc_silu(x4[u].w + t4.w)) << 16));
          *reinterpret_cast<uint2*>(dst + (size_t)(c0g + c) * ld + col) = o;
        }
      }
    }
  };

  for (int st = 0; st < P.nsteps; ++st) {
    // ===================== operand preparation for this step (epilogue warps; this CTA's quarter of the chains) ==========
    if (tid == 96) DC_STAMP(0);
    if (warp >= 3) {
      const int c0g = b0 + rank * DC_CHAINS;   // first global chain prepared by this CTA
      const int irev = P.eps_out ? 0 : P.T - 1 - st;
      for (int i0 = et; i0 < DC_CHAINS * nz; i0 += 8 * 256) {   // 8 independent loads per round trip (z comes from L2)
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 256, c = i / nz;
          v[u] = (i < DC_CHAINS * nz && c0g + c < P.B) ? P.z[(size_t)c0g * nz + i] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = i0 + u * 256, c = i / nz, k = i - c * nz;
          if (i < DC_CHAINS * nz) zs[k * DC_ZP + c] = v[u];
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 epilogue warps
      uint16_t* A0 = reinterpret_cast<uint16_t*>(P.A[0]);
      for (int i = et; i < (DC_CHAINS / 8) * half && !(P.dbg & 4); i += 256) {   // phase 2 pi z.B in fp32: it reaches tens of radians
        const int cg = i / half, j = i - cg * half;
        const float* zr = zs + cg * 8;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int k = 0; k < nz; ++k) {
          const float w = Bs[k * half + j];
          const float4 z0 = *reinterpret_cast<const float4*>(zr + k * DC_ZP);
          const float4 z1 = *reinterpret_cast<const float4*>(zr + k * DC_ZP + 4);
          acc[0] = fmaf(z0.x, w, acc[0]); acc[1] = fmaf(z0.y, w, acc[1]); acc[2] = fmaf(z0.z, w, acc[2]); acc[3] = fmaf(z0.w, w, acc[3]);
          acc[4] = fmaf(z1.x, w, acc[4]); acc[5] = fmaf(z1.y, w, acc[5]); acc[6] = fmaf(z1.z, w, acc[6]); acc[7] = fmaf(z1.w, w, acc[7]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int b = c0g + cg * 8 + c;
          if (b < P.B) {
            float sn, cs;
            sincosf(6.283185307179586f * acc[c], &sn, &cs);
            uint16_t* row = A0 + (size_t)b * P.ld[0];
            row[j] = dc_cvt(fp16, sn);
            row[half + j] = dc_cvt(fp16, cs);
          }
        }
      }
      for (int i = et; i < DC_CHAINS * (nz / 2); i += 256) {   // the raw z slice
        const int c = i / (nz / 2), k2 = (i - c * (nz / 2)) * 2;
        if (c0g + c < P.B) {
          const uint32_t o = (uint32_t)dc_cvt(fp16, zs[k2 * DC_ZP + c]) | ((uint32_t)dc_cvt(fp16, zs[(k2 + 1) * DC_ZP + c]) << 16);
          *reinterpret_cast<uint32_t*>(A0 + (size_t)(c0g + c) * P.ld[0] + 2 * half + k2) = o;
        }
      }
      // ctx activations: step 0 prepares every layer; later steps only the last layer's slice (the others were refreshed
      // during the previous step's layer phases, as soon as their readers were done)
      for (int l = (st == 0 ? 0 : DEN_LAYERS - 1); l < DEN_LAYERS; ++l) ctx_slice(l, irev);
      if (tid == 96) DC_STAMP(1);
      // no __threadfence: every reader of these rows (TMA loads, z reads) is in this cluster, and the cluster barrier's
      // release / acquire orders them at cluster scope; the proxy fence hands the generic-proxy writes to the async proxy
      fence_proxy_async_all();
      if (tid == 96) DC_STAMP(2);
    }
    __syncwarp();
    cluster_sync_all();   // every CTA's operand rows are in L2
    if (tid == 96) DC_STAMP(3);

    const float* cf = P.coef + (size_t)st * 8;
    for (int l = 0; l < DEN_LAYERS; ++l) {
      const DcLayer& Ly = P.L[l];
      if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
          fence_proxy_async_all();   // rows written through the generic proxy (by any CTA of the cluster) -> async-proxy reads
          DC_STAMP(8 + 8 * l + 0);
          for (int kb = 0; kb < Ly.kb_total; ++kb) {
            mbar_wait(bar_empty(stage), phase ^ 1u);
            const uint32_t sa = smem_base + (uint32_t)stage * DC_STAGE;
            mbar_expect_tx(bar_full(stage), DC_A_BYTES + Ly.tpc * DC_B_BYTES);
            // every CTA of the cluster multiplies the same 128 operand rows: each loads its 16-row slice once and
            // multicasts it into all eight rings; the rows feed both of this CTA's N tiles
            tma_load_2d_mcast(sa + (uint32_t)rank * (DC_A_BYTES / DC_CL), &P.tmA[l], bar_full(stage), kb * DC_BK,
                              b0 + rank * (DC_BM / DC_CL), (uint16_t)((1u << DC_CL) - 1u));
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
        }
      } else if (warp == 2) {
        // ===================== weight TMA producer (a second issuing thread: one thread sustains ~1 bulk-tensor copy per
        // 0.15-0.19 us, which is what bounded the k-block rate when warp 0 issued the operand rows and the weights) ==========
        if (lane == 0) {
          for (int kb = 0; kb < Ly.kb_total; ++kb) {
            mbar_wait(bar_empty(stage), phase ^ 1u);
            const uint32_t sa = smem_base + (uint32_t)stage * DC_STAGE;
            for (int t = 0; t < Ly.tpc; ++t)
              tma_load_2d(sa + DC_A_BYTES + t * DC_B_BYTES, &P.tmB[l], bar_full(stage), kb * DC_BK,
                          (rank + t * DC_CL) * DC_BN + (kb < Ly.kb_h ? DC_BN / 2 : 0));
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
          for (int kb = 0; kb < Ly.kb_total; ++kb) {
            mbar_wait(bar_full(stage), phase);
            if (kb == 0) DC_STAMP(8 + 8 * l + 1);
            tc_fence_after();
            const uint32_t sa = smem_base + (uint32_t)stage * DC_STAGE;
            const uint64_t adesc = make_sdesc(sa);
            const bool first_kb = kb == 0 || kb == Ly.kb_h;
            // accumulator t, (main, skip) half for an h k-block, (gate, hyper-bias) half for a c k-block; the two tiles'
            // MMAs are independent accumulation chains and are issued interleaved
            const uint32_t d0 = tmem_base + (uint32_t)(kb < Ly.kb_h ? DC_BN / 2 : 0), d1 = d0 + (uint32_t)DC_BN;
            const uint64_t bdesc0 = make_sdesc(sa + DC_A_BYTES), bdesc1 = make_sdesc(sa + DC_A_BYTES + DC_B_BYTES);
            const uint32_t acc0 = first_kb ? 0u : 1u;
            if (Ly.tpc == 2) {
#pragma unroll
              for (int k = 0; k < DC_BK / 16; ++k) {
                umma_bf16(d0, adesc + (uint64_t)(2 * k), bdesc0 + (uint64_t)(2 * k), P.idesc, k > 0 ? 1u : acc0);
                umma_bf16(d1, adesc + (uint64_t)(2 * k), bdesc1 + (uint64_t)(2 * k), P.idesc, k > 0 ? 1u : acc0);
              }
            } else {
#pragma unroll
              for (int k = 0; k < DC_BK / 16; ++k)
                umma_bf16(d0, adesc + (uint64_t)(2 * k), bdesc0 + (uint64_t)(2 * k), P.idesc, k > 0 ? 1u : acc0);
            }
            umma_commit_mcast(bar_empty(stage), (uint16_t)((1u << DC_CL) - 1u));
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_tfull(0));
          DC_STAMP(8 + 8 * l + 2);
        }
      } else {
        // ===================== epilogue warps =====================
        const int q = warp & 3, grp = (warp - 3) >> 2;   // TMEM lane quarter; the two warps of a quarter take one N tile each
        const int b = b0 + q * 32 + lane;
        const bool ok = b < P.B;
        const bool fin = l == DEN_LAYERS - 1;
        DenEpi d;
        if (fin) {
          d.z = P.z; d.eps_out = P.eps_out; d.nz = nz; d.residual = P.residual; d.use_philox = P.use_philox;
          d.noise = P.noise ? P.noise + (size_t)st * P.B * nz : nullptr;
          d.seed = P.seed; d.seed_ptr = nullptr; d.chain0 = P.chain0; d.step = (unsigned long long)st;
          d.c_pred = __ldg(cf + 0); d.c_eps = __ldg(cf + 1); d.c_zt = __ldg(cf + 2); d.c_x = __ldg(cf + 3); d.c_std = __ldg(cf + 4);
          d.last = __ldg(cf + 5) != 0.f;
        }
        // while the producer / MMA warps work on this layer: layer l-1's ctx slice for the NEXT step (its readers, the
        // TMA loads of layer l-1 in this step, finished before the cluster barrier that opened this phase)
        if (l > 0 && st + 1 < P.nsteps && !(P.dbg & 1)) ctx_slice(l - 1, P.T - 2 - st);
        // the two warps of a TMEM lane quarter take one N tile each (tile rank + grp * 8; 16 features)
        for (int one = 0; one < 1; ++one) {
          if (grp >= Ly.tpc) break;
          const int nt = rank + grp * DC_CL;
          mbar_wait(bar_tfull(0), (uint32_t)acc_it & 1u);
          if (tid == 96) DC_STAMP(8 + 8 * l + 3);
          tc_fence_after();
          const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * DC_BN);
          uint32_t vg[16], vh[16], vm[16], vs[16];
          tmem_ld16(t0, vg);
          tmem_ld16(t0 + (uint32_t)DC_Q, vh);
          tmem_ld16(t0 + (uint32_t)(2 * DC_Q), vm);
          tmem_ld16(t0 + (uint32_t)(3 * DC_Q), vs);
          tmem_ld_wait();
          tc_fence_before();
          if (P.dbg & 8) continue;
          const int f0 = nt * DC_Q;
          const float* sb = sbias + Ly.boff + 4 * f0;
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 bq = *reinterpret_cast<const float4*>(sb + 4 * i);   // (bg, 0, b, bs)
            const float gate = __uint_as_float(vg[i]) + bq.x;
            o[i] = fmaf(__uint_as_float(vm[i]) + bq.z, __fdividef(1.f, 1.f + __expf(-gate)),
                        __uint_as_float(vh[i]) + __uint_as_float(vs[i]) + bq.w);
          }
          if (!ok) continue;
          if (fin) {
#pragma unroll
            for (int j = 0; j < 4; ++j) den_final_quad(d, b, f0 + 4 * j, o + 4 * j);
          } else {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = o[2 * j], c = o[2 * j + 1];
              a = a > 0.f ? a : 0.01f * a;
              c = c > 0.f ? c : 0.01f * c;
              w[j] = pack2(fp16, a, c);
            }
            uint16_t* o1 = reinterpret_cast<uint16_t*>(Ly.dst1) + (long long)b * Ly.ld1 + Ly.off1 + f0;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o1), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                         "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            if (Ly.dst2) {
              uint16_t* o2 = reinterpret_cast<uint16_t*>(Ly.dst2) + (long long)b * Ly.ld2 + Ly.off2 + f0;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o2), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                           "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
          }
        }
        if (tid == 96) DC_STAMP(8 + 8 * l + 4);
        if (P.dbg & 2) __threadfence();
        fence_proxy_async_all();
        if (tid == 96) DC_STAMP(8 + 8 * l + 5);
      }
      ++acc_it;
      __syncwarp();
      cluster_sync_all();   // layer l complete in every CTA of the cluster (its rows / the new z are visible in L2)
      if (tid == 96) DC_STAMP(8 + 8 * l + 6);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// Policy (measured, T = 100, profiles/r01_denoiser_timings.txt): 6.0 ms at 128 chains and 6.1 ms at 1 024 against 6.5 / 7.0 ms
// for the per-layer launches (CUDA-graph replayed); with more than 8 clusters in flight the clusters contend (12 ms at
// 2 304 chains) and the per-layer path, which fills all SMs per layer, wins.  DAMC_DEN_CLUSTER=0 disables the kernel,
// =2 forces it for every B.
bool den_cluster_supported(const DenPack* d, int B) {
  static const int mode = []{ const char* e = getenv("DAMC_DEN_CLUSTER"); return e ? atoi(e) : 1; }();
  if (mode == 0 || d->nz > 128 || d->nz % 8) return false;
  for (int i = 0; i < DEN_LAYERS; ++i)
    // 4*dout/64 N tiles must split evenly over the 8 CTAs (lockstep rings), at most DC_TPC per CTA
    if (d->din[i] % 64 || d->dout[i] % 128 || d->dout[i] > 128 * DC_TPC) return false;
  if (mode == 2) return true;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  (void)sms;
  return ceil_div(B, DC_BM) <= 8;
}

int den_cluster_run(const DenPack* d, int precision, const DenWs& w, float* z, float* eps_out, int B, int T, int nsteps,
                    const float* host_coef, const float* noise, int use_philox, uint64_t seed, uint64_t chain0,
                    cudaStream_t s) {
  DAMC_TRY(den_tc_pack_bn(d, precision, 2, s));
  const DenTcPack* t = d->tc[precision];
  DcParams P{};
  const int fp16 = precision == DAMC_PREC_FP16;
  const int skip_to[3] = {6, 5, 4};
  int boff = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int kt = d->din[i] + d->dout[i];
    DAMC_TRY(tc_encode_2d(&P.tmA[i], fp16, w.A[i], kt, B, DC_BM / DC_CL));   // box = one CTA's 16-row slice
    DAMC_TRY(tc_encode_2d(&P.tmB[i], fp16, t->Wq[2][i], kt, 4 * d->dout[i], DC_BN / 2));   // bn = 64 row order
    DcLayer& L = P.L[i];
    L.kb_h = d->din[i] / DC_BK; L.kb_total = kt / DC_BK; L.tpc = 4 * d->dout[i] / DC_BN / DC_CL; L.boff = boff;
    boff += 4 * d->dout[i];
    if (i < DEN_LAYERS - 1) {
      L.dst1 = w.A[i + 1]; L.ld1 = d->din[i + 1] + d->dout[i + 1]; L.off1 = 0;
      if (i < 3) { const int j = skip_to[i]; L.dst2 = w.A[j]; L.ld2 = d->din[j] + d->dout[j]; L.off2 = d->dout[j - 1]; }
    }
    P.bias4[i] = t->bias4[i];
    P.A[i] = w.A[i]; P.ld[i] = kt; P.din[i] = d->din[i]; P.dout[i] = d->dout[i]; P.coff[i] = d->coff[i];
  }
  P.bias_floats = boff;
  P.Bp = d->Bp; P.cx = w.cx; P.ct = w.ct; P.coef = w.coef;
  P.z = z; P.eps_out = eps_out; P.noise = noise; P.seed = seed; P.chain0 = chain0; P.use_philox = use_philox;
  P.residual = d->residual; P.B = B; P.T = T; P.nsteps = nsteps; P.nz = d->nz; P.csum = d->csum; P.fp16 = fp16;
  P.dbg = getenv("DAMC_DC_DBG") ? atoi(getenv("DAMC_DC_DBG")) : 0;
  static unsigned long long* tlog = nullptr;
  if ((P.dbg & 16) && !tlog) { cudaMalloc(&tlog, 128 * 8); cudaMemset(tlog, 0, 128 * 8); }
  P.tlog = (P.dbg & 16) ? tlog : nullptr;
  const uint32_t opfmt = fp16 ? 0u : 1u;
  P.idesc = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)((DC_BN / 2) >> 3) << 17) | ((uint32_t)(DC_BM >> 4) << 24);
  const size_t fixed = 4 * ((size_t)d->nz * (d->nz / 2) + (size_t)d->nz * DC_ZP + boff) + 8 * (2 * 8 + 4) + 16 + 1024 + 64;
  P.stages = (int)std::min<size_t>(8, (227 * 1024 - fixed) / DC_STAGE);
  if (P.stages < 3) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "denoiser (cluster kernel): not enough shared memory for a pipeline");
  const size_t smem = (size_t)P.stages * DC_STAGE + fixed;
  DAMC_CUDA(cudaMemcpyAsync(w.coef, host_coef, sizeof(float) * 8 * (size_t)nsteps, cudaMemcpyHostToDevice, s));
  DAMC_CUDA(cudaFuncSetAttribute(den_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = DC_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.gridDim = dim3(DC_CL * ceil_div(B, DC_BM));
  cfg.blockDim = dim3(DC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  profile_mark(s, true);
  DAMC_CUDA(cudaLaunchKernelEx(&cfg, den_cluster_kernel, P));
  profile_mark(s, false);
  count_launch(3);
  if (P.tlog) {   // experiment mode: dump the stamps of step 1 (ns relative to the step start)
    unsigned long long h[128];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, P.tlog, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[dc] prep: compute %llu fence %llu barrier %llu\n", h[1] - h[0], h[2] - h[1], h[3] - h[2]);
    for (int l = 0; l < DEN_LAYERS; ++l) {
      const unsigned long long* e = h + 8 + 8 * l;
      const unsigned long long start = l == 0 ? h[3] : h[8 + 8 * (l - 1) + 6];
      fprintf(stderr, "[dc] layer %d: issue +%llu  first-full +%llu  mma-done +%llu  epi-wake +%llu  epi-done +%llu  fence +%llu  barrier +%llu\n", l,
              e[0] - start, e[1] - start, e[2] - start, e[3] - start, e[4] - start, e[5] - start, e[6] - start);
    }
  }
  return DAMC_OK;
}

}  // namespace damc
