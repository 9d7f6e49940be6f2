// gen_last.cu -- the generator's LAST layer of one Langevin step as ONE kernel: x_hat = tanh(ConvT(a) + b), the Gaussian
// likelihood gradient (x_hat - x)/sigma^2 (1 - x_hat^2), and the layer's input-gradient dU/da (masked by the LeakyReLU
// sign bits of a) -- without materialising the scatter-form product Y or the im2col'd gradient in HBM.
//
// Replaces, for the layer netG.gen[-2] (ConvTranspose2d(C -> nc, k3-s1-p1 | k4-s2-p1)) + Tanh, the forward, the
// likelihood term and autograd's backward through them in reference workspace/src/MCMC.py:55-60 (diffusion_net.py:40-45).
// The three-launch form it supersedes (scatter GEMM -> last_finish_kernel -> K = 64 dgrad GEMM) moved 1.66 GB per step at
// 1 024 CIFAR-10 chains; this kernel reads a (0.54 GB) and x once and writes dU/da (0.54 GB) once.
//
// A CTA owns a block of input rows of one image (the whole image when it fits in shared memory) and runs three phases,
// pipelined across consecutive blocks by dedicated warp groups:
//   P1  scatter GEMM   Y[pix, (kh,kw,co)] = a[pix, :] . W[:, co, kh, kw]   tcgen05, A tiles by TMA, W resident in smem,
//       accumulators in TMEM; the S warps add each Y row into the fp32 pre-activation image in smem (col2im) in
//       barrier-separated passes whose writes never collide (fixed order => bit-reproducible sums);
//   P2  S warps: tanh, x_hat / loss output, likelihood gradient g (operand type) into smem;
//   P3  E warps gather g into the K-major SWIZZLE_128B operand tile [128 pixels][64 = (kh,kw,co) slots] in smem; tcgen05
//       multiplies it with the resident dgrad weights [C][64]; the E warps scale by 1 | slope from the 1-bit masks and
//       store dU/da rows with 256-bit stores (parity-planar or flat layout, as the next dgrad GEMM reads it).
// Row blocks that do not cover the image recompute the scatter GEMM of `halo` neighbouring tiles on each side.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <atomic>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "tc_ptx.cuh"

namespace damc {

constexpr int LF_THREADS = 640;           // warp 0 TMA, 1 scatter-MMA, 2 dgrad-MMA, 3 idle, 4-7 S group A, 8-15 E group, 16-19 S group B
constexpr int LF_TILE_BYTES = 128 * 128;  // 128 rows x one 128-byte swizzle row (64 16-bit elements)
constexpr int LF_MAX_STAGES = 6;
constexpr int LF_MAX_CHUNKS = 4;         // dgrad N chunks per tile (C <= 512)

struct LastParams {
  int B, C, Hi, Wi, Ho, Wo, pad;
  int Np_sc;                   // scatter columns k*k*nc padded to a multiple of 16
  int Ht, tile_rows, tpi;      // input rows per 128-row tile, live rows of a tile, tiles per image
  int Rt, nblk_img, nblocks;   // tiles per block, blocks per image, blocks in the launch
  int halo_t;                  // recomputed tiles on each side of a row block
  int kb_sc;                   // C / 64
  int chunkN, nchunks;         // dgrad N chunk (TMEM accumulator stage) and chunks per tile
  int stages;
  uint32_t idesc_sc, idesc_dg;
  int op_fp16;
  uint32_t off_a2, off_wsc, off_wdg, off_out, off_g, off_x, off_bar;   // bytes from the 1024-aligned smem base
  int out_floats, g_bytes;
  int wo_shift;                // log2(Wo) or -1
  int xr;                      // output rows of x held in smem at a time (>= n_need: the whole block in one chunk)
  int row_in_warp;             // 32 % Wi == 0: an image row never straddles two warps of the S group
  int dual;                    // two pre-activation images: S group A accumulates the even tiles, B the odd ones (no races between them)
  int debug;                   // DAMC_LAST_DEBUG=1: CTA 0 prints where its warp groups spent their clocks
  const float* x;              // [B][nc][Ho][Wo]
  float* xhat;                 // same shape or null
  float* loss;                 // scalar accumulator or null
  const float* bias;           // [nc]
  float inv_sigma2, gscale, slope;
  const uint32_t* maskbits;    // sign bits of a (NHWC element order), 1 = pre-activation > 0
  void* gout;                  // dU/da, operand type
  int planar_out;
  // score mode (damc_posterior_score): forward + squared residual only -- no gradient image, no dgrad; every block writes
  // the sum of (x_hat - x)^2 over the output rows it owns to sq_part[b * nblk_img + block-in-image]
  int do_dgrad;
  float* sq_part;
  uint32_t off_red;
};

// 1-D bulk copy global -> shared (bytes % 16 == 0, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"((uint64_t)src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

#define LF_T(acc, stmt) do { if (P.debug) { const uint32_t _t = (uint32_t)clock(); stmt; acc += (uint32_t)clock() - _t; } else { stmt; } } while (0)

struct LastBlock {
  int b, t0, t1, pt0, pt1, r0, r1, need_lo, n_need;
};

template <int K, int S>
__device__ __forceinline__ LastBlock last_block(const LastParams& P, int blk) {
  LastBlock L;
  L.b = blk / P.nblk_img;
  const int j = blk - L.b * P.nblk_img;
  L.t0 = j * P.Rt;
  L.t1 = min(P.tpi, L.t0 + P.Rt);
  L.pt0 = max(0, L.t0 - P.halo_t);
  L.pt1 = min(P.tpi, L.t1 + P.halo_t);
  L.r0 = L.t0 * P.Ht;
  L.r1 = L.t1 * P.Ht;
  L.need_lo = L.r0 * S - P.pad;                                   // first output row the block's dgrad reads
  L.n_need = (L.r1 - 1) * S - P.pad + K - 1 - L.need_lo + 1;      // ... and how many (rows outside the image are zero)
  return L;
}

template <int K, int S, int NC>
__global__ void __launch_bounds__(LF_THREADS, 1)
last_fused_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmWsc,
                  const __grid_constant__ CUtensorMap tmWdg, const __grid_constant__ LastParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + P.off_bar;
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (LF_MAX_STAGES + s); };
  const uint32_t bar_w = bars + 8u * (2 * LF_MAX_STAGES);
  auto bar_yfull = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 1 + a); };
  auto bar_yempty = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 3 + a); };
  const uint32_t bar_gfull = bars + 8u * (2 * LF_MAX_STAGES + 5);
  const uint32_t bar_gempty = bars + 8u * (2 * LF_MAX_STAGES + 6);
  auto bar_a2full = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 7 + a); };
  auto bar_a2empty = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 9 + a); };
  auto bar_t3full = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 11 + a); };
  auto bar_t3empty = [&](int a) { return bars + 8u * (2 * LF_MAX_STAGES + 13 + a); };
  const uint32_t bar_xfull = bars + 8u * (2 * LF_MAX_STAGES + 15);
  const uint32_t tmem_slot = bars + 8u * (2 * LF_MAX_STAGES + 16);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + P.off_bar + 8u * (2 * LF_MAX_STAGES + 16));
  float* const out_s = reinterpret_cast<float*>(gen_base + P.off_out);
  uint8_t* const g_s = gen_base + P.off_g;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmWsc);
    prefetch_tmap(&tmWdg);
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    mbar_init(bar_w, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_yfull(a), 1); mbar_init(bar_yempty(a), 4);
      mbar_init(bar_a2full(a), 8); mbar_init(bar_a2empty(a), 1);
      mbar_init(bar_t3full(a), 1); mbar_init(bar_t3empty(a), 8);
    }
    mbar_init(bar_gfull, 8);
    mbar_init(bar_gempty, 8);
    mbar_init(bar_xfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // bias-initialised pre-activation image and an all-zero gradient image (its border columns stay zero for the whole launch)
  float bias_r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) bias_r[c] = __ldg(P.bias + c);   // written by the pack kernels of an earlier, completed launch
  for (int i = threadIdx.x; i < P.out_floats; i += LF_THREADS) out_s[i] = __ldg(P.bias + i % NC);
  if (P.dual) for (int i = threadIdx.x; i < P.out_floats; i += LF_THREADS) out_s[P.out_floats + i] = 0.f;
  for (int i = threadIdx.x; i < P.g_bytes / 4; i += LF_THREADS) reinterpret_cast<uint32_t*>(g_s)[i] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  const uint32_t tmem_y = tmem_base, tmem_d = tmem_base + 128u;   // 2 x 64 scatter columns | 2 x 128 dgrad columns
  const int gpitch = (P.Wo + 2) * 8;                               // bytes per row of the gradient image (4 x 16-bit per pixel)

  if (warp == 0) {
    // ===================== TMA producer: resident weights once, then the activation tiles of every block ==========
    if (lane == 0) {
      mbar_expect_tx(bar_w, (uint32_t)(P.kb_sc * P.Np_sc * 128 + P.C * 128));
      for (int kb = 0; kb < P.kb_sc; ++kb) tma_load_2d(base + P.off_wsc + (uint32_t)(kb * P.Np_sc * 128), &tmWsc, bar_w, kb * 64, 0);
      for (int ch = 0; ch < P.nchunks; ++ch) tma_load_2d(base + P.off_wdg + (uint32_t)(ch * P.chunkN * 128), &tmWdg, bar_w, 0, ch * P.chunkN);
      int stage = 0;
      uint32_t phase = 0;
      for (int blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x) {
        const LastBlock L = last_block<K, S>(P, blk);
        for (int t = L.pt0; t < L.pt1; ++t)
          for (int kb = 0; kb < P.kb_sc; ++kb) {
            mbar_wait(bar_empty(stage), phase ^ 1u);
            mbar_expect_tx(bar_full(stage), (uint32_t)P.tile_rows * 128u);
            tma_load_5d(base + (uint32_t)stage * LF_TILE_BYTES, &tmA, bar_full(stage), kb * 64, 0, t * P.Ht, L.b, 0);
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    // ===================== scatter-GEMM issuer =====================
    if (lane == 0) {
      mbar_wait(bar_w, 0u);
      tc_fence_after();
      int stage = 0, it = 0;
      uint32_t phase = 0, dbg0 = 0, dbg1 = 0;
      const uint32_t tstart = (uint32_t)clock();
      for (int blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x) {
        const LastBlock L = last_block<K, S>(P, blk);
        for (int t = L.pt0; t < L.pt1; ++t, ++it) {
          const int as = it & 1;
          LF_T(dbg0, mbar_wait(bar_yempty(as), ((uint32_t)(it >> 1) & 1u) ^ 1u));
          tc_fence_after();
          for (int kb = 0; kb < P.kb_sc; ++kb) {
            LF_T(dbg1, mbar_wait(bar_full(stage), phase));
            tc_fence_after();
            const uint64_t adesc = make_sdesc(base + (uint32_t)stage * LF_TILE_BYTES);
            const uint64_t bdesc = make_sdesc(base + P.off_wsc + (uint32_t)(kb * P.Np_sc * 128));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tmem_y + (uint32_t)as * 64u, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), P.idesc_sc,
                        (kb > 0 || k4 > 0) ? 1u : 0u);
            umma_commit(bar_empty(stage));
            if (++stage == P.stages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(bar_yfull(as));
        }
      }
      if (P.debug && blockIdx.x == 0)
        printf("last_fused CTA0 scatter-MMA: total %u clk, wait yempty %u, wait A tiles %u, tiles %d\n", (uint32_t)clock() - tstart, dbg0, dbg1, it);
    }
  } else if (warp == 2) {
    // ===================== dgrad-GEMM issuer =====================
    if (lane == 0 && P.do_dgrad) {
      mbar_wait(bar_w, 0u);
      tc_fence_after();
      int jt = 0, jc = 0;
      uint32_t dbg0 = 0, dbg1 = 0;
      for (int blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x) {
        const LastBlock L = last_block<K, S>(P, blk);
        for (int t = L.t0; t < L.t1; ++t, ++jt) {
          const int buf = jt & 1;
          LF_T(dbg0, mbar_wait(bar_a2full(buf), (uint32_t)(jt >> 1) & 1u));
          tc_fence_after();
          const uint64_t adesc = make_sdesc(base + P.off_a2 + (uint32_t)buf * LF_TILE_BYTES);
          for (int ch = 0; ch < P.nchunks; ++ch, ++jc) {
            const int st = jc & 1;
            LF_T(dbg1, mbar_wait(bar_t3empty(st), ((uint32_t)(jc >> 1) & 1u) ^ 1u));
            tc_fence_after();
            const uint64_t bdesc = make_sdesc(base + P.off_wdg + (uint32_t)(ch * P.chunkN * 128));
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tmem_d + (uint32_t)st * 128u, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), P.idesc_dg, k4 > 0 ? 1u : 0u);
            umma_commit(bar_t3full(st));
          }
          umma_commit(bar_a2empty(buf));
        }
      }
      if (P.debug && blockIdx.x == 0) printf("last_fused CTA0 dgrad-MMA: wait operand tile %u, wait accumulator %u\n", dbg0, dbg1);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 16) {
    // ===================== S groups A (warps 4-7) and B (16-19): col2im accumulation (P1), likelihood gradient (P2) =====
    // P1: with two pre-activation images (P.dual) A takes the even tiles and TMEM accumulator 0, B the odd tiles and
    // accumulator 1, each adding into its own image -- no ordering between the groups is needed, and the sum of the two images
    // is formed in P2.  Without room for a second image A takes every tile and B only helps in P2.
    // P2: both groups, 256 threads over the block's output pixels.
    const bool is_s = warp < 8;
    const int gidx = is_s ? 0 : 1;
    float* const out_g = out_s + (P.dual ? gidx * P.out_floats : 0);
    const int pass_bar = 1 + 2 * gidx;   // named barrier of this group's passes (1 | 3); 2 = both groups
    const int q = warp & 3, sid = is_s ? threadIdx.x - 128 : threadIdx.x - 512 + 128;
    constexpr int NSC = K * K * NC;                 // live scatter columns
    constexpr int NLD = (NSC + 15) / 16 * 16;       // = Np_sc
    const uint32_t t_lane = tmem_y + ((uint32_t)(q * 32) << 16);
    const float gmul = P.inv_sigma2 * P.gscale;
    int it = 0, nb = 0;
    float loss_acc = 0.f;
    uint32_t dbg0 = 0, dbg1 = 0, dbg2 = 0, dbg3 = 0;
    uint32_t nx = 0;   // x chunks consumed so far (parity of bar_xfull)
    const float* x_s = reinterpret_cast<const float*>(gen_base + P.off_x);
    // rows [need_lo + lr0, +xr) of the block, clipped to the image: one contiguous run per channel
    auto issue_x = [&](const LastBlock& L, int lr0) {
      const int oy0 = max(0, L.need_lo + lr0), oy1 = min(P.Ho, L.need_lo + min(L.n_need, lr0 + P.xr));
      const uint32_t bytes = (uint32_t)((oy1 - oy0) * P.Wo * 4);
      mbar_expect_tx(bar_xfull, bytes * NC);
#pragma unroll
      for (int c = 0; c < NC; ++c)
        bulk_load_1d(base + P.off_x + (uint32_t)(c * P.xr * P.Wo * 4), P.x + (((size_t)L.b * NC + c) * P.Ho + oy0) * P.Wo, bytes, bar_xfull);
    };
    for (int blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x, ++nb) {
      const LastBlock L = last_block<K, S>(P, blk);
      if (sid == 0) issue_x(L, 0);   // this block's x rows (first chunk) travel while the scatter GEMM runs
      for (int t = L.pt0; t < L.pt1; ++t, ++it) {
        if (P.dual ? ((it & 1) != gidx) : !is_s) continue;
        const int as = it & 1;
        LF_T(dbg0, mbar_wait(bar_yfull(as), (uint32_t)(it >> 1) & 1u));
        tc_fence_after();
        const uint32_t tp1 = P.debug ? (uint32_t)clock() : 0u;
        uint32_t v[NLD];
        if constexpr (NLD == 16) {
          tmem_ld16(t_lane + (uint32_t)as * 64u, v);
        } else if constexpr (NLD == 32) {
          tmem_ld32(t_lane + (uint32_t)as * 64u, v);
        } else {
          static_assert(NLD == 48, "scatter width");
          tmem_ld32(t_lane + (uint32_t)as * 64u, v);
          tmem_ld16(t_lane + (uint32_t)as * 64u + 32u, v + 32);
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_yempty(as));   // the accumulator is in registers: the next tile may overwrite it
        const int r = q * 32 + lane;
        const bool valid = r < P.tile_rows;
        const int ry = r / P.Wi;
        const int iy = t * P.Ht + ry, ix = r - ry * P.Wi;
        // col2im: every thread adds its pixel's K*K*NC products into the pre-activation image.  Two threads may not add to
        // the same pixel concurrently, and the order of the additions must be fixed (bit-reproducible results), so the
        // slots are visited in passes whose targets are distinct across the group, separated by barriers.
        auto tgt = [&](int kh, int kw, bool& ok) -> float* {
          const int oy = iy * S - P.pad + kh, ox = ix * S - P.pad + kw, lrow = oy - L.need_lo;
          ok = valid && oy >= 0 && oy < P.Ho && lrow >= 0 && lrow < L.n_need && ox >= 0 && ox < P.Wo;
          return out_g + (lrow * P.Wo + ox) * NC;
        };
        if constexpr (S == 2) {
          // stride 2: the 2 x 2 slots {kh0, kh0+1} x {kw0, kw0+1} of ALL threads land on distinct pixels (distinct output
          // parities) -- one pass of 4 slots, loads batched ahead of the stores.  Passes that differ only in kw0 collide
          // only between horizontal neighbours: lanes of one warp when an image row does not straddle warps.
#pragma unroll
          for (int kh0 = 0; kh0 < K; kh0 += 2)
#pragma unroll
            for (int kw0 = 0; kw0 < K; kw0 += 2) {
              float* tp[4];
              bool ok[4];
              float tv[4][NC];
#pragma unroll
              for (int j = 0; j < 4; ++j) tp[j] = tgt(kh0 + (j >> 1), kw0 + (j & 1), ok[j]);
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < NC; ++c) tv[j][c] = ok[j] ? tp[j][c] : 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (ok[j]) {
#pragma unroll
                  for (int c = 0; c < NC; ++c)
                    tp[j][c] = tv[j][c] + __uint_as_float(v[((kh0 + (j >> 1)) * K + kw0 + (j & 1)) * NC + c]);
                }
              if (P.row_in_warp && kw0 + 2 < K) __syncwarp(); else named_bar_sync(pass_bar, 128);
            }
        } else {
          if (P.row_in_warp) {
            // stride 1, rows inside a warp: the three kw terms of an output pixel come from lanes ix+1, ix, ix-1 -- summed
            // through shuffles, then ONE add per kh (distinct pixels across the group)
            const bool has_r = ix + 1 < P.Wi, has_l = ix > 0;
#pragma unroll
            for (int kh = 0; kh < K; ++kh) {
              float tsum[NC];
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                const float fr = __shfl_down_sync(0xffffffffu, __uint_as_float(v[(kh * K + 0) * NC + c]), 1);
                const float fl = __shfl_up_sync(0xffffffffu, __uint_as_float(v[(kh * K + 2) * NC + c]), 1);
                tsum[c] = ((has_r ? fr : 0.f) + __uint_as_float(v[(kh * K + 1) * NC + c])) + (has_l ? fl : 0.f);
              }
              bool ok;
              float* tp = tgt(kh, 1, ok);
              if (ok) {
#pragma unroll
                for (int c = 0; c < NC; ++c) tp[c] += tsum[c];
              }
              named_bar_sync(pass_bar, 128);
            }
          } else {
#pragma unroll
            for (int kh = 0; kh < K; ++kh)
#pragma unroll
              for (int kw = 0; kw < K; ++kw) {
                bool ok;
                float* tp = tgt(kh, kw, ok);
                if (ok) {
#pragma unroll
                  for (int c = 0; c < NC; ++c) tp[c] += __uint_as_float(v[(kh * K + kw) * NC + c]);
                }
                named_bar_sync(pass_bar, 128);
              }
          }
        }
        if (P.debug) dbg1 += (uint32_t)clock() - tp1;
      }
      // ---- P2: x_hat, loss, likelihood gradient -------------------------------------------------------------------
      // x arrives in smem by bulk copies issued at the START of the block (a dependent global load per pixel would expose a
      // loaded-DRAM round trip per iteration, ~3 k clocks each); blocks whose rows do not fit take it in several chunks.
      const int own0 = L.r0 * S, own1 = L.r1 * S;
      named_bar_sync(2, 256);   // the pre-activation image is complete (S) -- the helpers may start (H)
      if (P.do_dgrad) LF_T(dbg2, mbar_wait(bar_gempty, ((uint32_t)nb & 1u) ^ 1u));   // the E group has gathered the previous block's gradient image
      float sq_acc = 0.f;
      const uint32_t tp2 = P.debug ? (uint32_t)clock() : 0u;
      for (int lr0 = 0; lr0 < L.n_need; lr0 += P.xr) {
        const int lr1 = min(L.n_need, lr0 + P.xr);
        if (lr0 > 0) {   // later chunk: every thread is done with the buffer, then one thread refills it
          named_bar_sync(2, 256);
          if (sid == 0) issue_x(L, lr0);
        }
        mbar_wait(bar_xfull, nx & 1u);
        ++nx;
        const int cy0 = max(0, L.need_lo + lr0);   // first image row held by the buffer
        // (a 4-pixel batched form of this loop measured 3-8 % slower on the 3-channel shapes, `#pragma unroll 2` 0-3 % slower)
        for (int p = lr0 * P.Wo + sid; p < lr1 * P.Wo; p += 256) {
          const int lrow = P.wo_shift >= 0 ? (p >> P.wo_shift) : p / P.Wo;
          const int ox = p - lrow * P.Wo, oy = L.need_lo + lrow;
          float g[4] = {0.f, 0.f, 0.f, 0.f};
          if (oy >= 0 && oy < P.Ho) {
            float* op = out_s + p * NC;
            const bool own = oy >= own0 && oy < own1;
            float h[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { h[c] = op[c]; op[c] = bias_r[c]; }
            if (P.dual) {
#pragma unroll
              for (int c = 0; c < NC; ++c) { h[c] += op[P.out_floats + c]; op[P.out_floats + c] = 0.f; }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
              // tanh(h) = 1 - 2 / (e^{2h} + 1): absolute error ~1e-7 (x_hat enters only through x_hat - x and 1 - x_hat^2)
              const float xh = 1.f - __fdividef(2.f, __expf(2.f * h[c]) + 1.f);
              if (P.xhat != nullptr && own) P.xhat[(((size_t)L.b * NC + c) * P.Ho + oy) * P.Wo + ox] = xh;
              const float rr = xh - x_s[(c * P.xr + (oy - cy0)) * P.Wo + ox];
              g[c] = rr * gmul * (1.f - xh * xh);
              if (own) { loss_acc += 0.5f * P.inv_sigma2 * rr * rr; sq_acc = fmaf(rr, rr, sq_acc); }
            }
          }
          if (P.do_dgrad) {
            const uint32_t w0 = pack2(P.op_fp16, g[0], g[1]), w1 = pack2(P.op_fp16, g[2], g[3]);
            *reinterpret_cast<uint2*>(g_s + (size_t)lrow * gpitch + (size_t)(ox + 1) * 8) = make_uint2(w0, w1);
          }
        }
      }
      if (P.debug) dbg3 += (uint32_t)clock() - tp2;
      __syncwarp();
      if (P.do_dgrad) {
        if (lane == 0) mbar_arrive(bar_gfull);   // release: this warp's gradient pixels are visible to the E group
      } else if (P.sq_part != nullptr) {
        // fixed-order reduction (lanes by shuffle tree, then the 8 warps in order): the per-chain squared error is bit-reproducible
        sq_acc = warp_sum(sq_acc);
        float* red = reinterpret_cast<float*>(gen_base + P.off_red);
        if (lane == 0) red[(is_s ? 0 : 4) + q] = sq_acc;
        named_bar_sync(2, 256);
        if (sid == 0) P.sq_part[blk] = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7])));
      }
      named_bar_sync(2, 256);                  // every pre-activation is re-initialised before the next block accumulates
    }
    if (P.debug && blockIdx.x == 0 && sid == 0)
      printf("last_fused CTA0 S group: wait scatter acc %u, col2im %u, wait gempty %u, P2 %u, blocks %d\n", dbg0, dbg1, dbg2, dbg3, nb);
    if (P.loss != nullptr) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && loss_acc != 0.f) atomicAdd(P.loss, loss_acc);
    }
  } else if (warp >= 8 && warp < 16 && P.do_dgrad) {
    // ===================== E group: operand gather (P3a) and the masked dgrad epilogue (P3b) =====================
    const int q = warp & 3, half = (warp - 8) >> 2, eid = threadIdx.x - 256;
    int jb = 0, jc = 0, nb = 0;
    uint32_t dbg0 = 0, dbg1 = 0, dbg2 = 0, dbg3 = 0, dbg4 = 0, dbg5 = 0;
    const uint32_t tstart_e = (uint32_t)clock();
    const int cw = P.chunkN >> 1;   // columns of a chunk handled by this warp
    const uint32_t t_lane = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * cw);
    uint32_t mnext[LF_MAX_CHUNKS][2];
    // sign-bit words of this lane's row of tile t: cw columns of every chunk (1 bit per element of a, NHWC order)
    auto load_masks = [&](int b, int t, uint32_t (&mw)[LF_MAX_CHUNKS][2]) {
      const int r = q * 32 + lane;
      const int ry = r / P.Wi;
      const long long m = r < P.tile_rows ? ((long long)b * P.Hi + t * P.Ht + ry) * P.Wi + (r - ry * P.Wi) : 0ll;
      const uint32_t* mrow = P.maskbits + ((m * P.C) >> 5);
#pragma unroll
      for (int ch = 0; ch < LF_MAX_CHUNKS; ++ch) {
        mw[ch][0] = mw[ch][1] = 0u;
        if (ch < P.nchunks) {
          const int n0 = ch * P.chunkN + half * cw;
          mw[ch][0] = __ldg(mrow + (n0 >> 5));
          if (cw > 32) mw[ch][1] = __ldg(mrow + (n0 >> 5) + 1);
        }
      }
    };
    for (int blk = blockIdx.x; blk < P.nblocks; blk += gridDim.x, ++nb) {
      const LastBlock L = last_block<K, S>(P, blk);
      load_masks(L.b, L.t0, mnext);   // the block's first tile: in flight while the S group finishes the gradient image
      LF_T(dbg0, mbar_wait(bar_gfull, (uint32_t)nb & 1u));
      auto build = [&](int t) {
        const int buf = jb & 1;
        LF_T(dbg1, mbar_wait(bar_a2empty(buf), ((uint32_t)(jb >> 1) & 1u) ^ 1u));
        const uint32_t tb = P.debug ? (uint32_t)clock() : 0u;
        const int r = eid & 127, hf = eid >> 7;   // warp-uniform half: no divergence between the two compile-time slot lists
        if (r < P.tile_rows) {
          const int ry = r / P.Wi;
          const int iy = t * P.Ht + ry, ix = r - ry * P.Wi;
          const uint32_t row = base + P.off_a2 + (uint32_t)buf * LF_TILE_BYTES + (uint32_t)r * 128u;
          const uint8_t* gpix = g_s + (size_t)(iy * S - P.pad - L.need_lo) * gpitch + (size_t)(ix * S - P.pad + 1) * 8;   // slot (0,0)
          auto gather_half = [&](auto HF) {   // slots [8 HF, 8 HF + 8): (kh, kw) are compile-time constants
            constexpr int hfc = decltype(HF)::value;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = hfc * 4 + jj;
              uint2 s01[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int slot = 2 * j + u;
                s01[u] = make_uint2(0u, 0u);
                if (slot < K * K) s01[u] = *reinterpret_cast<const uint2*>(gpix + (size_t)(slot / K) * gpitch + (size_t)(slot % K) * 8);
              }
              const uint32_t dst = row + (uint32_t)((j ^ (r & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(s01[0].x), "r"(s01[0].y), "r"(s01[1].x), "r"(s01[1].y) : "memory");
            }
          };
          if (hf == 0) gather_half(std::integral_constant<int, 0>{}); else gather_half(std::integral_constant<int, 1>{});
        }
        fence_proxy_async();   // generic-proxy writes of the operand tile -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a2full(buf));
        ++jb;
        if (P.debug) dbg2 += (uint32_t)clock() - tb;
      };
      auto done_gather = [&]() {   // this warp has read everything it needs from the block's gradient image
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_gempty);
      };
      build(L.t0);
      if (L.t0 + 1 == L.t1) done_gather();
      for (int t = L.t0; t < L.t1; ++t) {
        if (t + 1 < L.t1) {
          build(t + 1);
          if (t + 2 == L.t1) done_gather();
        }
        const uint32_t tq = P.debug ? (uint32_t)clock() : 0u;
        const int r = q * 32 + lane;
        const bool valid = r < P.tile_rows;
        const int ry = r / P.Wi;
        const int iy = t * P.Ht + ry, ix = r - ry * P.Wi;
        const long long m = valid ? ((long long)L.b * P.Hi + iy) * P.Wi + ix : 0ll;
        long long o;
        if (P.planar_out) {
          const int Hh = P.Hi >> 1, Wh = P.Wi >> 1;
          o = (long long)((iy & 1) * 2 + (ix & 1)) * P.B * Hh * Wh * P.C + (((long long)L.b * Hh + (iy >> 1)) * Wh + (ix >> 1)) * P.C;
        } else {
          o = m * P.C;
        }
        const unsigned long long out_row = (unsigned long long)P.gout + 2ull * (unsigned long long)(valid ? o : 0ll);
        // this tile's sign bits were requested one tile ago; request the next tile's now (a load issued where it is used
        // would expose a loaded-memory round trip per chunk)
        uint32_t mcur[LF_MAX_CHUNKS][2];
#pragma unroll
        for (int ch = 0; ch < LF_MAX_CHUNKS; ++ch) { mcur[ch][0] = mnext[ch][0]; mcur[ch][1] = mnext[ch][1]; }
        if (t + 1 < L.t1) load_masks(L.b, t + 1, mnext);
        if (P.debug) dbg5 += (uint32_t)clock() - tq;
#pragma unroll
        for (int ch = 0; ch < LF_MAX_CHUNKS; ++ch) {
          if (ch < P.nchunks) {
            const int st = jc & 1;
            const int n0 = ch * P.chunkN + half * cw;
            LF_T(dbg3, mbar_wait(bar_t3full(st), (uint32_t)(jc >> 1) & 1u));
            tc_fence_after();
            const uint32_t te = P.debug ? (uint32_t)clock() : 0u;
#pragma unroll
            for (int i32 = 0; i32 < 2; ++i32) {
              const int c32 = 32 * i32;
              if (c32 < cw) {
                uint32_t v[32];
                tmem_ld32(t_lane + (uint32_t)st * 128u + (uint32_t)c32, v);
                tmem_ld_wait();
                const uint32_t mword = mcur[ch][i32];
#pragma unroll
                for (int h16 = 0; h16 < 2; ++h16) {
                  uint32_t w[8];
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const int c0 = 16 * h16 + 2 * e;
                    const float s0 = (mword & (1u << c0)) ? 1.f : P.slope, s1 = (mword & (1u << (c0 + 1))) ? 1.f : P.slope;
                    w[e] = pack2(P.op_fp16, __uint_as_float(v[c0]) * s0, __uint_as_float(v[c0 + 1]) * s1);
                  }
                  if (valid) {
                    const unsigned long long a = out_row + 2ull * (unsigned long long)(n0 + c32 + 16 * h16);
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                  }
                }
              }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_t3empty(st));
            if (P.debug) dbg4 += (uint32_t)clock() - te;
            ++jc;
          }
        }
      }
    }
    if (P.debug && blockIdx.x == 0 && eid == 0)
      printf("last_fused CTA0 E group: total %u, wait gfull %u, wait operand buffer %u, gather %u, tile setup + mask words %u, wait dgrad acc %u, epilogue %u, chunks %d\n",
             (uint32_t)clock() - tstart_e, dbg0, dbg1, dbg2, dbg5, dbg3, dbg4, jc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
struct LastPlan {
  LastParams P;
  size_t smem;
  int variant;   // 0: k3-s1 nc3, 1: k3-s1 nc1, 2: k4-s2 nc3
};

static bool last_plan(const GenPack* g, int B, LastPlan* out) {
  const GenLayer& y = g->layers[g->nlayers - 1];
  const bool same = y.k == 3 && y.stride == 1 && y.pad == 1, up = y.k == 4 && y.stride == 2 && y.pad == 1;
  int variant = -1;
  if (same && y.cout == 3) variant = 0;
  else if (same && y.cout == 1) variant = 1;
  else if (up && y.cout == 3) variant = 2;
  if (variant < 0 || !is_tc_precision(g->precision) || !g->use_tc || !g->use_bits || !g->last_scatter) return false;
  if (y.cin % 64 || y.Win > 128 || y.Hin * y.Win < 128) return false;
  LastParams P{};
  P.B = B; P.C = y.cin; P.Hi = y.Hin; P.Wi = y.Win; P.Ho = y.Hout; P.Wo = y.Wout; P.pad = y.pad;
  P.Np_sc = y.np_sc;
  P.Ht = 128 / y.Win;
  if (y.Hin % P.Ht) return false;
  P.tile_rows = P.Ht * y.Win;
  P.tpi = y.Hin / P.Ht;
  P.kb_sc = y.cin / 64;
  P.chunkN = y.cin % 128 == 0 ? 128 : 64;
  P.nchunks = y.cin / P.chunkN;
  if (P.Np_sc > 64 || P.Np_sc != (y.k * y.k * y.cout + 15) / 16 * 16) return false;
  const int halo_rows = same ? 2 : 1;
  const size_t wbytes = (size_t)P.kb_sc * P.Np_sc * 128 + (size_t)y.cin * 128;
  const size_t fixed = 2 * LF_TILE_BYTES + wbytes + 8 * (2 * LF_MAX_STAGES + 16) + 112 + 1024;
  const size_t cap = 227 * 1024;
  auto need_rows = [&](int Rt) { return (Rt * P.Ht - 1) * y.stride - y.pad + y.k - 1 - (0 * y.stride - y.pad) + 1; };
  if (P.nchunks > LF_MAX_CHUNKS) return false;
  int best_rt = 0, best_stages = 0, best_xr = 0;
  bool best_dual = false;
  for (int Rt = P.tpi; Rt >= 1 && !best_rt; --Rt) {
    if (Rt < P.tpi && Rt > 16) continue;
    const int nr = need_rows(Rt);
    for (int xchunks = 1; xchunks <= 3 && !best_rt; ++xchunks) {   // x rows staged per bulk copy: the whole block, or 1/2, 1/3 of it
      const int xr = ceil_div(nr, xchunks);
      const size_t outb = align_up((size_t)nr * y.Wout * y.cout * 4, 16);
      size_t img = outb + align_up((size_t)nr * (y.Wout + 2) * 8, 16) + align_up((size_t)y.cout * xr * y.Wout * 4, 16);
      if (fixed + img + 4 * LF_TILE_BYTES > cap) continue;   // at least a 4-stage activation ring
      best_dual = fixed + img + outb + 4 * LF_TILE_BYTES <= cap;   // room for the second pre-activation image
      if (best_dual) img += outb;
      best_rt = Rt;
      best_xr = xr;
      best_stages = (int)std::min<size_t>(LF_MAX_STAGES, (cap - fixed - img) / LF_TILE_BYTES);
    }
  }
  if (!best_rt) return false;
  if ((y.Wout * 4) % 16) return false;   // bulk copies move whole 16-byte units
  P.Rt = best_rt;
  P.stages = best_stages;
  P.xr = best_xr;
  P.row_in_warp = (32 % y.Win == 0) ? 1 : 0;
  P.halo_t = P.Rt == P.tpi ? 0 : ceil_div(halo_rows, P.Ht);
  P.nblk_img = ceil_div(P.tpi, P.Rt);
  P.nblocks = B * P.nblk_img;
  const int nr = need_rows(P.Rt);
  P.out_floats = nr * y.Wout * y.cout;
  P.wo_shift = -1;
  for (int l2 = 0; l2 < 12; ++l2) if ((1 << l2) == y.Wout) P.wo_shift = l2;
  P.g_bytes = (int)align_up((size_t)nr * (y.Wout + 2) * 8, 16);
  uint32_t off = (uint32_t)P.stages * LF_TILE_BYTES;
  P.off_a2 = off; off += 2 * LF_TILE_BYTES;
  P.off_wsc = off; off += (uint32_t)(P.kb_sc * P.Np_sc * 128);   // multiples of 2 KB: every k-block slab stays 1024-aligned
  off = (uint32_t)align_up(off, 1024);
  P.off_wdg = off; off += (uint32_t)(y.cin * 128);
  P.dual = best_dual ? 1 : 0;
  P.off_out = off; off += (uint32_t)align_up((size_t)P.out_floats * 4 * (P.dual ? 2 : 1), 16);
  P.off_g = off; off += (uint32_t)P.g_bytes;
  P.off_x = off; off += (uint32_t)align_up((size_t)y.cout * P.xr * y.Wout * 4, 16);
  P.off_red = (uint32_t)align_up(off, 16); off = P.off_red + 32;
  P.off_bar = (uint32_t)align_up(off, 8);
  const size_t total = P.off_bar + 8 * (2 * LF_MAX_STAGES + 16) + 16 + 1024;
  if (total > cap) return false;
  const uint32_t opfmt = g->precision == DAMC_PREC_FP16 ? 0u : 1u;
  P.op_fp16 = g->precision == DAMC_PREC_FP16 ? 1 : 0;
  P.idesc_sc = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(P.Np_sc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  P.idesc_dg = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(P.chunkN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  out->P = P;
  out->smem = total;
  out->variant = variant;
  return true;
}

bool last_fused_supported(const GenPack* g) {
  const char* e = getenv("DAMC_LAST_FUSED");   // read per handle: tests pack one generator with and one without the fusion
  LastPlan pl;
  return !(e && e[0] == '0') && last_plan(g, 1, &pl);
}

int last_fused_parts(const GenPack* g) {
  LastPlan pl;
  return last_plan(g, 1, &pl) ? pl.P.nblk_img : 1;
}

int launch_last_fused(const GenPack* g, const GenWorkspace& ws, int B, const float* x, float sigma, float* xhat, float* loss,
                      float* sq_part, cudaStream_t stream) {
  LastPlan pl;
  if (!last_plan(g, B, &pl)) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "fused last layer: unsupported shape");
  const int L = g->nlayers;
  const GenLayer& y = g->layers[L - 1];
  LastParams& P = pl.P;
  P.x = x; P.xhat = xhat; P.loss = loss; P.bias = y.bias;
  P.inv_sigma2 = 1.0f / (sigma * sigma);
  P.gscale = generator_grad_scale(g, sigma);
  P.slope = g->slope;
  P.maskbits = ws.mask[L - 2];
  P.gout = ws.grad[L - 2];
  P.planar_out = g->layers[L - 2].type == L_UP ? 1 : 0;
  P.do_dgrad = sq_part == nullptr ? 1 : 0;
  P.sq_part = sq_part;
  { const char* e = getenv("DAMC_LAST_DEBUG"); P.debug = (e && e[0] == '1') ? 1 : 0; }
  CUtensorMap tmA, tmWsc, tmWdg;
  DAMC_TRY(tc_encode_act(&tmA, g->precision, ws.act[L - 2], y.cin, y.Win, y.Hin, B, P.Ht));
  DAMC_TRY(tc_encode_2d(&tmWsc, P.op_fp16, y.w_scatter_tc, y.cin, P.Np_sc, P.Np_sc));
  DAMC_TRY(tc_encode_2d(&tmWdg, P.op_fp16, y.w_dgrad_tc, 64, y.cin, P.chunkN));
  constexpr int kMaxDev = 64;
  static std::atomic<int> sms_of[kMaxDev];
  int dev = 0;
  DAMC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "fused last layer: device ordinal %d out of range", dev);
  int num_sms = sms_of[dev].load(std::memory_order_acquire);
  if (!num_sms) {
    const int big = 227 * 1024;
    DAMC_CUDA(cudaFuncSetAttribute(last_fused_kernel<3, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(last_fused_kernel<3, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(last_fused_kernel<4, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    sms_of[dev].store(num_sms, std::memory_order_release);
  }
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(std::min(P.nblocks, num_sms));
  cfg.blockDim = dim3(LF_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pl.variant == 0) DAMC_CUDA(cudaLaunchKernelEx(&cfg, last_fused_kernel<3, 1, 3>, tmA, tmWsc, tmWdg, pl.P));
  else if (pl.variant == 1) DAMC_CUDA(cudaLaunchKernelEx(&cfg, last_fused_kernel<3, 1, 1>, tmA, tmWsc, tmWdg, pl.P));
  else DAMC_CUDA(cudaLaunchKernelEx(&cfg, last_fused_kernel<4, 2, 3>, tmA, tmWsc, tmWdg, pl.P));
  return DAMC_OK;
}

}  // namespace damc
