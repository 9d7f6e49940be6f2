// gen_tc.cu -- tcgen05 / TMEM / TMA engine for the generator's shifted-window GEMMs (DAMC_PREC_BF16).
//
// Replaces the cuDNN conv-transpose forward and dgrad kernels PyTorch dispatches for netG(z) and autograd.grad through
// netG (reference workspace/src/MCMC.py:55-60, diffusion_net.py:26-45).  Same formulation as gen_simt.cu -- every layer
// pass is  D[m,n] = sum_taps A_tap[m,:] . W_tap[n,:]  with A_tap a shifted window of an NHWC (or parity-planar) tensor --
// but the operands are bf16, fetched by TMA and multiplied by the 5th-generation tensor cores into TMEM:
//
//   * A tile  : one TMA 5-D box {64 ch, W, Ht, Bt, plane} of the source tensor at a tap-shifted coordinate.  Rows of
//               the box are pixels, each 64 bf16 = one 128-byte swizzle row, so the box lands directly in the UMMA
//               K-major SWIZZLE_128B layout; out-of-image taps are zero-filled by the TMA unit (no padding, no im2col).
//   * B tile  : TMA 2-D box {64, BN} of the pre-packed weights [tap][n][c] (K-major as well).
//   * MMA     : tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BN <= 256, K = 16 x 4 per 64-channel block, fp32
//               accumulators in TMEM (2 x 256 columns: the epilogue of tile i overlaps the main loop of tile i+1).
//   * roles   : warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator, warps 2..9 = epilogue (tcgen05.ld, fused
//               bias / LeakyReLU / mask / tanh-likelihood, 16-byte stores).  Persistent CTAs, one per SM, static tile
//               order with the N tiles of one M tile adjacent (A re-reads hit L2).
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <atomic>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "gen_epilogue.cuh"
#include "tc_ptx.cuh"

namespace damc {

constexpr int TC_BM = 128, TC_BK = 64, TC_MAX_STAGES = 8, TC_TMEM_COLS = 512;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;  // 16 KB
constexpr int TC_STAGING_PER_WARP = 32 * 64 * 2;  // per epilogue warp: 32 rows x 64 bf16

// n / d for 0 <= n < 2^31 without the ~25-instruction integer division: mul = ceil(2^(31+s) / d), s = ceil(log2 d)
struct FastDiv {
  uint32_t mul, shr;  // d == 1: mul = 0
  int d;
  __device__ __forceinline__ int div(int n) const {
    return mul ? (int)((uint32_t)(((unsigned long long)(uint32_t)n * mul) >> 31) >> shr) : n;
  }
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const { q = div(n); r = n - q * d; }
};
static FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  if (d <= 1) { f.mul = 0; f.shr = 0; f.d = 1; return f; }
  uint32_t sh = 0;
  while ((1u << sh) < (uint32_t)d) ++sh;
  f.shr = sh;
  f.mul = (uint32_t)((((unsigned long long)1 << (31 + sh)) + (unsigned long long)d - 1) / (unsigned long long)d);
  return f;
}

struct TcParams {
  GemmPlan plan;
  FastDiv fd_ntiles, fd_munits, fd_ncls, fd_tpi, fd_perimg, fd_wm, fd_bias;
  const float* bias_src;   // bias vector copied into smem at kernel start (bias_floats > 0), at staging + bias_off bytes
  int bias_floats, bias_off;
  int BN, stages, b_stage_bytes;
  int bk;           // channels per k-block = one 128-byte swizzle row: 64 (16-bit operands) or 32 (tf32: fp32 containers)
  int tf32;         // kind::tf32 MMAs on fp32 tensors (DAMC_PREC_TF32); epilogues store tf32-rounded fp32
  int m_tiles, n_tiles, kb_per_tap, kb_total, kb_per_split;
  int Ht, Bt, tiles_per_img, tile_rows;
  uint32_t a_box_bytes, b_box_bytes, idesc;
  int wm_shift, per_img_shift;  // log2(Wm), log2(Ht*Wm) when powers of two, else -1 (row decode without divisions)
  int cls_inner;    // merged parity classes: class index is the fastest tile dimension
  int den_kb_h;     // denoiser launches: k-blocks [0, den_kb_h) carry the layer input h, the rest the ctx activation c
  int direct;       // staged-epilogue launches: lanes store their own row pieces with 256-bit stores, no smem staging tiles
  int dbg;          // experiment switches (env DAMC_TC_DBG, profiles/r01_epilogue_ablation.txt): 1 no global stores, 2 no epilogue
                    // math, 4 smem-staged 16-byte stores instead of direct 256-bit ones, 32 / 64 / 128 stores confined to 1 / 32 / 256 MB
  int stage_cols;   // 0: per-thread row stores; 64 | 128: epilogue staged through smem for coalesced 16-byte rows
};

// ---- epilogue for one row x 16 consecutive columns ---------------------------------------------------------------------
struct RowCtx {
  bool ok;
  int m, b, y, x;
};

template <bool F32>
__device__ __forceinline__ void epi_chunk16(const GemmPlan& p, const RowCtx& r, int split, int n0, const uint32_t raw[16],
                                            float& loss_acc) {
  const Epilogue& e = p.epi;
  if (!r.ok || n0 >= p.Np) return;
  if (e.kind == EPI_FWD_ACT && n0 + 16 <= p.N && !F32) {
    const long long o = (long long)r.b * e.o_b + (long long)(r.y * e.sy + e.py) * e.o_y +
                        (long long)(r.x * e.sx + e.px) * e.o_x + n0;
    const float4* bp = reinterpret_cast<const float4*>(e.bias + (n0 % e.bias_mod));
    uint32_t w[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bb = __ldg(bp + q);
      float h0 = __uint_as_float(raw[4 * q + 0]) + bb.x, h1 = __uint_as_float(raw[4 * q + 1]) + bb.y;
      float h2 = __uint_as_float(raw[4 * q + 2]) + bb.z, h3 = __uint_as_float(raw[4 * q + 3]) + bb.w;
      h0 = h0 > 0.f ? h0 : e.slope * h0; h1 = h1 > 0.f ? h1 : e.slope * h1;
      h2 = h2 > 0.f ? h2 : e.slope * h2; h3 = h3 > 0.f ? h3 : e.slope * h3;
      w[2 * q] = pack2(p.op_fp16, h0, h1);
      w[2 * q + 1] = pack2(p.op_fp16, h2, h3);
    }
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + o);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    return;
  }
  if (e.kind == EPI_DGRAD_MASK && n0 + 16 <= p.N && !F32) {
    const uint4* ap = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.act) + (long long)r.m * p.N + n0);
    long long o;
    if (e.planar_out) {
      const int Hh = p.Hm >> 1, Wh = p.Wm >> 1;
      o = (long long)((r.y & 1) * 2 + (r.x & 1)) * p.B * Hh * Wh * p.N +
          (((long long)r.b * Hh + (r.y >> 1)) * Wh + (r.x >> 1)) * p.N + n0;
    } else {
      o = (long long)r.m * p.N + n0;
    }
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(e.out) + o);
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      const uint4 a = __ldg(ap + hlf);
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
      uint32_t w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        // bf16 sign/zero test on the raw bits: positive and non-zero  <=>  (bits & 0x7fff) != 0 && sign == 0
        const uint32_t lo = aw[q] & 0xffffu, hi = aw[q] >> 16;
        const float s0 = (lo != 0u && lo < 0x8000u) ? 1.f : e.slope;
        const float s1 = (hi != 0u && hi < 0x8000u) ? 1.f : e.slope;
        w[q] = pack2(p.op_fp16, __uint_as_float(raw[8 * hlf + 2 * q]) * s0, __uint_as_float(raw[8 * hlf + 2 * q + 1]) * s1);
      }
      dst[hlf] = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return;
  }
  if (e.kind == EPI_STORE_F32 && n0 + 16 <= p.Np) {  // padded weight columns are zero, so storing all Np columns is exact
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (long long)r.m * e.nz_out + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      dst[q] = make_float4(__uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]), __uint_as_float(raw[4 * q + 2]),
                           __uint_as_float(raw[4 * q + 3]));
    return;
  }
  if (e.kind == EPI_STORE_F32_BIAS && n0 + 16 <= p.N && e.bias != nullptr) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(e.out) + (long long)r.m * e.nz_out + n0);
    const float4* bp = reinterpret_cast<const float4*>(e.bias + n0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bb = __ldg(bp + q);
      dst[q] = make_float4(__uint_as_float(raw[4 * q]) + bb.x, __uint_as_float(raw[4 * q + 1]) + bb.y,
                           __uint_as_float(raw[4 * q + 2]) + bb.z, __uint_as_float(raw[4 * q + 3]) + bb.w);
    }
    return;
  }
#pragma unroll 1
  for (int j = 0; j < 16; ++j)
    if (n0 + j < p.N) {
      if (F32) epilogue_elem<tf32_t>(p, split, r.m, r.b, r.y, r.x, n0 + j, __uint_as_float(raw[j]), loss_acc);
      else if (p.op_fp16) epilogue_elem<__half>(p, split, r.m, r.b, r.y, r.x, n0 + j, __uint_as_float(raw[j]), loss_acc);
      else epilogue_elem<__nv_bfloat16>(p, split, r.m, r.b, r.y, r.x, n0 + j, __uint_as_float(raw[j]), loss_acc);
    }
}

// ---- staged epilogue: TMEM -> registers -> (bias+LeakyReLU | mask) -> swizzled smem -> coalesced 16-byte row stores ------
// A warp owns 32 accumulator rows (its TMEM lane quarter) and walks 64-column segments.  Row segments are moved between
// global memory and a per-warp XOR-swizzled smem tile so that 8 consecutive lanes cover one contiguous 128-byte piece of
// a row (coalesced both ways).  For the mask epilogue the activation rows (sign source) are fetched into registers TWO
// segments ahead -- across tile boundaries -- so that each lane keeps 16 independent 16-byte loads in flight.
struct StagedEpi {
  static constexpr int CPR = 8, SW = 64, RPI = 4, NIT = 8;
  static constexpr uint32_t row_bytes = SW * 2;
  unsigned long long out_bits;
  unsigned long long mrow_bits;  // this lane's row of mask words (1 bit per element), or 0
  int sub, j, dbg;
  int direct;  // 1: each lane stores its own row piece with 256-bit stores (no smem round trip); 0: staged through smem

  __device__ __forceinline__ static unsigned long long act_ptr_bits(const GemmPlan& p, const RowCtx& rc) {
    return (unsigned long long)(reinterpret_cast<const __nv_bfloat16*>(p.epi.act) + (rc.ok ? (long long)rc.m * p.N : 0ll));
  }
  template <bool F32>
  __device__ __forceinline__ void set_out(const GemmPlan& p, const RowCtx& rc, int cls) {
    const Epilogue& e = p.epi;
    long long o = 0;   // element offset of this lane's output row
    if (rc.ok) {
      if (e.kind == EPI_DGRAD_MASK) {
        if (e.planar_out) {
          const int Hh = p.Hm >> 1, Wh = p.Wm >> 1;
          o = (long long)((rc.y & 1) * 2 + (rc.x & 1)) * p.B * Hh * Wh * p.N +
              (((long long)rc.b * Hh + (rc.y >> 1)) * Wh + (rc.x >> 1)) * p.N;
        } else {
          o = (long long)rc.m * p.N;
        }
      } else {
        const int py = p.ncls > 1 ? (cls >> 1) : e.py, px = p.ncls > 1 ? (cls & 1) : e.px;
        o = (long long)rc.b * e.o_b + (long long)(rc.y * e.sy + py) * e.o_y + (long long)(rc.x * e.sx + px) * e.o_x;
      }
    }
    out_bits = rc.ok ? (unsigned long long)e.out + (unsigned long long)o * (F32 ? 4ull : 2ull) : 0ull;
    mrow_bits = 0ull;
    if (e.maskbits != nullptr && rc.ok) {
      const long long elem0 = e.kind == EPI_DGRAD_MASK ? (long long)rc.m * p.N : o;
      mrow_bits = (unsigned long long)(e.maskbits + (elem0 >> 5));
    }
  }
  // mask words of the 64-column segment at n_base for this lane's row (dgrad; rows outside the problem read word 0)
  __device__ __forceinline__ static uint2 load_mask_words(const GemmPlan& p, const RowCtx& rc, int n_base) {
    const uint32_t* w = p.epi.maskbits + (rc.ok ? (((long long)rc.m * p.N + n_base) >> 5) : 0ll);
    return __ldg(reinterpret_cast<const uint2*>(w));
  }
  // issue the 8 coalesced 16-byte loads of one 32-row x 64-column activation segment
  __device__ __forceinline__ void request(uint4 (&r)[NIT], unsigned long long act_bits, int n_base) const {
    // Unconditional loads: rows outside the problem read row 0 (their results are never stored).  A predicated load
    // followed by a select would make every load wait for its own result and serialise the batch.
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const unsigned long long pb = __shfl_sync(0xffffffffu, act_bits, i * RPI + sub);
      r[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(pb) + n_base) + j);
    }
  }
  __device__ __forceinline__ void stash(const uint4 (&r)[NIT], uint32_t my_stage) const {
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int R = i * RPI + sub;
      const uint32_t dst = my_stage + (uint32_t)R * row_bytes + (uint32_t)((j ^ (R & (CPR - 1))) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[i].x), "r"(r[i].y), "r"(r[i].z), "r"(r[i].w) : "memory");
    }
    __syncwarp();
  }
  // accumulators of one segment -> epilogue math -> smem -> global
  template <bool F32>
  __device__ __forceinline__ void process(const GemmPlan& p, uint32_t t_seg, uint32_t my_stage, int lane, int n_base,
                                          uint2 mw, const float* sbias /* smem bias + (n_base % bias_mod) */) const {
    const Epilogue& e = p.epi;
    const bool is_mask = e.kind == EPI_DGRAD_MASK;
    const bool bits = e.maskbits != nullptr;
    uint32_t wout[2] = {0u, 0u};
    if (dbg & 2) return;
#pragma unroll 1
    for (int c = 0; c < SW; c += 32) {
      uint32_t v[32];
      tmem_ld32(t_seg + (uint32_t)c, v);
      tmem_ld_wait();
      const uint32_t mword = c == 0 ? mw.x : mw.y;
      uint32_t oword = 0u;
      uint32_t wprev[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int jj = (c >> 3) + g;
        const uint32_t addr = my_stage + (uint32_t)lane * row_bytes + (uint32_t)((jj ^ (lane & (CPR - 1))) << 4);
        uint32_t w[4];
        if constexpr (F32) {
          // tf32 mode: 8 columns = 32 bytes of this lane's fp32 row, rounded to tf32 (the next GEMM's operand) -> one 256-bit store
          uint32_t f[8];
          if (is_mask) {
#pragma unroll
            for (int t2 = 0; t2 < 8; ++t2)
              f[t2] = __float_as_uint(tf32_rna(__uint_as_float(v[8 * g + t2]) * ((mword & (1u << (8 * g + t2))) ? 1.f : e.slope)));
          } else {
            const float4* bp = reinterpret_cast<const float4*>(sbias + c + 8 * g);
            const float4 b0v = bp[0], b1v = bp[1];
            const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
            for (int t2 = 0; t2 < 8; ++t2) {
              const float h0 = __uint_as_float(v[8 * g + t2]) + bb[t2];
              const bool p0 = h0 > 0.f;
              oword |= (p0 ? 1u : 0u) << (8 * g + t2);
              f[t2] = __float_as_uint(tf32_rna(p0 ? h0 : h0 * e.slope));
            }
          }
          if (out_bits && !(dbg & 1)) {
            const unsigned long long a = out_bits + 4ull * (unsigned long long)(n_base + c + 8 * g);
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a), "r"(f[0]), "r"(f[1]), "r"(f[2]),
                         "r"(f[3]), "r"(f[4]), "r"(f[5]), "r"(f[6]), "r"(f[7]) : "memory");
          }
          continue;
        }
        if (is_mask && bits) {
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            const float g0 = __uint_as_float(v[8 * g + 2 * t2]), g1 = __uint_as_float(v[8 * g + 2 * t2 + 1]);
            const float s0 = (mword & (1u << (8 * g + 2 * t2))) ? 1.f : e.slope;       // select, not branch
            const float s1 = (mword & (1u << (8 * g + 2 * t2 + 1))) ? 1.f : e.slope;
            w[t2] = pack2(p.op_fp16, g0 * s0, g1 * s1);
          }
        } else if (is_mask) {
          uint32_t aw[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(aw[0]), "=r"(aw[1]), "=r"(aw[2]), "=r"(aw[3]) : "r"(addr) : "memory");
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            // two bf16 activations per word: "> 0" is a signed compare of each half (as int32: high half >= 0x10000)
            float g0 = __uint_as_float(v[8 * g + 2 * t2]), g1 = __uint_as_float(v[8 * g + 2 * t2 + 1]);
            if ((int)(aw[t2] << 16) <= 0) g0 *= e.slope;
            if ((int)aw[t2] < 0x10000) g1 *= e.slope;
            w[t2] = pack2(p.op_fp16, g0, g1);
          }
        } else {
          const float4* bp = reinterpret_cast<const float4*>(sbias + c + 8 * g);
          const float4 b0v = bp[0], b1v = bp[1];
          const float bb[8] = {b0v.x, b0v.y, b0v.z, b0v.w, b1v.x, b1v.y, b1v.z, b1v.w};
#pragma unroll
          for (int t2 = 0; t2 < 4; ++t2) {
            const float h0 = __uint_as_float(v[8 * g + 2 * t2]) + bb[2 * t2], h1 = __uint_as_float(v[8 * g + 2 * t2 + 1]) + bb[2 * t2 + 1];
            const bool p0 = h0 > 0.f, p1 = h1 > 0.f;   // selects, not branches
            oword |= (p0 ? 1u : 0u) << (8 * g + 2 * t2) | (p1 ? 1u : 0u) << (8 * g + 2 * t2 + 1);
            w[t2] = pack2(p.op_fp16, p0 ? h0 : h0 * e.slope, p1 ? h1 : h1 * e.slope);
          }
        }
        if (direct) {
          // 8 columns = 16 bytes of this lane's own row; two g's make one 256-bit store (full 32-byte sectors)
          if (g & 1) {
            if (out_bits && !(dbg & 1)) {
              unsigned long long a = out_bits + 2ull * (unsigned long long)(n_base + c + 8 * (g - 1));
              if (dbg & (32 | 64 | 128)) {   // experiment: same store instructions, confined to the first 1 / 32 / 256 MB of the output
                const unsigned long long base = (unsigned long long)p.epi.out;
                const unsigned long long span = (dbg & 32) ? (1ull << 20) : (dbg & 64) ? (32ull << 20) : (256ull << 20);
                a = base + ((a - base) & (span - 32ull));
              }
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a), "r"(wprev[0]), "r"(wprev[1]),
                           "r"(wprev[2]), "r"(wprev[3]), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
            }
          } else {
            wprev[0] = w[0]; wprev[1] = w[1]; wprev[2] = w[2]; wprev[3] = w[3];
          }
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        }
      }
      wout[c >> 5] = oword;
    }
    if (!is_mask && bits && mrow_bits)  // sign bits of this row's 64 pre-activations (consumed by the dgrad epilogue)
      *reinterpret_cast<uint2*>(reinterpret_cast<uint32_t*>(mrow_bits) + (n_base >> 5)) = make_uint2(wout[0], wout[1]);
    if (direct) return;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NIT; ++i) {  // write-out: 8 consecutive lanes cover one contiguous 128-byte row piece
      const int R = i * RPI + sub;
      const unsigned long long pb = __shfl_sync(0xffffffffu, out_bits, R);
      const uint32_t src = my_stage + (uint32_t)R * row_bytes + (uint32_t)((j ^ (R & (CPR - 1))) << 4);
      uint4 o4;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o4.x), "=r"(o4.y), "=r"(o4.z), "=r"(o4.w) : "r"(src) : "memory");
      if (pb && !(dbg & 1)) *(reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(pb) + n_base) + j) = o4;
    }
    __syncwarp();
  }
};

// taps of output-parity class cls = py*2+px of a k4-s2-p1 transposed convolution (see gen_driver.cu: up_fwd_taps)
__device__ __forceinline__ Tap up_fwd_tap(int cls, int t) {
  const int py = cls >> 1, px = cls & 1, ty = t >> 1, tx = t & 1;
  Tap r;
  r.plane = 0;
  r.dy = (signed char)(py == 0 ? (ty == 0 ? 0 : -1) : (ty == 0 ? 1 : 0));
  r.dx = (signed char)(px == 0 ? (tx == 0 ? 0 : -1) : (tx == 0 ? 1 : 0));
  r.pad = 0;
  return r;
}

// ---- the kernel --------------------------------------------------------------------------------------------------------
// EW = epilogue warps: 8 for MMA-bound launches; 16 (four per TMEM lane quarter) when K is so short that the epilogue
// is the critical path and needs the extra issue slots.
// CG = CTAs per MMA: 1, or 2 (cta_group::2: a CTA pair computes a 256 x 256 tile; each CTA stages its own 128 rows of
// A and 128 of the 256 weight rows, so L2->smem traffic and B smem reads per FLOP drop by a third; the leader CTA's
// MMA thread issues for both, commits multicast to both CTAs' barriers).
// F32 = tf32 mode (fp32 containers, kind::tf32 MMAs, fp32 row stores); a template parameter so that the 16-bit
// instantiations keep their register allocation.
template <int EW, int CG, bool F32>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
convgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment of the tile ring is required by SWIZZLE_128B
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = TC_A_BYTES + P.b_stage_bytes;
  const uint32_t bars = smem_base + (uint32_t)P.stages * stage_bytes;  // [full x S][empty x S][tfull x 2][tempty x 2][tmem ptr]
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (P.stages + s); };
  auto bar_tfull = [&](int a) { return bars + 8u * (2 * P.stages + a); };
  auto bar_tempty = [&](int a) { return bars + 8u * (2 * P.stages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * P.stages + 4);
  const uint32_t staging = (tmem_slot + 16u + 127u) & ~127u;  // [4 warps][32 rows][<=256 B], XOR-swizzled 16-B chunks
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  // Programmatic dependent launch: let the next kernel of the stream begin its own prologue now; it (like this kernel,
  // below) blocks in griddepcontrol.wait until its predecessor has completed and flushed before touching any tensor.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // MMA unit (CTA or CTA pair)
  const int nunits = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int s = 0; s < P.stages; ++s) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull(a), 1); mbar_init(bar_tempty(a), EW * CG); }
    fence_barrier_init();
  }
  if (warp == 1) { if (CG == 2) tmem_alloc_2sm(tmem_slot, TC_TMEM_COLS); else tmem_alloc(tmem_slot, TC_TMEM_COLS); }
  const bool den_kind = P.plan.epi.kind == EPI_DEN_LAYER || P.plan.epi.kind == EPI_DEN_FINAL;
  // the launch's bias vector (forward: [Cout]; denoiser: bias quads of the whole layer) lives in smem behind the staging tiles
  float* den_bias = reinterpret_cast<float*>(smem_raw + (staging - smem_u32(smem_raw)) + P.bias_off);
  // The bias vector is written by the handle's refill kernels, which precede a non-PDL kernel of the same call (stage_z /
  // the hoist kernels), so it is complete before any PDL-launched kernel of the chain can start: safe to read early.
  for (int i = threadIdx.x; i < P.bias_floats; i += blockDim.x) den_bias[i] = __ldg(P.bias_src + i);
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before any remote arrive / TMA signal
  tc_fence_after();
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything above overlapped the previous kernel's tail
  const uint32_t tmem_base = *tmem_slot_ptr;

  const GemmPlan& p = P.plan;
  const int m_units = (P.m_tiles + CG - 1) / CG;
  const int ncls = p.ncls > 1 ? p.ncls : 1;
  const int total_tiles = m_units * P.n_tiles * p.ksplit * ncls;

  // tile -> (m tile, n tile, K split | parity class).  sp carries the split index, or the class when ncls > 1.
  auto decode = [&](int tile, int& mt, int& nt, int& sp) {
    if (ncls > 1 && P.cls_inner) {  // classes innermost: the 4 classes of one input window run back to back and share it through L2
      int q;
      P.fd_ncls.divmod(tile, q, sp);
      int mu;
      P.fd_ntiles.divmod(q, mu, nt);
      mt = mu * CG + (int)cta_rank;
      return;
    }
    int r;
    P.fd_ntiles.divmod(tile, r, nt);
    int mu;
    P.fd_munits.divmod(r, sp, mu);
    mt = mu * CG + (int)cta_rank;   // this CTA's 128-row tile (may lie past the end: rows are masked)
  };
  auto tile_origin = [&](int mt, int& b0, int& y0) {
    if (P.Bt == 1) { int rem; P.fd_tpi.divmod(mt, b0, rem); y0 = rem * P.Ht; }
    else { b0 = mt * P.Bt; y0 = 0; }
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // lane 0 issues the loads.  (Tried and dropped: idle lanes prefetching the next tile's A window into L2 with
    // cp.async.bulk.prefetch.tensor -- 3-7 % slower on the MMA-bound launches.)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; lane == 0 && tile < total_tiles; tile += nunits) {
      int mt, nt, sp, b0, y0;
      decode(tile, mt, nt, sp);
      {
        tile_origin(mt, b0, y0);
        const int cls = ncls > 1 ? sp : 0;
        const int kb0 = ncls > 1 ? 0 : sp * P.kb_per_split, kb1 = min(P.kb_total, kb0 + P.kb_per_split);
        const int wrow0 = cls * p.ntaps * p.Np;  // class block inside the weight tensor
        for (int kb = kb0; kb < kb1; ++kb) {
          const int t = kb / P.kb_per_tap, c0 = (kb - t * P.kb_per_tap) * P.bk;
          const Tap tp = ncls > 1 ? up_fwd_tap(cls, t) : p.taps[t];
          mbar_wait(bar_empty(stage), phase ^ 1u);
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          if (CG == 2) {
            // both CTAs' loads report to the LEADER's full barrier, which expects the bytes of the whole pair
            if (cta_rank == 0) mbar_expect_tx(bar_full(stage), 2u * (P.a_box_bytes + P.b_box_bytes));
            const uint32_t lead_full = mapa_u32(bar_full(stage), 0u);
            tma_load_5d_2sm(sa, &tmA, lead_full, c0, (int)tp.dx, y0 + (int)tp.dy, b0, (int)tp.plane);
            tma_load_2d_2sm(sa + TC_A_BYTES, &tmB, lead_full, c0, wrow0 + t * p.Np + nt * P.BN + (int)cta_rank * (P.BN / 2));
          } else {
            mbar_expect_tx(bar_full(stage), P.a_box_bytes + P.b_box_bytes);
            tma_load_5d(sa, &tmA, bar_full(stage), c0, (int)tp.dx, y0 + (int)tp.dy, b0, (int)tp.plane);
            // denoiser tiles hold [gate | hyper-bias | main | skip] column blocks: an h k-block only feeds (main, skip),
            // a c k-block only (gate, hyper-bias) -- load just that half of the weight tile
            const int brow = den_kind ? nt * P.BN + (kb < P.den_kb_h ? P.BN / 2 : 0) : wrow0 + t * p.Np + nt * P.BN;
            tma_load_2d(sa + TC_A_BYTES, &tmB, bar_full(stage), c0, brow);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs the issue loop (converged: every lane waits on the barriers), one elected lane issues (umma_elect).
    // DAMC_TC_ISSUE=lane0 at build time keeps the round-1 form (lane 0 alone in the loop) for A/B.
#ifdef DAMC_TC_ISSUE_LANE0
    if (lane == 0 && cta_rank == 0) {
#else
    if (cta_rank == 0) {
#endif
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, ++it) {
        int mt, nt, sp;
        decode(tile, mt, nt, sp);
        const int kb0 = ncls > 1 ? 0 : sp * P.kb_per_split, kb1 = min(P.kb_total, kb0 + P.kb_per_split);
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(bar_tempty(as), aphase ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * 256u;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + (uint32_t)stage * stage_bytes;
          const uint64_t adesc = make_sdesc(sa), bdesc = make_sdesc(sa + TC_A_BYTES);
          // denoiser: half-width MMA into the (main, skip) or the (gate, hyper-bias) column block of the accumulator
          const uint32_t d_blk = den_kind ? d_tmem + (uint32_t)(kb < P.den_kb_h ? P.BN / 2 : 0) : d_tmem;
          const bool first_kb = den_kind ? (kb == kb0 || kb == P.den_kb_h) : kb == kb0;
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // +32 bytes (16 bf16 | 8 tf32) along K inside the 128-byte swizzle row
            const uint32_t acc = (!first_kb || k > 0) ? 1u : 0u;
#ifdef DAMC_TC_ISSUE_LANE0
            if (F32) {
              if (CG == 2) umma_tf32_2sm(d_blk, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, acc);
              else umma_tf32(d_blk, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, acc);
            } else {
              if (CG == 2) umma_bf16_2sm(d_blk, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, acc);
              else umma_bf16(d_blk, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, acc);
            }
#else
            umma_elect<F32, CG>(d_blk, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), P.idesc, acc);
#endif
          }
#ifdef DAMC_TC_ISSUE_LANE0
          if (CG == 2) umma_commit_2sm(bar_empty(stage)); else umma_commit(bar_empty(stage));  // frees the smem slot
#else
          umma_commit_elect<CG>(bar_empty(stage));  // frees the smem slot
#endif
          if (++stage == P.stages) { stage = 0; phase ^= 1u; }
        }
        if (kb1 <= kb0) {  // empty K range (trailing split): nothing was accumulated -> signal with zeros impossible;
          // the host guarantees kb_per_split * (ksplit-1) < kb_total, so this cannot happen
        }
#ifdef DAMC_TC_ISSUE_LANE0
        if (CG == 2) umma_commit_2sm(bar_tfull(as)); else umma_commit(bar_tfull(as));  // accumulator complete
#else
        umma_commit_elect<CG>(bar_tfull(as));  // accumulator complete
#endif
      }
    }
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    constexpr int NG = EW / 4;        // warps per lane quarter
    const int grp = (warp - 2) >> 2;  // the NG warps of a quarter split the columns
    float loss_acc = 0.f;
    auto row_ctx = [&](int tile, int& nt, int& sp) {  // row of this thread inside the tile -> (chain, y, x) on the M grid
      int mt, b0, y0;
      decode(tile, mt, nt, sp);
      tile_origin(mt, b0, y0);
      const int r = q * 32 + lane;
      RowCtx rc;
      const int per_img = P.Ht * p.Wm;
      int bt, rem, yy, xx;
      if (P.per_img_shift >= 0) { bt = r >> P.per_img_shift; rem = r & (per_img - 1); }
      else P.fd_perimg.divmod(r, bt, rem);
      if (P.wm_shift >= 0) { yy = rem >> P.wm_shift; xx = rem & (p.Wm - 1); }
      else P.fd_wm.divmod(rem, yy, xx);
      rc.b = b0 + bt;
      rc.y = y0 + yy;
      rc.x = xx;
      rc.ok = r < P.tile_rows && rc.b < p.B && rc.y < p.Hm;
      rc.m = (rc.b * p.Hm + rc.y) * p.Wm + rc.x;
      return rc;
    };
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    if (P.stage_cols == 64) {
      // ---- staged path: flat software pipeline over (tile, segment) items of this warp ----
      const bool is_mask = p.epi.kind == EPI_DGRAD_MASK;
      const bool bits = p.epi.maskbits != nullptr;  // dgrad reads 8 bytes of sign bits per row segment, not 128
      const int nseg_w = (P.BN / 64 - grp + NG - 1) / NG;  // segments of a tile handled by this warp: grp, grp+NG, ...
      const int ntl = (total_tiles - unit + nunits - 1) / nunits;
      const uint32_t tempty_lead[2] = {CG == 2 ? mapa_u32(bar_tempty(0), 0u) : 0u, CG == 2 ? mapa_u32(bar_tempty(1), 0u) : 0u};
      auto release_acc = [&](int as) {  // this warp is done with accumulator `as` (reported to the MMA-issuing CTA)
        if (CG == 2) mbar_arrive_cluster(tempty_lead[as]); else mbar_arrive(bar_tempty(as));
      };
      const uint32_t my_stage = staging + (uint32_t)(warp - 2) * (32u * 128u);
      StagedEpi se;
      se.dbg = P.dbg;
      se.direct = P.direct;
      se.sub = lane / StagedEpi::CPR;
      se.j = lane % StagedEpi::CPR;
      if (nseg_w == 0) {
        for (int k = 0; k < ntl; ++k) {
          mbar_wait_relaxed(bar_tfull(k & 1), (uint32_t)(k >> 1) & 1u);
          __syncwarp();
          if (lane == 0) release_acc(k & 1);
        }
      } else {
        const int nitems = ntl * nseg_w;
        uint4 ra[StagedEpi::NIT], rb[StagedEpi::NIT];
        // (tile ordinal k, segment ordinal sg) of the current item and of the item whose rows are requested ahead
        struct ItemPos { int k, sg; };
        auto advance = [&](ItemPos& ip) { if (++ip.sg == nseg_w) { ip.sg = 0; ++ip.k; } };
        auto request = [&](uint4 (&r)[StagedEpi::NIT], const ItemPos& ip) {
          if (!is_mask || ip.k >= ntl) return;
          int nt, sp;
          const RowCtx rc = row_ctx(unit + ip.k * nunits, nt, sp);
          const int n_base = nt * P.BN + (grp + NG * ip.sg) * 64;
          if (n_base >= p.N) return;
          if (bits) { const uint2 mw = StagedEpi::load_mask_words(p, rc, n_base); r[0].x = mw.x; r[0].y = mw.y; }
          else se.request(r, StagedEpi::act_ptr_bits(p, rc), n_base);
        };
        constexpr int AHEAD = EW == 8 ? 2 : 1;  // 16 warps: one register set each (keeps the kernel spill-free)
        ItemPos cur{0, 0}, nxt{0, 0};
        request(ra, nxt);
        advance(nxt);
        if (AHEAD == 2) { request(rb, nxt); advance(nxt); }
        int nt = 0, sp = 0;
        // one (tile, segment) item; r holds its activation rows (requested AHEAD items ago) and is re-armed for item+AHEAD.
        // The two register sets alternate (no register copies: a copy would wait on the loads it moves).
        auto do_item = [&](uint4 (&r)[StagedEpi::NIT]) {
          const int k = cur.k, sg = cur.sg;
          const int as = k & 1;
          if (sg == 0) {
            const RowCtx rc = row_ctx(unit + k * nunits, nt, sp);
            se.template set_out<F32>(p, rc, sp);
          }
          const int seg = (grp + NG * sg) * 64, n_base = nt * P.BN + seg;
          uint2 mw = make_uint2(0u, 0u);
          if (is_mask) {
            if (bits) mw = make_uint2(r[0].x, r[0].y); else se.stash(r, my_stage);
            request(r, nxt);
            advance(nxt);
          }
          if (sg == 0) {
            mbar_wait_relaxed(bar_tfull(as), (uint32_t)(k >> 1) & 1u);
            tc_fence_after();
          }
          if (n_base < p.N) {
            int bq, brem = 0;
            if (!is_mask) P.fd_bias.divmod(n_base, bq, brem);
            se.template process<F32>(p, t_lane + (uint32_t)as * 256u + (uint32_t)seg, my_stage, lane, n_base, mw,
                       (P.bias_floats ? den_bias : p.epi.bias) + brem);
          }
          if (sg == nseg_w - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release_acc(as);
          }
          advance(cur);
        };
        if (AHEAD == 2) {
          for (int item = 0; item < nitems; item += 2) {
            do_item(ra);
            if (item + 1 < nitems) do_item(rb);
          }
        } else {
          for (int item = 0; item < nitems; ++item) do_item(ra);
        }
      }
    } else if (den_kind) {
      // ---- denoiser layer: accumulator columns [gate Q | hyper-bias Q | main Q | skip Q], Q = BN/4 features per tile ----
      const DenEpi& d = p.epi.den;
      const bool fin = p.epi.kind == EPI_DEN_FINAL;
      const int Q = P.BN >> 2, nchunk = Q >> 4;  // 16-feature chunks; the NG warps of a lane quarter alternate over them
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, ++it) {
        int nt, sp;
        const RowCtx rc = row_ctx(tile, nt, sp);
        const int as = it & 1;
        mbar_wait_relaxed(bar_tfull(as), (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = t_lane + (uint32_t)as * 256u;
#pragma unroll 1
        for (int ch = grp; ch < nchunk; ch += NG) {
          const uint32_t t0 = t_row + (uint32_t)(ch << 4);
          uint32_t vg[16], vh[16], vm[16], vs[16];
          tmem_ld16(t0, vg);
          tmem_ld16(t0 + (uint32_t)Q, vh);
          tmem_ld16(t0 + (uint32_t)(2 * Q), vm);
          tmem_ld16(t0 + (uint32_t)(3 * Q), vs);
          tmem_ld_wait();
          const int f0 = nt * Q + (ch << 4);  // first output feature of this chunk
          float o[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 bq = *reinterpret_cast<const float4*>(den_bias + 4 * (f0 + i));   // (bg, 0, b, bs)
            const float gate = __uint_as_float(vg[i]) + bq.x;
            o[i] = fmaf(__uint_as_float(vm[i]) + bq.z, __fdividef(1.f, 1.f + __expf(-gate)),
                        __uint_as_float(vh[i]) + __uint_as_float(vs[i]) + bq.w);
          }
          if (!rc.ok) continue;
          if (fin) {
#pragma unroll
            for (int j = 0; j < 4; ++j) den_final_quad(d, rc.b, f0 + 4 * j, o + 4 * j);
          } else {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float a = o[2 * j], b = o[2 * j + 1];
              a = a > 0.f ? a : 0.01f * a;
              b = b > 0.f ? b : 0.01f * b;
              w[j] = pack2(p.op_fp16, a, b);
            }
            // 16 features = 32 bytes of this lane's operand row: one full-sector 256-bit store per destination
            uint16_t* o1 = reinterpret_cast<uint16_t*>(d.dst1) + (long long)rc.b * d.ld1 + d.off1 + f0;
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o1), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                         "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            if (d.dst2) {
              uint16_t* o2 = reinterpret_cast<uint16_t*>(d.dst2) + (long long)rc.b * d.ld2 + d.off2 + f0;
              asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o2), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                           "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_u32(bar_tempty(as), 0u)); else mbar_arrive(bar_tempty(as)); }
      }
    } else {
      int it = 0;
      for (int tile = unit; tile < total_tiles; tile += nunits, ++it) {
        int nt, sp;
        const RowCtx rc = row_ctx(tile, nt, sp);
        const int as = it & 1;
        mbar_wait_relaxed(bar_tfull(as), (uint32_t)(it >> 1) & 1u);
        tc_fence_after();
        const uint32_t t_row = t_lane + (uint32_t)as * 256u;
        for (int c = grp * 32; c < P.BN; c += 32 * NG) {
          uint32_t v[32];
          if (c + 32 <= P.BN) {
            tmem_ld32(t_row + (uint32_t)c, v);
            tmem_ld_wait();
            epi_chunk16<F32>(p, rc, sp, nt * P.BN + c, v, loss_acc);
            epi_chunk16<F32>(p, rc, sp, nt * P.BN + c + 16, v + 16, loss_acc);
          } else {
            tmem_ld16(t_row + (uint32_t)c, v);
            tmem_ld_wait();
            epi_chunk16<F32>(p, rc, sp, nt * P.BN + c, v, loss_acc);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(mapa_u32(bar_tempty(as), 0u)); else mbar_arrive(bar_tempty(as)); }
      }
    }
    if (p.epi.kind == EPI_FWD_LAST && p.epi.loss != nullptr) {
      loss_acc = warp_sum(loss_acc);
      if (lane == 0 && loss_acc != 0.f) atomicAdd(p.epi.loss, loss_acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // neither CTA may retire while the pair's MMAs / multicast arrives can still touch it
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, TC_TMEM_COLS); else tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)sym;
  }
  return fn;
}

int tc_available() { return get_encode() != nullptr; }

// host-side check of the magic-number division the kernel's tile / row decoding relies on (same formula as FastDiv::div)
int tc_selftest_fastdiv() {
  auto div = [](const FastDiv& f, int n) { return f.mul ? (int)((uint32_t)(((unsigned long long)(uint32_t)n * f.mul) >> 31) >> f.shr) : n; };
  const int ns[] = {0, 1, 2, 3, 63, 64, 65, 127, 128, 1000, 4095, 4096, 65535, 65536, 1000003, (1 << 24) + 7, 0x7ffffffe, 0x7fffffff};
  uint32_t x = 12345u;
  for (int d = 1; d <= 70000; d = d < 4200 ? d + 1 : d + 977) {
    const FastDiv f = make_fastdiv(d);
    for (int n : ns)
      if (div(f, n) != n / d) DAMC_FAIL(DAMC_ERR_INVALID, "FastDiv: %d / %d -> %d", n, d, div(f, n));
    for (int i = 0; i < 64; ++i) {
      x = x * 1664525u + 1013904223u;
      const int n = (int)(x >> 1);
      if (div(f, n) != n / d) DAMC_FAIL(DAMC_ERR_INVALID, "FastDiv: %d / %d -> %d", n, d, div(f, n));
    }
  }
  return DAMC_OK;
}

// K-major operand matrix [rows][cols] (16-bit elements, cols contiguous) as a 2-D tensor map with a {64, box_rows} box and
// the 128-byte swizzle the UMMA descriptors of this engine expect
int tc_encode_2d(void* tensor_map, int fp16, const void* base, int cols, int rows, int box_rows) {
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(tensor_map);
  EncodeTiledFn enc = get_encode();
  if (!enc) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(tm, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                         dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d (cols=%d rows=%d box=%d)", (int)r, cols, rows, box_rows);
  return DAMC_OK;
}

// NHWC activation tensor [B][Hm][Wm][Cs] (16-bit elements) as the 5-D map of the A operand: box {64 ch, Wm, Ht, 1, 1}, one
// 128-byte swizzle row per pixel (used by the fused last-layer kernel, gen_last.cu)
int tc_encode_act(void* tensor_map, int precision, const void* base, int Cs, int Wm, int Hm, int B, int Ht) {
  CUtensorMap* tm = reinterpret_cast<CUtensorMap*>(tensor_map);
  EncodeTiledFn enc = get_encode();
  if (!enc) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (!is_tc_precision(precision)) DAMC_FAIL(DAMC_ERR_INVALID, "tc_encode_act: 16-bit operand precisions only");
  const cuuint64_t dims[5] = {(cuuint64_t)Cs, (cuuint64_t)Wm, (cuuint64_t)Hm, (cuuint64_t)B, 1};
  const cuuint64_t row = (cuuint64_t)Cs * 2;
  const cuuint64_t strides[4] = {row, row * Wm, row * Wm * Hm, row * Wm * Hm * B};
  const cuuint32_t box[5] = {(cuuint32_t)TC_BK, (cuuint32_t)Wm, (cuuint32_t)Ht, 1, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(tm, precision == DAMC_PREC_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d (Cs=%d W=%d H=%d B=%d)", (int)r, Cs, Wm, Hm, B);
  return DAMC_OK;
}

// few chains: narrower N tiles spread one layer over more SMs and shorten each CTA's serial load -> MMA -> epilogue chain
int tc_den_tile_width(int B, int Np) {
  int bn = Np < 256 ? Np : 256;
  while (bn > 64 && (long long)ceil_div(B, TC_BM) * (Np / bn) < 96) bn /= 2;
  return bn;
}

struct TcLaunch {
  CUtensorMap tmA, tmB;
  TcParams P;
  int ew, cg;
  size_t smem;
};

static int tc_prepare_into(const GemmPlan& p, int precision, TcLaunch* L) {
  EncodeTiledFn enc = get_encode();
  if (!enc) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const bool f32 = precision == DAMC_PREC_TF32;
  const int bk = f32 ? TC_BK / 2 : TC_BK;          // elements per 128-byte swizzle row
  const cuuint64_t esz = f32 ? 4 : 2;
  if (p.Cs % bk) DAMC_FAIL(DAMC_ERR_INVALID, "tcgen05 GEMM: K per tap (%d) must be a multiple of %d", p.Cs, bk);
  if (p.Np % 16) DAMC_FAIL(DAMC_ERR_INVALID, "tcgen05 GEMM: padded N (%d) must be a multiple of 16", p.Np);
  if (p.Wm > TC_BM) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: pixel-grid width %d > %d", p.Wm, TC_BM);
  if (p.ncls > 1 && (p.ncls != 4 || p.ntaps != 4 || p.ksplit != 1 || p.epi.kind != EPI_FWD_ACT || p.Np % 64))
    DAMC_FAIL(DAMC_ERR_INVALID, "tcgen05 GEMM: merged parity classes need the k4-s2-p1 forward shape");
  TcParams& P = L->P;
  P = TcParams{};
  P.plan = p;
  const bool fp16 = precision == DAMC_PREC_FP16;
  P.plan.op_fp16 = fp16 ? 1 : 0;
  P.plan.op_f32 = f32 ? 1 : 0;
  P.bk = bk;
  P.tf32 = f32 ? 1 : 0;
  const CUtensorMapDataType tm_dtype = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                           : fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  P.BN = p.Np < 256 ? p.Np : 256;
  const bool den_kind = p.epi.kind == EPI_DEN_LAYER || p.epi.kind == EPI_DEN_FINAL;
  if (f32 && (den_kind || (p.epi.kind == EPI_DGRAD_MASK && p.epi.maskbits == nullptr)))
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: the tf32 mode covers the generator GEMMs with 1-bit masks only");
  if (den_kind) {
    P.BN = tc_den_tile_width(p.B, p.Np);
    if (p.epi.den.bn != P.BN) DAMC_FAIL(DAMC_ERR_INVALID, "tcgen05 GEMM: denoiser weights packed for tile width %d, launch uses %d", p.epi.den.bn, P.BN);
  }
  if (p.Hm * p.Wm >= TC_BM) {
    P.Bt = 1;
    P.Ht = TC_BM / p.Wm;
    P.tiles_per_img = ceil_div(p.Hm, P.Ht);
    P.m_tiles = p.B * P.tiles_per_img;
  } else {
    P.Ht = p.Hm;
    P.Bt = TC_BM / (p.Hm * p.Wm);
    P.tiles_per_img = 1;
    P.m_tiles = ceil_div(p.B, P.Bt);
  }
  // (Tried: 256 x 128 pair tiles when 256 x 256 ones quantise badly, e.g. 3.46 waves at 128 chains.  They move 1.5x the
  // L2 -> SM bytes per FLOP and the pair GEMMs are L2 -> SM bound: 21.9 ms vs 18.4 ms at 128 chains.  Dropped.)
  P.n_tiles = ceil_div(p.Np, P.BN);
  P.tile_rows = p.Wm * P.Ht * P.Bt;
  auto log2_or_neg = [](int v) { int l = 0; while ((1 << l) < v) ++l; return (1 << l) == v ? l : -1; };
  P.wm_shift = log2_or_neg(p.Wm);
  P.per_img_shift = log2_or_neg(P.Ht * p.Wm);
  P.kb_per_tap = p.Cs / bk;
  P.kb_total = p.ntaps * P.kb_per_tap;
  P.kb_per_split = ceil_div(P.kb_total, p.ksplit);
  if ((long long)P.kb_per_split * (p.ksplit - 1) >= P.kb_total)
    DAMC_FAIL(DAMC_ERR_INVALID, "tcgen05 GEMM: ksplit %d leaves an empty K range (%d blocks)", p.ksplit, P.kb_total);
  P.fd_ntiles = make_fastdiv(P.n_tiles);
  P.fd_ncls = make_fastdiv(p.ncls > 1 ? p.ncls : 1);
  P.fd_tpi = make_fastdiv(P.tiles_per_img);
  P.fd_perimg = make_fastdiv(P.Ht * p.Wm);
  P.fd_wm = make_fastdiv(p.Wm);
  P.a_box_bytes = (uint32_t)P.tile_rows * 128u;   // one 128-byte swizzle row per pixel / weight row
  P.b_box_bytes = (uint32_t)P.BN * 128u;
  P.stage_cols = 0;
  P.cls_inner = getenv("DAMC_TC_CLS_OUTER") ? 0 : 1;
  P.dbg = getenv("DAMC_TC_DBG") ? atoi(getenv("DAMC_TC_DBG")) : 0;
  if ((p.epi.kind == EPI_FWD_ACT || p.epi.kind == EPI_DGRAD_MASK) && !getenv("DAMC_TC_NOSTAGE")) {
    if (P.BN % 64 == 0 && p.N % 64 == 0 && (p.epi.kind != EPI_FWD_ACT || p.epi.bias_mod % 64 == 0)) P.stage_cols = 64;
  }
  if (!P.stage_cols && p.epi.maskbits != nullptr && (p.epi.kind == EPI_FWD_ACT || p.epi.kind == EPI_DGRAD_MASK))
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: bit masks need the staged epilogue (N %% 64 == 0); set DAMC_TC_NOBITS=1");
  const int ew = (P.stage_cols && P.kb_per_split <= 4 && P.BN >= 256 && !getenv("DAMC_TC_EW8")) ? 16 : 8;
  // CTA-pair MMA for the MMA-bound launches: full 256-wide N tiles, long K, an even grid of pairs
  static const bool allow_2sm = []{ const char* e = getenv("DAMC_TC_2SM"); return !(e && e[0] == '0'); }();
  // (also for genuinely 128-wide layers -- SVHN 256->128, CelebA-HQ 256->128 forward: a 256 x 128 pair tile stages 24 KB per
  //  k-block and CTA instead of 32 KB for two independent 128 x 128 tiles, and each MMA reads half of the weight rows)
  // Measured (profiles/r02_launches_*): no gain -- SVHN 256->128 forward 1 440 us as pairs vs 1 313 us as single CTAs, CelebA-HQ
  // 744 vs 666 us; both forms move > 21 TB/s of L2 -> SM operand traffic at full tensor rate, which the fabric does not
  // deliver (17-18 TB/s measured), so 128-wide layers are L2 -> SM bound either way.  Off by default.
  static const bool pair_n128 = []{ const char* e = getenv("DAMC_TC_PAIR128"); return e && e[0] == '1'; }();
  const bool pair_shape = (P.BN == 256 && p.Np % 256 == 0) || (pair_n128 && P.BN == 128 && p.Np == 128 && !den_kind);
  const int cg = (allow_2sm && ew == 8 && pair_shape && P.kb_per_split >= 16 && P.m_tiles >= 2) ? 2 : 1;
  P.fd_munits = make_fastdiv(ceil_div(P.m_tiles, cg));
  if (cg == 2) P.b_box_bytes /= 2;  // each CTA of the pair stages half of the 256 weight rows
  if (den_kind) P.b_box_bytes /= 2; // denoiser: a k-block loads only the column block (half of the tile's rows) it feeds
  P.b_stage_bytes = (int)align_up(P.b_box_bytes, 1024);
  const int stage_bytes = TC_A_BYTES + P.b_stage_bytes;
  if (den_kind && (P.BN % 64 || p.Np % P.BN || p.ksplit != 1 || p.ncls > 1 || p.epi.den.din % TC_BK || p.epi.den.din >= p.Cs))
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: denoiser epilogue needs 4*dout (%d) to be a multiple of 256", p.Np);
  P.den_kb_h = den_kind ? p.epi.den.din / TC_BK : 0;
  // per-warp staging tiles, then (denoiser) the layer's bias quads
  P.bias_src = nullptr;
  P.bias_floats = 0;
  P.bias_off = 0;
  P.fd_bias = make_fastdiv(1);
  if (den_kind) { P.bias_src = p.epi.den.bias4; P.bias_floats = p.Np; P.bias_off = 0; }
  else if (P.stage_cols && p.epi.kind == EPI_FWD_ACT) {
    P.bias_src = p.epi.bias; P.bias_floats = p.epi.bias_mod; P.bias_off = -1;  // set below: right behind the staging tiles
    P.fd_bias = make_fastdiv(p.epi.bias_mod);
  }
  // the sign-from-activations dgrad variant (DAMC_TC_NOBITS) stashes activation rows in the staging tiles; every other
  // staged launch writes its rows directly from registers and needs no tiles
  P.direct = (P.stage_cols && !(P.dbg & 4) && !(p.epi.kind == EPI_DGRAD_MASK && p.epi.maskbits == nullptr)) ? 1 : 0;
  const int tiles_bytes = P.stage_cols ? (P.direct ? 0 : ew * TC_STAGING_PER_WARP) : 0;
  auto stages_for = [&](int sb) { return std::min(TC_MAX_STAGES, (int)((227 * 1024 - 2048 - sb) / stage_bytes)); };
  if (!den_kind && stages_for(tiles_bytes + P.bias_floats * 4 + 128) < stages_for(tiles_bytes + 128))
    P.bias_floats = 0;  // a pipeline stage is worth more than the smem bias: the epilogue reads the bias through L1 instead
  if (P.bias_off < 0) P.bias_off = tiles_bytes;
  const int staging_bytes = tiles_bytes + P.bias_floats * 4 + 128;
  P.stages = stages_for(staging_bytes);
  if (P.stages < 2) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: not enough shared memory for a pipeline (BN=%d)", P.BN);
  // instruction descriptor: D=f32 (1<<4), A=B=bf16 (1<<7, 1<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
  const uint32_t opfmt = f32 ? 2u : fp16 ? 0u : 1u;  // operand format: 0 = f16, 1 = bf16 (kind::f16); 2 = tf32 (kind::tf32)
  const int mma_n = den_kind ? P.BN / 2 : P.BN;
  P.idesc = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)((TC_BM * cg) >> 4) << 24);

  CUtensorMap& tmA = L->tmA;
  CUtensorMap& tmB = L->tmB;
  {
    const int nplanes = [&] { int mx = 0; for (int t = 0; t < p.ntaps; ++t) mx = std::max(mx, (int)p.taps[t].plane); return mx + 1; }();
    const cuuint64_t dims[5] = {(cuuint64_t)p.Cs, (cuuint64_t)p.Wm, (cuuint64_t)p.Hm, (cuuint64_t)p.B, (cuuint64_t)nplanes};
    const cuuint64_t row = (cuuint64_t)p.Cs * esz;
    const cuuint64_t plane_bytes = nplanes > 1 ? (cuuint64_t)p.plane_stride * esz : row * p.Wm * p.Hm * p.B;
    const cuuint64_t strides[4] = {row, row * p.Wm, row * p.Wm * p.Hm, plane_bytes};
    const cuuint32_t box[5] = {(cuuint32_t)bk, (cuuint32_t)p.Wm, (cuuint32_t)P.Ht, (cuuint32_t)P.Bt, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(&tmA, tm_dtype, 5, const_cast<void*>(p.A), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d (Cs=%d W=%d H=%d B=%d planes=%d)", (int)r, p.Cs, p.Wm, p.Hm, p.B, nplanes);
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)p.Cs, (cuuint64_t)(p.ncls > 1 ? p.ncls : 1) * p.ntaps * p.Np};
    const cuuint64_t strides[1] = {(cuuint64_t)p.Cs * esz};
    const cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)(den_kind ? P.BN / 2 : P.BN / cg)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmB, tm_dtype, 2, const_cast<void*>(p.Wtc), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DAMC_FAIL(DAMC_ERR_CUDA, "cuTensorMapEncodeTiled(W) failed: %d (Cs=%d rows=%d BN=%d)", (int)r, p.Cs, p.ntaps * p.Np, P.BN);
  }
  L->smem = (size_t)P.stages * stage_bytes + 8 * (2 * P.stages + 4) + 16 + 1024 + staging_bytes;
  L->ew = ew;
  L->cg = cg;
  return DAMC_OK;
}

int tc_launch(TcLaunch* L, cudaStream_t stream) {
  const TcParams& P = L->P;
  const GemmPlan& p = P.plan;
  const size_t smem = L->smem;
  const int ew = L->ew, cg = L->cg;
  // per-device launch state: the dynamic-smem opt-in is a per-device function attribute and the persistent grid is sized
  // from the device's own SM count (one process may drive several GPUs; MCMC.py picks the device per tensor)
  constexpr int kMaxDev = 64;
  static std::atomic<int> sms_of[kMaxDev];
  int dev = 0;
  DAMC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDev) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "tcgen05 GEMM: device ordinal %d out of range", dev);
  int num_sms = sms_of[dev].load(std::memory_order_acquire);
  if (!num_sms) {
    const int big = 227 * 1024;
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<8, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<16, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<8, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<8, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<16, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaFuncSetAttribute(convgemm_tc_kernel<8, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    DAMC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    sms_of[dev].store(num_sms, std::memory_order_release);   // idempotent: a racing thread repeats the same calls
  }
  static const bool use_pdl = []{ const char* e = getenv("DAMC_TC_PDL"); return !(e && e[0] == '0'); }();
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  if (cg == 2) {
    const int units = ceil_div(P.m_tiles, 2) * P.n_tiles * p.ksplit * (p.ncls > 1 ? p.ncls : 1);
    cfg.gridDim = dim3(2 * std::min(units, num_sms / 2));
    cfg.blockDim = dim3(64 + 32 * 8);
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
    cfg.numAttrs = na;
    if (P.tf32) DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<8, 2, true>, L->tmA, L->tmB, L->P));
    else DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<8, 2, false>, L->tmA, L->tmB, L->P));
    return DAMC_OK;
  }
  const int total = P.m_tiles * P.n_tiles * p.ksplit * (p.ncls > 1 ? p.ncls : 1);
  cfg.gridDim = dim3(std::min(total, num_sms));
  cfg.blockDim = dim3(64 + 32 * ew);
  cfg.numAttrs = na;
  if (P.tf32) {
    if (ew == 16) DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<16, 1, true>, L->tmA, L->tmB, L->P));
    else DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<8, 1, true>, L->tmA, L->tmB, L->P));
  } else {
    if (ew == 16) DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<16, 1, false>, L->tmA, L->tmB, L->P));
    else DAMC_CUDA(cudaLaunchKernelEx(&cfg, convgemm_tc_kernel<8, 1, false>, L->tmA, L->tmB, L->P));
  }
  return DAMC_OK;
}


int tc_prepare(const GemmPlan& p, int precision, TcLaunch** out) {
  TcLaunch* L = new TcLaunch();
  const int r = tc_prepare_into(p, precision, L);
  if (r != DAMC_OK) { delete L; return r; }
  *out = L;
  return DAMC_OK;
}
GemmPlan* tc_plan(TcLaunch* l) { return &l->P.plan; }
void tc_free(TcLaunch* l) { delete l; }

int launch_gemm_tc(const GemmPlan& p, int precision, cudaStream_t stream) {
  TcLaunch L;
  DAMC_TRY(tc_prepare_into(p, precision, &L));
  return tc_launch(&L, stream);
}

}  // namespace damc
