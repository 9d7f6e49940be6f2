// denoiser_seq.cu -- DAMC reverse steps with the context branch hoisted out of the sequential loop (16-bit tensor-core modes).
//
// One reverse step of _netQ_U.forward (reference workspace/src/diffusion_net.py:597-620; Q.p = Diffusion_UnetA :463-533, layer
// ConcatSquashLinearSkipCtx :417-445) is 7 layers   out = (W h + b) * sigmoid(Wg c + bg) + (Wb c) + (Ws h + bs),
// c = SiLU(Wc [temb, xemb] + bc).  The gate and the hyper-bias depend on the chain's xemb and on the step index only -- not on z --
// so 42 % of a step's FLOPs (and every ctx activation) do not belong to the chain of dependent operations at all:
//   * den_seq_kernel<1> ("gate pass"): for a window of steps, G[t][b][f] = half2(s, Wb c + bs + b s), s = sigmoid(Wg c + bg), for every layer,
//     one CTA per (step, 128-chain tile): the ctx activations c = SiLU(cx[b] + ct[t]) are formed in shared memory as K-major
//     operand tiles (never in HBM), the (gate | hyper-bias) weights stream through a TMA ring, N = 256 tcgen05 MMAs.  Fully
//     parallel over steps x chains.
//   * den_seq_kernel<0> ("step pass"): ONE CTA owns 128 chains for ALL steps of the window.  Per step: the Fourier embedding
//     phase z.B as a split-fp16 tensor-core GEMM (z = zh + zl, B = Bh + Bl; zh Bh + zl Bh + zh Bl, fp32 accumulate: 2^-22
//     relative), then the seven (main | skip) GEMMs.  Activations never leave the SM: every epilogue writes the next layer's
//     K-major SWIZZLE_128B operand blocks into one of two 64 KB buffers (X / Y); the U-net skips of the first two layers go
//     through L2 and come back by TMA, the third is still resident when it is needed.  Only the 16-bit (main | skip) weights
//     (1.66 MB per step) stream from L2; the epilogue reads G, forms out, LeakyReLU, and the last layer applies
//     eps = z + out and the reverse update of z in fp32 (same arithmetic as the other schedules: den_final_vals).
// This replaces 8 launches per step (per-layer schedule) / the 8-CTA cluster kernel: no launch or cluster-barrier latency
// between layers, no operand round trips through L2, N = 256 MMAs instead of N = 64.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>

#include "damc_common.cuh"
#include "damc_internal.h"
#include "tc_ptx.cuh"

namespace damc {

constexpr int SQ_STEP_NWQ = 2;         // step pass: worker warps per TMEM lane quarter (each thread: one chain row, 128 / NWQ features of a tile)
constexpr int SQ_THREADS = 64 + 128 * SQ_STEP_NWQ;   // warp 0: TMA, warp 1: MMA issuer + TMEM, then the workers
constexpr int SQ_GATE_PW = 8;          // gate pass: producer warps (warps 10 ...) that form the ctx operand tiles, 128 / SQ_GATE_PW chain rows each
constexpr int SQ_THREADS_GATE = 320 + 32 * SQ_GATE_PW;   // warps 2-9: epilogue (two per lane quarter), then the producers
constexpr int SQ_BLK = 128 * 128;      // one K-major k-block of a 128-row operand tile (64 columns x 16 bit): 16 KB
constexpr int SQ_NBLK = 8;             // X = blocks 0-3, Y = blocks 4-7
constexpr int SQ_WSTAGE = 256 * 128;   // one weight k-block: 256 rows x 128 B
constexpr int SQ_PF_AHEAD = 3;       // tiles between the L2 prefetch of a tile's gate words and the tile (the producer itself runs up to 3 k-blocks ahead)
constexpr int SQ_WBOX = 64;          // weight rows per TMA request
constexpr int SQ_STAGES_STEP = 3;   // step pass: the weight stream feeds the chain of dependent MMAs
constexpr int SQ_STAGES_GATE = 3;   // (two stages leave the gate pass bound by the ~2 us a 256-row weight box takes to arrive)
constexpr int SQ_MAXCSUM = 65535;   // tile offsets are 16-bit
constexpr int SQ_MAXKB = 8;
constexpr int SQ_MAXTILES = 12;
constexpr int SQ_EMB_MAP = 7;
enum { SQ_EMB = 0, SQ_LAYER = 1, SQ_FINAL = 2, SQ_GATE = 3 };
enum { SQ_W_NONE = 0, SQ_W_H0 = 1, SQ_W_H1 = 2, SQ_W_STG = 3, SQ_W_C = 4 };

struct SqK { unsigned char ablk, wait; unsigned short wcol; };
// U-net skip rows leave by TMA store from the operand blocks the epilogue has just written: issued by the MMA thread after the
// hready wait of k-block `kb` (two 64-column blocks from `blk0` to columns `col0`, `col0 + 64` of skip tensor `sel`); kb = 255: none
struct SqSt { unsigned char kb, sel, blk0, col0; };
struct SqTile {
  SqK k[SQ_MAXKB];
  SqSt st[2];
  unsigned char rdwait, fullwait, pad0, pad1;   // before this tile's accumulator commit: rdwait = 1 + N: wait until at most N store
                                                // groups still read shared memory; fullwait: until every store has completed
  unsigned char nkb, wmap, kind, layer;
  unsigned short wrow0, wrows;
  unsigned char oblk, defer, skipsel, commit_ldone;
  unsigned char stage_nblk, stage_blk0, stage_map, layer_last;
  unsigned short stage_col0, goff, ocol0;
  unsigned char gfence, xcommit;   // xcommit: arrive on bar_xfree once the MMAs of the first `xcommit` k-blocks are complete
};

struct SqParams {
  CUtensorMap tmW[8];      // per layer: (main | skip) rows [step pass] or (gate | hyper-bias) rows [gate pass]; [7]: embedding
  CUtensorMap tmSkip[2];   // skip tensors of layers 0 / 1, [Bpad][dout]
  SqTile tile[SQ_MAXTILES];
  int ntiles;
  int B, Bpad, nz, csum, T;
  int s0, nsteps;          // steps [s0, s0 + nsteps) of the sampler (execution order); G holds exactly this window
  int fp16, residual, use_philox;
  uint32_t idesc_l, idesc_e;
  float* z;
  const float4* noiseT;    // this window's normals, tile-transposed [step][chain tile][nz/4][128 rows] (null: no noise)
  float4* zT;              // the chains' z in the same tile-transposed layout [chain tile][nz/4][128 rows] between the steps
  int with_noise;
  unsigned long long seed, chain0;
  const float* coef;       // [T][8] device
  const float* cx;         // [B][csum]
  const float* ct;         // [T][csum]
  // Every per-row stream of the epilogues is tile-transposed -- [..][16-byte word][128 rows] -- so that the 32 lanes of a warp (32
  // consecutive rows, one TMEM lane each) read 512 contiguous bytes.  Row-major rows would make every 16-byte load touch 32 lines:
  // 4 096 LSU wavefronts per tile, which also starve the MMA / TMA threads' own shared-memory instructions (measured: 3.5 us per tile).
  uint4* G;                // [nsteps][chain tile][csum/4][128] half2(s, hb + bs + b s) x 4
  void* skip[2];
  int skip_ld[2];
  const float* bias3[DEN_LAYERS];   // [3][dout]: b_main, b_skip, b_gate
  int dout[DEN_LAYERS], coff[DEN_LAYERS];
  unsigned long long* tlog;   // DAMC_SQ_DBG=1: time stamps of CTA 0, second pass (8 per tile)
};

static_assert(sizeof(SqParams) <= 4096, "kernel parameter block");

__device__ __forceinline__ unsigned long long sq_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// tcgen05.mma / tcgen05.commit issued by ONE elected lane of a CONVERGED warp.  Inside `if (lane == 0)` the compiler has to wrap
// every such instruction (uniform-datapath operands) into an elect-and-retry loop and re-materialise its uniform registers; with
// the whole warp running the loop and the election inside the asm statement, a k-block costs the issuing warp half the instructions
// -- and that warp's instruction stream, not the tensor pipe, paced the MMA-bound stretches (0.55 us per k-block).
template <bool PAIR>
__device__ __forceinline__ void sq_umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred pe, pa;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred pe, pa;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "setp.ne.b32 pa, %4, 0;\n\t"
        "@pe tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pa;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void sq_commit(uint32_t bar) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
        ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred pe;\n\t"
        "elect.sync _|pe, 0xffffffff;\n\t"
        "@pe tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar) : "memory");
}
// pull a contiguous global range into L2 (no destination in the SM): the epilogues' per-tile streams are requested only a tile's MMA
// time before they are needed -- less than a DRAM round trip under load
__device__ __forceinline__ void sq_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// exp2 without the denormal-range fix-up __expf carries (3 of its 5 instructions): an argument below -126 flushes to 0, which is the
// right limit for 1 / (1 + e^-a); the gate pass evaluates 2 816 of these per chain and step and is SFU / issue bound
__device__ __forceinline__ float sq_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sq_silu_fast(float a) { return __fdividef(a, 1.f + sq_ex2(-1.4426950408889634f * a)); }
__device__ __forceinline__ void sq_st16(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// byte offset of the 16-byte chunk holding columns [col, col + 8) of row r inside the block set starting at block 0
__device__ __forceinline__ uint32_t sq_chunk(int blk, int r, int col) {
  return (uint32_t)blk * SQ_BLK + (uint32_t)r * 128u + (uint32_t)((((col & 63) >> 3) ^ (r & 7)) << 4);
}

// FP16: operand type fp16 (else bf16), a compile-time constant in every conversion.  PAIR (gate pass only): two CTAs of a cluster work
// on two items in lockstep as ONE tcgen05 unit (cta_group::2, M = 256: each CTA's 128 chain rows against weight tiles of which each CTA
// holds -- and receives from L2 -- only half): the weight ring is 6 stages of 16 KB instead of 3 of 32 KB, and a ring delivers
// (bytes in flight) / (1.75 us of TMA latency).  The leader CTA's MMA thread issues for the pair; its barriers collect both CTAs' arrivals.
template <int MODE, bool FP16, bool PAIR>
__global__ void __launch_bounds__(MODE == 0 ? SQ_THREADS : SQ_THREADS_GATE, 1) den_seq_kernel(const __grid_constant__ SqParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* const gen_base = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t blocks = base, ring = base + (uint32_t)SQ_NBLK * SQ_BLK;
  constexpr int SQ_STAGES = (MODE == 0 ? SQ_STAGES_STEP : SQ_STAGES_GATE) * (PAIR ? 2 : 1);
  constexpr uint32_t WST = PAIR ? SQ_WSTAGE / 2 : SQ_WSTAGE;   // bytes of a weight stage in this CTA
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  constexpr int NWQ = MODE == 0 ? SQ_STEP_NWQ : 2, NWORK = 4 * NWQ;   // worker warps per lane quarter / in all
  constexpr int FPW = 128 / NWQ, NCH = FPW / 16;                      // features of a tile per worker thread, in 16-feature chunks
  const uint32_t off_bars = (uint32_t)SQ_NBLK * SQ_BLK + (uint32_t)SQ_STAGES * WST;
  const uint32_t bars = base + off_bars;
  auto bar_wfull = [&](int s) { return bars + 8u * s; };
  auto bar_wempty = [&](int s) { return bars + 8u * (SQ_STAGES + s); };
  auto bar_accfull = [&](int s) { return bars + 8u * (2 * SQ_STAGES + s); };
  auto bar_accempty = [&](int s) { return bars + 8u * (2 * SQ_STAGES + 2 + s); };
  auto bar_hready = [&](int s) { return bars + 8u * (2 * SQ_STAGES + 4 + s); };
  auto bar_cready = [&](int s) { return bars + 8u * (2 * SQ_STAGES + 6 + s); };
  const uint32_t bar_stg = bars + 8u * (2 * SQ_STAGES + 8), bar_ldone = bars + 8u * (2 * SQ_STAGES + 9);
  const uint32_t bar_xfree = bars + 8u * (2 * SQ_STAGES + 10);
  auto bar_cfree = [&](int s) { return bars + 8u * (2 * SQ_STAGES + 11 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * SQ_STAGES + 13);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(gen_base + off_bars + 8u * (2 * SQ_STAGES + 13));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 8; ++i) prefetch_tmap(&P.tmW[i]);
    if (MODE == 0) { prefetch_tmap(&P.tmSkip[0]); prefetch_tmap(&P.tmSkip[1]); }
    for (int s = 0; s < SQ_STAGES; ++s) { mbar_init(bar_wfull(s), 1); mbar_init(bar_wempty(s), 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar_accfull(s), 1); mbar_init(bar_accempty(s), NWORK * (PAIR ? 2 : 1)); mbar_init(bar_hready(s), NWORK);
      mbar_init(bar_cready(s), SQ_GATE_PW * (PAIR ? 2 : 1));
      mbar_init(bar_cfree(s), 1);
    }
    mbar_init(bar_stg, 1);
    mbar_init(bar_ldone, 1);
    mbar_init(bar_xfree, 1);
    fence_barrier_init();
  }
  if (warp == 1) { if (PAIR) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before any remote arrive / TMA signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  // passes: step pass = the steps of the window for this CTA's chain tile; gate pass = this CTA's (step, chain tile) items
  const int tiles_b = P.Bpad >> 7;
  const int nitems = MODE == 0 ? P.nsteps : P.nsteps * tiles_b;
  const int first = MODE == 0 ? 0 : (int)blockIdx.x, stride = MODE == 0 ? 1 : (int)gridDim.x;
  // iterations of this CTA; a pair runs the same number in both CTAs (the one without an item of its own repeats the last item:
  // same values to the same addresses)
  const int niter = PAIR ? (nitems + stride - 1) / stride : (nitems - first + stride - 1) / stride;
  auto item_of = [&](int it) { return min(first + it * stride, nitems - 1); };
  // leader-side barriers as cluster addresses (for the non-leader's arrivals)
  auto lead = [&](uint32_t bar) { return PAIR ? mapa_u32(bar, 0u) : bar; };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, ldph = 0;
      auto prefetch_g = [&](int pitem, int pt) {   // gate words of tile pt of step pitem: 128 features x 128 rows, contiguous
        if (pitem >= nitems || P.tile[pt].kind == SQ_EMB) return;
        sq_prefetch_l2(P.G + (((size_t)pitem * tiles_b + blockIdx.x) * (size_t)(P.csum >> 2) + (P.tile[pt].goff >> 2)) * 128, 128u * 128u * 4u);
      };
      if (MODE == 0) for (int pt = 1; pt < SQ_PF_AHEAD; ++pt) prefetch_g(0, pt);
      for (int it = 0; it < niter; ++it) {
        const int item = item_of(it);
        for (int ti = 0; ti < P.ntiles; ++ti) {
          const SqTile& T = P.tile[ti];
          if (MODE == 0) {   // SQ_PF_AHEAD tiles ahead; the step's normals half a step ahead
            const int pt = ti + SQ_PF_AHEAD;
            if (pt < P.ntiles) prefetch_g(item, pt); else prefetch_g(item + 1, pt - P.ntiles);
            if (ti == 4 && P.noiseT)
              sq_prefetch_l2(P.noiseT + ((size_t)item * tiles_b + blockIdx.x) * (size_t)(P.nz >> 2) * 128, 128u * (uint32_t)P.nz * 4u);
          }
          if (MODE == 0 && T.stage_nblk) {   // U-net skip blocks come back from L2 once the buffer they land in is no longer read
            mbar_wait(bar_ldone, ldph);
            ldph ^= 1u;
            mbar_expect_tx(bar_stg, (uint32_t)T.stage_nblk * SQ_BLK);
            for (int j = 0; j < T.stage_nblk; ++j)
              tma_load_2d(blocks + (uint32_t)(T.stage_blk0 + j) * SQ_BLK, &P.tmSkip[T.stage_map], bar_stg, T.stage_col0 + 64 * j,
                          (int)blockIdx.x * 128);
          }
          // one ring stage per k-block; the embedding's 8 KB weight blocks travel four to a stage
          const int per = T.kind == SQ_EMB ? 4 : 1;
          for (int kb = 0; kb < T.nkb; kb += per) {
            const int nsub = min(per, T.nkb - kb);
            mbar_wait(bar_wempty(stage), phase ^ 1u);
            if (PAIR) {
              // this CTA's half of the tile's weight rows; both CTAs' loads report to the LEADER's full barrier, which expects the
              // bytes of the whole pair
              if (cta_rank == 0) mbar_expect_tx(bar_wfull(stage), (uint32_t)T.wrows * 128u);
              const uint32_t lead_full = mapa_u32(bar_wfull(stage), 0u);
              const int half_rows = T.wrows >> 1;
              for (int rb = 0; rb < half_rows; rb += SQ_WBOX)
                tma_load_2d_2sm(ring + (uint32_t)stage * WST + (uint32_t)rb * 128u, &P.tmW[T.wmap], lead_full, T.k[kb].wcol,
                                T.wrow0 + (int)cta_rank * half_rows + rb);
            } else {
              mbar_expect_tx(bar_wfull(stage), (uint32_t)nsub * T.wrows * 128u);
              // SQ_WBOX-row boxes: a stage arrives as several concurrent requests instead of one 256-row box walked row by row
              for (int j = 0; j < nsub; ++j)
                for (int rb = 0; rb < T.wrows; rb += SQ_WBOX)
                  tma_load_2d(ring + (uint32_t)stage * WST + (uint32_t)(j * T.wrows + rb) * 128u, &P.tmW[T.wmap], bar_wfull(stage),
                              T.k[kb + j].wcol, T.wrow0 + rb);
            }
            if (kb == 0 && P.tlog != nullptr && blockIdx.x == 0 && it == 1) P.tlog[10 * ti + 8] = sq_now();
            if (++stage == SQ_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (cta_rank == 0) {   // the WHOLE warp runs this loop (converged): waits by every lane, issue by an elected one (sq_umma)
      int stage = 0;
      uint32_t phase = 0, hph[2] = {0u, 0u}, stgph = 0u;
      uint32_t cnt = 0, lcnt = 0;
      for (int it = 0; it < niter; ++it)
        for (int ti = 0; ti < P.ntiles; ++ti) {
          const SqTile& T = P.tile[ti];
          const uint32_t as = cnt & 1u;
          const bool lg = P.tlog != nullptr && blockIdx.x == 0 && it == 1 && lane == 0;
          if (lg) P.tlog[10 * ti + 0] = sq_now();
          mbar_wait(bar_accempty(as), ((cnt >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (lg) P.tlog[10 * ti + 1] = sq_now();
          const uint32_t d_tmem = tmem_base + as * 256u;
          const uint32_t idesc = T.kind == SQ_EMB ? P.idesc_e : (PAIR ? (P.idesc_l & ~(0x1fu << 24)) | ((uint32_t)(256 >> 4) << 24) : P.idesc_l);
          const uint32_t cbase = MODE == 1 ? 4u * (lcnt & 1u) : 0u;
          // this warp's instruction stream paces the MMA-bound stretches: the tile's fields are read once, a k-block reads ONE
          // 32-bit word of the program, and the rare branches (waits, skip stores) sit behind a single test each
          const int nkb = T.nkb, per = T.kind == SQ_EMB ? 4 : 1, xcommit = T.xcommit;
          const int stkb0 = T.st[0].kb, stkb1 = T.st[1].kb;
          const uint32_t wrows128 = (uint32_t)T.wrows * 128u;
          for (int kb = 0; kb < nkb; kb += per) {
            const int nsub = min(per, nkb - kb);
            const uint32_t kw = *reinterpret_cast<const uint32_t*>(&T.k[kb]);   // {ablk, wait, wcol}
            const int wc = (int)((kw >> 8) & 0xffu);
            if (wc) {
              if (wc == SQ_W_H0) { mbar_wait(bar_hready(0), hph[0]); hph[0] ^= 1u; }
              else if (wc == SQ_W_H1) { mbar_wait(bar_hready(1), hph[1]); hph[1] ^= 1u; }
              else if (wc == SQ_W_STG) { mbar_wait(bar_stg, stgph); stgph ^= 1u; }
              else if (wc == SQ_W_C) { mbar_wait(bar_cready(lcnt & 1u), (lcnt >> 1) & 1u); }
            }
            if (MODE == 0 && (kb == stkb0 || kb == stkb1)) {
#pragma unroll
              for (int q = 0; q < 2; ++q)
                if (T.st[q].kb == kb && lane == 0) {
                  tma_store_2d(&P.tmSkip[T.st[q].sel], blocks + (uint32_t)T.st[q].blk0 * SQ_BLK, T.st[q].col0, (int)blockIdx.x * 128);
                  tma_store_2d(&P.tmSkip[T.st[q].sel], blocks + (uint32_t)(T.st[q].blk0 + 1) * SQ_BLK, T.st[q].col0 + 64, (int)blockIdx.x * 128);
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (lg && kb == 0) P.tlog[10 * ti + 2] = sq_now();
            mbar_wait(bar_wfull(stage), phase);
            tc_fence_after();
            if (lg && kb == 0) P.tlog[10 * ti + 3] = sq_now();
            for (int j = 0; j < nsub; ++j) {
              const uint32_t ablk = j == 0 ? (kw & 0xffu) : (uint32_t)T.k[kb + j].ablk;
              const uint64_t adesc = make_sdesc(blocks + (cbase + ablk) * SQ_BLK);
              const uint64_t bdesc = make_sdesc(ring + (uint32_t)stage * WST + (uint32_t)j * wrows128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                sq_umma<PAIR>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb + j > 0 || k > 0) ? 1u : 0u);
            }
            sq_commit<PAIR>(bar_wempty(stage));
            if (xcommit && xcommit == kb + nsub) sq_commit<PAIR>(bar_xfree);
            if (++stage == SQ_STAGES) { stage = 0; phase ^= 1u; }
          }
          if (MODE == 0 && lane == 0) {   // (the stores were issued microseconds ago: these waits do not stall)
            if (T.rdwait == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else if (T.rdwait == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (T.fullwait) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          }
          __syncwarp();   // lane 0's bulk-group waits precede the commit that lets the epilogue overwrite the stored blocks
          sq_commit<PAIR>(bar_accfull(as));
          if (lg) P.tlog[10 * ti + 4] = sq_now();
          if (T.commit_ldone) sq_commit<PAIR>(bar_ldone);
          if (MODE == 1 && T.layer_last) sq_commit<PAIR>(bar_cfree(lcnt & 1u));   // the ctx buffer of this layer may be rewritten
          if (T.layer_last) ++lcnt;
          ++cnt;
        }
    }
  } else if (MODE == 1 && warp >= 2 + NWORK) {
    // ===================== gate pass, 4 producer warps: ctx activations c = SiLU(cx[b] + ct[t]) of global layer g (item, layer) ->
    // c buffer g & 1 (blocks 4 (g & 1) ...).  A warp owns 32 chain rows; 16 lanes cover the 64 columns of a block row (float4
    // each), two rows per instruction, 16 rows per pass.  The cx / ct loads of the pass after the one being converted are always
    // in flight (a cursor over the flat sequence of (item, layer, 64-column block, half)): issued one pass at a time they would
    // cost a DRAM round trip per pass.
    const int pw = warp - 2 - NWORK;
    constexpr bool fp16 = FP16;
    constexpr int PROWS = 128 / SQ_GATE_PW, PH = PROWS / 16;   // rows per producer warp, 16-row passes per block
    int c_it = 0, c_l = 0, c_j = 0, c_h = 0;
    float4 xn[8], tn;
    const int c4 = (lane & 15) * 4;
    auto issue = [&]() {
      if (c_it >= niter) return;
      const int c_item = item_of(c_it);
      const int ctile = c_item / P.nsteps, tl = c_item - ctile * P.nsteps;
      const int irev = P.T - 1 - (P.s0 + tl);
      const int col = P.coff[c_l] + 64 * c_j + c4;
      tn = __ldg(reinterpret_cast<const float4*>(P.ct + (size_t)irev * P.csum + col));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long bb = (long long)ctile * 128 + pw * PROWS + c_h * 16 + 2 * u + (lane >> 4);
        xn[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bb < P.B)   // read once per step: keep it out of L1
          asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(xn[u].x), "=f"(xn[u].y), "=f"(xn[u].z), "=f"(xn[u].w) : "l"(P.cx + bb * P.csum + col));
      }
    };
    issue();
    uint32_t g = 0;
    for (int it = 0; it < niter; ++it)
      for (int l = 0; l < DEN_LAYERS; ++l, ++g) {
        if (g >= 2) mbar_wait_relaxed(bar_cfree(g & 1u), ((g >> 1) - 1u) & 1u);   // the MMAs of layer g - 2 have read this buffer
        const int nb = P.dout[l] >> 6;
        for (int j = 0; j < nb; ++j)
          for (int h = 0; h < PH; ++h) {
            float4 x4[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) x4[u] = xn[u];
            const float4 t4 = tn;
            if (++c_h == PH) { c_h = 0; if (++c_j == (P.dout[c_l] >> 6)) { c_j = 0; if (++c_l == DEN_LAYERS) { c_l = 0; ++c_it; } } }
            issue();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int R = pw * PROWS + h * 16 + 2 * u + (lane >> 4);
              const float a0 = x4[u].x + t4.x, a1 = x4[u].y + t4.y, a2 = x4[u].z + t4.z, a3 = x4[u].w + t4.w;
              const float s0 = sq_silu_fast(a0), s1 = sq_silu_fast(a1), s2 = sq_silu_fast(a2), s3 = sq_silu_fast(a3);
              const uint32_t dst = blocks + sq_chunk((int)(4u * (g & 1u)) + j, R, c4) + (uint32_t)((c4 & 4) << 1);
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(pack2(fp16, s0, s1)), "r"(pack2(fp16, s2, s3)) : "memory");
            }
          }
        fence_proxy_async();   // generic-proxy writes of the tile -> the tensor core's reads
        __syncwarp();
        if (lane == 0) { if (PAIR) mbar_arrive_cluster(lead(bar_cready(g & 1u))); else mbar_arrive(bar_cready(g & 1u)); }
      }
  } else {
    // ===================== 8 worker warps: two per TMEM lane quarter; a thread owns one chain row and half of a tile's features ====
    const int ew = warp - 2, q = warp & 3, sub = ew >> 2, r = q * 32 + lane;
    constexpr int RW = 128 / NWORK;   // rows per warp in the row-major passes
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr bool fp16 = FP16;
    // this warp's part of an operand tile is written (generic proxy): hand it to the async proxy.  all_spaces: also the skip rows
    // this thread stored to global memory since the last such fence (read back by TMA much later)
    auto signal = [&](uint32_t bar, bool all_spaces) {
      tc_fence_before();
      if (all_spaces) fence_proxy_async_all(); else fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    uint32_t cnt = 0;
    if (MODE == 0) {
      const long long brow = (long long)blockIdx.x * 128 + r;
      const bool rok = brow < P.B;
      // ---- before the first step: zh -> X2,X3 ; zl -> Y0,Y1 (what the last layer's epilogue leaves behind for every later step) ----
      {
        const int c4 = lane * 4;
#pragma unroll 1
        for (int r0 = 0; r0 < RW; r0 += 4) {
          float4 zq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const long long bb = (long long)blockIdx.x * 128 + ew * RW + r0 + u;
            zq[u] = bb < P.B ? *reinterpret_cast<const float4*>(P.z + bb * P.nz + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int R = ew * RW + r0 + u;
            const float v[4] = {zq[u].x, zq[u].y, zq[u].z, zq[u].w};
            __half h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) { h[e] = __float2half_rn(v[e]); l[e] = __float2half_rn(v[e] - __half2float(h[e])); }
            const uint32_t o = sq_chunk(2 + (c4 >> 6), R, c4) + (uint32_t)((c4 & 4) << 1);
            const uint32_t ol = sq_chunk(4 + (c4 >> 6), R, c4) + (uint32_t)((c4 & 4) << 1);
            const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
            const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(blocks + o), "r"(*reinterpret_cast<const uint32_t*>(&h01)),
                         "r"(*reinterpret_cast<const uint32_t*>(&h23)) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(blocks + ol), "r"(*reinterpret_cast<const uint32_t*>(&l01)),
                         "r"(*reinterpret_cast<const uint32_t*>(&l23)) : "memory");
            P.zT[((size_t)blockIdx.x * (P.nz >> 2) + lane) * 128 + R] = zq[u];   // (strided, once per launch)
            if (!fp16) {   // bf16 operands: the first layer's z columns are bf16(z) (Y2,Y3), not the fp16 split's high half
              const uint32_t ob = sq_chunk(6 + (c4 >> 6), R, c4) + (uint32_t)((c4 & 4) << 1);
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(blocks + ob), "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])) : "memory");
            }
          }
        }
        signal(bar_hready(1), false);
      }
      for (int st = 0; st < P.nsteps; ++st) {
        const int sg = P.s0 + st;   // sampler step (execution order)
        const float* cf = P.coef + 8 * (size_t)sg;
        DenEpi d{};
        d.z = P.z; d.eps_out = nullptr; d.noise = nullptr;
        d.nz = P.nz; d.residual = P.residual; d.use_philox = 0;   // Philox normals are drawn by den_seq_noise_kernel (same bits)
        d.c_pred = __ldg(cf); d.c_eps = __ldg(cf + 1); d.c_zt = __ldg(cf + 2); d.c_x = __ldg(cf + 3); d.c_std = __ldg(cf + 4);
        d.last = __ldg(cf + 5) != 0.f;
        d.seed = P.seed; d.chain0 = P.chain0; d.step = (unsigned long long)sg; d.seed_ptr = nullptr;
        const uint4* Gt = P.G + ((size_t)st * tiles_b + blockIdx.x) * (size_t)(P.csum >> 2) * 128 + r;
        float4* zt_ = P.zT + (size_t)blockIdx.x * (P.nz >> 2) * 128 + r;
        const float4* nt_ = P.noiseT ? P.noiseT + ((size_t)st * tiles_b + blockIdx.x) * (size_t)(P.nz >> 2) * 128 + r : nullptr;
        const bool write_z = st == P.nsteps - 1;   // row-major z leaves the kernel once, after the window's last step
#pragma unroll 1
        for (int ti = 0; ti < P.ntiles; ++ti, ++cnt) {
          const SqTile& T = P.tile[ti];
          const uint32_t as = cnt & 1u;
          const uint32_t t_acc = t_lane + as * 256u;
          if (T.kind == SQ_EMB) {
            mbar_wait(bar_accfull(as), (cnt >> 1) & 1u);
            tc_fence_after();
            // phases of this row's 64 / NWQ frequencies: sin -> X0, cos -> X1 (diffusion_net.py:497-499)
            constexpr int PE = 64 / NWQ;
            uint32_t v[32];
            if (PE == 32) tmem_ld32(t_acc + (uint32_t)(sub * PE), v); else tmem_ld16(t_acc + (uint32_t)(sub * PE), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < PE / 8; ++j) {
              uint32_t ws[4], wc[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // sin(2 pi p) has period 1 in p: exact reduction to [-pi, pi], where the SFU forms are accurate to 2^-21 absolute
                const float p0 = __uint_as_float(v[j * 8 + 2 * e]), p1 = __uint_as_float(v[j * 8 + 2 * e + 1]);
                const float a0 = 6.283185307179586f * (p0 - rintf(p0)), a1 = 6.283185307179586f * (p1 - rintf(p1));
                const float s0 = __sinf(a0), c0 = __cosf(a0), s1 = __sinf(a1), c1 = __cosf(a1);
                ws[e] = pack2(fp16, s0, s1);
                wc[e] = pack2(fp16, c0, c1);
              }
              const int col = sub * PE + j * 8;
              sq_st16(blocks + sq_chunk(0, r, col), ws[0], ws[1], ws[2], ws[3]);
              sq_st16(blocks + sq_chunk(1, r, col), wc[0], wc[1], wc[2], wc[3]);
            }
          } else {
            // the gate / hyper-bias words of this row's 64 features of the tile: requested before the accumulator is waited for
            uint4 g[FPW / 4];
            const uint4* gp = Gt + (size_t)((T.goff >> 2) + sub * (FPW / 4)) * 128;
#pragma unroll
            for (int i = 0; i < FPW / 4; ++i) g[i] = __ldg(gp + i * 128);
            // last layer: this row's z / noise quads, one 16-feature chunk at a time (row-strided 16-byte reads: L2 round trips that
            // must not sit between the accumulator and the update)
            float4 zq[2][4], nq[2][4];
            auto ld_zn = [&](int c, float4 (&zz)[4], float4 (&nn)[4]) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                zz[j] = zt_[(sub * (FPW / 4) + c * 4 + j) * 128];
                nn[j] = nt_ ? __ldg(nt_ + (sub * (FPW / 4) + c * 4 + j) * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            };
            if (T.kind == SQ_FINAL) ld_zn(0, zq[0], nq[0]);
            if (P.tlog && blockIdx.x == 0 && st == 1 && threadIdx.x == 64) P.tlog[10 * ti + 5] = sq_now();
            mbar_wait(bar_accfull(as), (cnt >> 1) & 1u);
            if (T.defer) mbar_wait(bar_xfree, (uint32_t)st & 1u);   // the output blocks are inputs of the next tile's first k-blocks
            tc_fence_after();
            if (P.tlog && blockIdx.x == 0 && st == 1 && threadIdx.x == 64) P.tlog[10 * ti + 6] = sq_now();
            const bool fin = T.kind == SQ_FINAL;
            const bool noisy = fin && !d.last && d.c_std != 0.f && nt_ != nullptr;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
              uint32_t vm[16], vs[16];
              tmem_ld16(t_acc + (uint32_t)(sub * FPW + c * 16), vm);
              tmem_ld16(t_acc + 128u + (uint32_t)(sub * FPW + c * 16), vs);
              if (fin && c < NCH - 1) ld_zn(c + 1, zq[(c + 1) & 1], nq[(c + 1) & 1]);   // the next chunk's z / noise quads: a chunk ahead
              tmem_ld_wait();
              float o[16];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const uint32_t gw[4] = {g[c * 4 + i4].x, g[c * 4 + i4].y, g[c * 4 + i4].z, g[c * 4 + i4].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  // gw = (sigmoid(gate), hyper-bias + bs + b sigmoid(gate)):  (main + b) sig + hb + skip + bs = main sig + skip + gw.y
                  const float2 sh = __half22float2(*reinterpret_cast<const __half2*>(&gw[e]));
                  const int i = i4 * 4 + e;
                  o[i] = fmaf(__uint_as_float(vm[i]), sh.x, __uint_as_float(vs[i]) + sh.y);
                }
              }
              const int f = sub * FPW + c * 16;   // feature inside the tile
              if (fin) {
                // eps = z + out, reverse update of z; the new z also leaves as the fp16 split (zh -> X2,X3 ; zl -> Y0,Y1) for the
                // next step's embedding GEMM (and as bf16 -> Y2,Y3, the first layer's z operand, in bf16 mode)
                uint32_t wh[8], wl[8], wb[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 z4 = zq[c & 1][j], n4 = nq[c & 1][j];
                  const float zt[4] = {z4.x, z4.y, z4.z, z4.w}, nrm[4] = {n4.x, n4.y, n4.z, n4.w};
                  float res[4];
                  den_final_math(d, zt, nrm, noisy, o + 4 * j, res);
                  zt_[((f >> 2) + j) * 128] = make_float4(res[0], res[1], res[2], res[3]);
                  if (write_z && rok) *reinterpret_cast<float4*>(P.z + brow * P.nz + f + 4 * j) = make_float4(res[0], res[1], res[2], res[3]);
                  __half h[4], l[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) { h[e] = __float2half_rn(res[e]); l[e] = __float2half_rn(res[e] - __half2float(h[e])); }
                  const __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
                  const __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
                  wh[2 * j] = *reinterpret_cast<const uint32_t*>(&h01); wh[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                  wl[2 * j] = *reinterpret_cast<const uint32_t*>(&l01); wl[2 * j + 1] = *reinterpret_cast<const uint32_t*>(&l23);
                  wb[2 * j] = pack_bf16x2(res[0], res[1]); wb[2 * j + 1] = pack_bf16x2(res[2], res[3]);
                }
                sq_st16(blocks + sq_chunk(2 + (f >> 6), r, f), wh[0], wh[1], wh[2], wh[3]);
                sq_st16(blocks + sq_chunk(2 + (f >> 6), r, f + 8), wh[4], wh[5], wh[6], wh[7]);
                sq_st16(blocks + sq_chunk(4 + (f >> 6), r, f), wl[0], wl[1], wl[2], wl[3]);
                sq_st16(blocks + sq_chunk(4 + (f >> 6), r, f + 8), wl[4], wl[5], wl[6], wl[7]);
                if (!fp16) {
                  sq_st16(blocks + sq_chunk(6 + (f >> 6), r, f), wb[0], wb[1], wb[2], wb[3]);
                  sq_st16(blocks + sq_chunk(6 + (f >> 6), r, f + 8), wb[4], wb[5], wb[6], wb[7]);
                }
              } else {
                uint32_t w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = pack2(fp16, fmaxf(o[2 * j], 0.01f * o[2 * j]), fmaxf(o[2 * j + 1], 0.01f * o[2 * j + 1]));
                const int blk = T.oblk + (f >> 6);
                sq_st16(blocks + sq_chunk(blk, r, f), w[0], w[1], w[2], w[3]);
                sq_st16(blocks + sq_chunk(blk, r, f + 8), w[4], w[5], w[6], w[7]);
                if (T.skipsel) {   // U-net skip of layers 0 / 1: through L2, back by TMA at the matching out layer
                  uint16_t* sk = reinterpret_cast<uint16_t*>(P.skip[T.skipsel - 1]) + brow * P.skip_ld[T.skipsel - 1] + T.ocol0 + f;
                  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(sk), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                               "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_accempty(as));
          signal(bar_hready(as), T.gfence != 0);
          if (P.tlog && blockIdx.x == 0 && st == 1) {
            if (threadIdx.x == 64) P.tlog[10 * ti + 7] = sq_now();
            if (lane == 0) atomicMax(P.tlog + 10 * ti + 9, sq_now());
          }
        }
      }
    } else {
      // ---------------------------------------------- gate pass ----------------------------------------------------------------
      // (the ctx operand tiles are formed by the producer warps above; these warps turn accumulators into gate words)
      for (int it = 0; it < niter; ++it) {
        const int item = item_of(it);
        // chain tile major: the steps of one chain tile run at about the same time on neighbouring CTAs and share its cx rows in L2
        const int ctile = item / P.nsteps, tl = item - ctile * P.nsteps;
        uint4* Gt = P.G + ((size_t)tl * tiles_b + ctile) * (size_t)(P.csum >> 2) * 128 + r;
        int ti = 0;
        for (int l = 0; l < DEN_LAYERS; ++l) {
          const int ntl = P.dout[l] >> 7;
          for (int n = 0; n < ntl; ++n, ++ti, ++cnt) {
            const SqTile& T = P.tile[ti];
            const uint32_t as = cnt & 1u;
            const uint32_t t_acc = t_lane + as * 256u;
            mbar_wait_relaxed(bar_accfull(as), (cnt >> 1) & 1u);
            tc_fence_after();
            // (same address in every lane; the cx stream bypasses L1, so these 17 KB of biases stay in it)
            const float* bm = P.bias3[l] + T.ocol0 + sub * FPW;
            const float* bs = bm + P.dout[l];
            const float* bg = bs + P.dout[l];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t vg[16], vh[16];
              tmem_ld16(t_acc + (uint32_t)(sub * FPW + c * 16), vg);
              tmem_ld16(t_acc + 128u + (uint32_t)(sub * FPW + c * 16), vh);
              tmem_ld_wait();
              uint32_t w[16];
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(bg + c * 16) + i4);
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(bs + c * 16) + i4);
                const float4 m4 = __ldg(reinterpret_cast<const float4*>(bm + c * 16) + i4);
                const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const int i = i4 * 4 + e;
                  const float gate = __uint_as_float(vg[i]) + gg[e];
                  const float sig = __fdividef(1.f, 1.f + sq_ex2(-1.4426950408889634f * gate));
                  // everything of out = (main + b) sig + hb + skip + bs that does not depend on h:  (sig, hb + bs + b sig)
                  w[i] = pack_f16x2(sig, fmaf(mm[e], sig, __uint_as_float(vh[i]) + ss[e]));
                }
              }
              uint4* gp = Gt + (size_t)((T.goff >> 2) + sub * (FPW / 4) + c * 4) * 128;
#pragma unroll
              for (int i4 = 0; i4 < 4; ++i4) gp[i4 * 128] = make_uint4(w[4 * i4], w[4 * i4 + 1], w[4 * i4 + 2], w[4 * i4 + 3]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (PAIR) mbar_arrive_cluster(lead(bar_accempty(as))); else mbar_arrive(bar_accempty(as)); }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // neither CTA may retire while the pair's MMAs / multicast arrives can still touch it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// The window's normals in the step pass's tile-transposed layout [step][chain tile][nz/4][128 rows]: Philox draws with the key of
// the other schedules' last epilogue (seed, GLOBAL chain index, step, quad) -- same bits, taken off the chain of dependent work --
// or the caller's injected noise [T-1][B][nz]
__global__ void __launch_bounds__(256) den_seq_noise_kernel(float4* __restrict__ out, const float* __restrict__ src, int B, int Bpad,
                                                            int nz, int T, int s0, int ns, unsigned long long seed,
                                                            unsigned long long chain0) {
  // grid (Bpad * nz / 4 / 256, ns): 32-bit index arithmetic (64-bit divisions cost more than the Philox rounds)
  const uint32_t q4 = (uint32_t)nz >> 2;
  const uint32_t j = blockIdx.x * 256u + threadIdx.x;   // (chain tile, quad, row) of step blockIdx.y
  if (j >= (uint32_t)Bpad * q4) return;
  const int r = (int)(j & 127u);
  const uint32_t t1 = j >> 7;
  const int ctile = (int)(t1 / q4), quad = (int)(t1 - (uint32_t)ctile * q4), st = (int)blockIdx.y;
  const size_t i = (size_t)st * Bpad * q4 + j;
  const int b = ctile * 128 + r, sg = s0 + st;
  (void)ns;
  float n[4] = {0.f, 0.f, 0.f, 0.f};
  if (b < B && sg < T - 1) {   // the last step adds no noise (and injected noise has T - 1 slabs)
    if (src) {
      const float4 v = *reinterpret_cast<const float4*>(src + ((size_t)sg * B + b) * nz + 4 * quad);
      n[0] = v.x; n[1] = v.y; n[2] = v.z; n[3] = v.w;
    } else {
      philox_normal4(seed, chain0 + (unsigned long long)b, (unsigned long long)sg, (uint32_t)quad, n);
    }
  }
  out[i] = make_float4(n[0], n[1], n[2], n[3]);
}

// ---- packing ---------------------------------------------------------------------------------------------------------------
// rows of N tile t (256 rows): [first 128 features of the tile | second] = [W_a rows | W_b rows]; K-major [2*dout][K]
template <typename T>
__global__ void pack_seq_pair(const float* __restrict__ Wa, const float* __restrict__ Wb, int dout, int K, T* __restrict__ dst,
                              const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2ll * dout * K) return;
  const int row = (int)(i / K), k = (int)(i - (long long)row * K);
  const int tile = row >> 8, rr = row & 255, qq = rr >> 7, n = tile * 128 + (rr & 127);
  dst[i] = T((qq ? Wb : Wa)[(size_t)n * K + k]);
}
// embedding: rows j < nz/2, columns [Bh | Bh | Bl] (each nz wide), fp16 split of Bproj[k][j]
__global__ void pack_seq_emb(const float* __restrict__ Bp, int nz, __half* __restrict__ dst, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const int half = nz >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half * nz) return;
  const int j = i / nz, k = i - j * nz;
  const float v = Bp[(size_t)k * half + j];
  const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
  __half* row = dst + (size_t)j * 3 * nz;
  row[k] = h; row[nz + k] = h; row[2 * nz + k] = l;
}

// K-major tf32-rounded copy [csum][nxemb] of the xemb half of every layer's ctx Linear (Wc[:, ntemb:])
__global__ void pack_seq_wcx(const float* __restrict__ Wc, int dout, int ntemb, int nxemb, float* __restrict__ dst, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)dout * nxemb) return;
  const int n = (int)(i / nxemb), k = (int)(i - (long long)n * nxemb);
  dst[i] = tf32_rna(Wc[(size_t)n * (ntemb + nxemb) + ntemb + k]);
}
// the ctx Linear acts on SiLU(xemb) (diffusion_net.py:426: Sequential(SiLU, Linear, SiLU)): operand rows tf32(SiLU(xemb))
__device__ __forceinline__ float sq_silu(float v) { return v / (1.f + expf(-v)); }
__global__ void __launch_bounds__(256) sq_silu_tf32(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n4) return;
  const float4 v = src[i];
  dst[i] = make_float4(tf32_rna(sq_silu(v.x)), tf32_rna(sq_silu(v.y)), tf32_rna(sq_silu(v.z)), tf32_rna(sq_silu(v.w)));
}

struct DenSeqPack {
  float* Wcx = nullptr;   // [csum][nxemb] tf32 (own allocation)
  void* slab = nullptr;
  void* Wms[DEN_LAYERS];
  void* Wgh[DEN_LAYERS];
  __half* Wemb = nullptr;
  CUtensorMap tmMs[DEN_LAYERS], tmGh[DEN_LAYERS], tmEmb;
};

void den_seq_free(DenSeqPack* p) {
  if (!p) return;
  if (p->slab) cudaFree(p->slab);
  if (p->Wcx) cudaFree(p->Wcx);
  delete p;
}

bool den_seq_supported(const DenPack* d) {
  const char* e = getenv("DAMC_DEN_SEQ");   // read per call: the tests compare the schedules in one process
  return !(e && e[0] == '0') && den_seq_shape_ok(d) && tc_available();
}

bool den_seq_shape_ok(const DenPack* d) {
  if (d->nz != 128 || d->din[0] != 2 * d->nz || d->dout[6] != d->nz) return false;
  for (int i = 0; i < DEN_LAYERS; ++i)
    if ((d->dout[i] != 128 && d->dout[i] != 256) || d->csum > SQ_MAXCSUM) return false;
  // U-net wiring this kernel's buffer plan is written for (diffusion_net.py:505-528): 128 -> 256 -> 256 | 256 | (256+256) -> 256,
  // (256+256) -> 128, (128+128) -> nz
  return d->dout[0] == 128 && d->dout[1] == 256 && d->dout[2] == 256 && d->dout[3] == 256 && d->dout[4] == 256 && d->dout[5] == 128 &&
         d->din[1] == 128 && d->din[2] == 256 && d->din[3] == 256 && d->din[4] == 512 && d->din[5] == 512 && d->din[6] == 256;
}

int den_seq_refill(const DenPack* d, int precision, cudaStream_t s, const int* dirty) {
  const DenTcPack* t = d->tc[precision];
  if (!t || !t->seq) return DAMC_OK;
  DenSeqPack* p = t->seq;
  const damc_denoiser_desc* h = &d->src;
  for (int i = 0; i < DEN_LAYERS; ++i) {
    const int di = d->din[i], dn = d->dout[i];
    const int b1 = (int)((2ll * dn * di + 255) / 256), b2 = (int)((2ll * dn * dn + 255) / 256);
    if (precision == DAMC_PREC_FP16) {
      pack_seq_pair<__half><<<b1, 256, 0, s>>>(h->W[i], h->Ws[i], dn, di, (__half*)p->Wms[i], dirty);
      pack_seq_pair<__half><<<b2, 256, 0, s>>>(h->Wg[i], h->Wb[i], dn, dn, (__half*)p->Wgh[i], dirty);
    } else {
      pack_seq_pair<__nv_bfloat16><<<b1, 256, 0, s>>>(h->W[i], h->Ws[i], dn, di, (__nv_bfloat16*)p->Wms[i], dirty);
      pack_seq_pair<__nv_bfloat16><<<b2, 256, 0, s>>>(h->Wg[i], h->Wb[i], dn, dn, (__nv_bfloat16*)p->Wgh[i], dirty);
    }
  }
  pack_seq_emb<<<ceil_div((d->nz / 2) * d->nz, 256), 256, 0, s>>>(h->Bproj, d->nz, p->Wemb, dirty);
  for (int i = 0; i < DEN_LAYERS; ++i)
    pack_seq_wcx<<<(unsigned)(((long long)d->dout[i] * d->nxemb + 255) / 256), 256, 0, s>>>(h->Wc[i], d->dout[i], d->ntemb, d->nxemb,
                                                                                              p->Wcx + (size_t)d->coff[i] * d->nxemb, dirty);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

static int den_seq_ensure(const DenPack* d, int precision, cudaStream_t s) {
  DAMC_TRY(den_tc_ensure(d, precision, s));
  DenTcPack* t = d->tc[precision];
  if (t->seq) return DAMC_OK;
  DenSeqPack* p = new DenSeqPack();
  size_t bytes = 0;
  for (int i = 0; i < DEN_LAYERS; ++i) bytes += align_up(4 * (size_t)d->dout[i] * d->din[i], 256) + align_up(4 * (size_t)d->dout[i] * d->dout[i], 256);
  const size_t ebytes = align_up((size_t)(d->nz / 2) * 3 * d->nz * 2, 256);
  if (cudaMalloc(&p->slab, bytes + ebytes) != cudaSuccess) { delete p; DAMC_FAIL(DAMC_ERR_CUDA, "denoiser (hoisted schedule): cudaMalloc(%zu) failed", bytes + ebytes); }
  char* q = (char*)p->slab;
  const int fp16 = precision == DAMC_PREC_FP16 ? 1 : 0;
  int r = DAMC_OK;
  for (int i = 0; i < DEN_LAYERS && r == DAMC_OK; ++i) {
    p->Wms[i] = q; q += align_up(4 * (size_t)d->dout[i] * d->din[i], 256);
    p->Wgh[i] = q; q += align_up(4 * (size_t)d->dout[i] * d->dout[i], 256);
    r = tc_encode_2d(&p->tmMs[i], fp16, p->Wms[i], d->din[i], 2 * d->dout[i], SQ_WBOX);
    if (r == DAMC_OK) r = tc_encode_2d(&p->tmGh[i], fp16, p->Wgh[i], d->dout[i], 2 * d->dout[i], SQ_WBOX);
  }
  p->Wemb = (__half*)q;
  if (r == DAMC_OK) r = tc_encode_2d(&p->tmEmb, 1, p->Wemb, 3 * d->nz, d->nz / 2, d->nz / 2);
  if (r == DAMC_OK && cudaMalloc(&p->Wcx, sizeof(float) * (size_t)d->csum * d->nxemb) != cudaSuccess) { r = DAMC_ERR_CUDA; set_error("denoiser (hoisted schedule): cudaMalloc failed"); }
  if (r != DAMC_OK) { den_seq_free(p); return r; }
  t->seq = p;
  return den_seq_refill(d, precision, s, nullptr);
}

// cx = SiLU(xemb) Wc_x^T + bc for all layers (diffusion_net.py:426-433, the xemb half of the ctx Linears) as ONE tcgen05 GEMM in TF32
// (operands rounded, not truncated) instead of the CUDA-core hoist kernel: 47 GFLOP at 16 384 chains, 3 ms -> 0.1 ms
bool den_seq_hoist_usable(const DenPack* d, int precision, int B) {
  (void)B;   // every batch size: a chain's result must not depend on how the batch is sharded
  return is_tc_precision(precision) && d->nxemb % 32 == 0 && d->csum % 16 == 0 && den_seq_supported(d);
}
int den_seq_hoist(const DenPack* d, int precision, const DenWs& w, const float* xemb, int B, cudaStream_t s) {
  DAMC_TRY(den_seq_ensure(d, precision, s));
  const DenSeqPack* p = d->tc[precision]->seq;
  if (!w.xr) DAMC_FAIL(DAMC_ERR_WORKSPACE, "denoiser (hoisted schedule): workspace was carved without the xemb copy");
  if ((uintptr_t)xemb & 15) DAMC_FAIL(DAMC_ERR_INVALID, "denoiser: xemb must be 16-byte aligned");
  const size_t n4 = (size_t)B * d->nxemb / 4;
  sq_silu_tf32<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4*>(xemb), reinterpret_cast<float4*>(w.xr), n4);
  DAMC_CUDA(cudaGetLastError());
  GemmPlan g{};
  g.A = w.xr; g.B = B; g.Hm = 1; g.Wm = 1; g.Cs = d->nxemb;
  g.ntaps = 1; g.taps[0] = Tap{0, 0, 0, 0};
  g.N = g.Np = d->csum; g.ksplit = 1;
  g.Wtc = p->Wcx;
  g.epi.kind = EPI_STORE_F32_BIAS;
  g.epi.bias = d->bc;
  g.epi.out = w.cx;
  g.epi.nz_out = d->csum;
  count_launch(2);
  return launch_gemm_tc(g, DAMC_PREC_TF32, s);
}

// steps per window: the gate pass writes G for this many steps, then the step pass consumes them
int den_seq_window(int B, int T, int csum) {
  const size_t per_step = (size_t)align_up(B, 128) * csum * 4;
  const size_t budget = (size_t)4 << 30;   // of gate words (fewer, longer launches: the gate pass wastes its last partial round of items)
  const char* e = getenv("DAMC_DEN_SEQ_WINDOW");   // test switch: force short windows (read per call, by den_ws and the run alike)
  if (e && atoi(e) > 0) return std::min(T, atoi(e));
  return (int)std::max<size_t>(1, std::min<size_t>((size_t)T, budget / per_step));
}

static void sq_k(SqTile& t, int i, int ablk, int wcol, int wait) { t.k[i].ablk = (unsigned char)ablk; t.k[i].wcol = (unsigned short)wcol; t.k[i].wait = (unsigned char)wait; }

int den_seq_run(const DenPack* d, int precision, const DenWs& w, float* z, int B, int T, int nsteps, const float* host_coef,
                const float* noise, int use_philox, uint64_t seed, uint64_t chain0, cudaStream_t s) {
  DAMC_TRY(den_seq_ensure(d, precision, s));
  const DenSeqPack* p = d->tc[precision]->seq;
  if (!w.G || !w.skip[0] || !w.skip[1] || !w.nbuf || !w.zT) DAMC_FAIL(DAMC_ERR_WORKSPACE, "denoiser (hoisted schedule): workspace was carved without its buffers");
  const int fp16 = precision == DAMC_PREC_FP16 ? 1 : 0;
  const int Bpad = (int)align_up(B, 128);
  DAMC_CUDA(cudaMemcpyAsync(w.coef, host_coef, sizeof(float) * 8 * (size_t)nsteps, cudaMemcpyHostToDevice, s));

  SqParams S{};   // step pass
  SqParams Gp{};  // gate pass
  for (SqParams* q : {&S, &Gp}) {
    q->B = B; q->Bpad = Bpad; q->nz = d->nz; q->csum = d->csum; q->T = T;
    q->fp16 = fp16; q->residual = d->residual; q->use_philox = use_philox;
    const uint32_t opfmt = fp16 ? 0u : 1u;
    q->idesc_l = (1u << 4) | (opfmt << 7) | (opfmt << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    q->idesc_e = (1u << 4) | ((uint32_t)((d->nz / 2) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // fp16 x fp16
    q->z = z; q->zT = reinterpret_cast<float4*>(w.zT); q->seed = seed; q->chain0 = chain0; q->coef = w.coef; q->cx = w.cx; q->ct = w.ct;
    q->G = reinterpret_cast<uint4*>(w.G);
    for (int i = 0; i < 2; ++i) { q->skip[i] = w.skip[i]; q->skip_ld[i] = d->dout[i]; }
    for (int i = 0; i < DEN_LAYERS; ++i) { q->bias3[i] = d->bias3[i]; q->dout[i] = d->dout[i]; q->coff[i] = d->coff[i]; }
  }
  for (int i = 0; i < DEN_LAYERS; ++i) { S.tmW[i] = p->tmMs[i]; Gp.tmW[i] = p->tmGh[i]; }
  S.tmW[SQ_EMB_MAP] = p->tmEmb; Gp.tmW[SQ_EMB_MAP] = p->tmEmb;
  DAMC_TRY(tc_encode_2d(&S.tmSkip[0], fp16, w.skip[0], d->dout[0], Bpad, 128));
  DAMC_TRY(tc_encode_2d(&S.tmSkip[1], fp16, w.skip[1], d->dout[1], Bpad, 128));

  // ---- step-pass program (buffer plan in the header comment; X = blocks 0-3, Y = 4-7) ----
  const char* ets = getenv("DAMC_SQ_TMASTORE");   // experiment switch: 0 = the epilogue threads store the skip rows themselves
  const bool tma_store = !(ets && ets[0] == '0');
  for (SqTile& t : S.tile) t.st[0].kb = t.st[1].kb = 255;
  for (SqTile& t : Gp.tile) t.st[0].kb = t.st[1].kb = 255;
  {
    int n = 0;
    auto layer_tile = [&](int l, int nt, int oblk) -> SqTile& {
      SqTile& t = S.tile[n++];
      t.kind = l == DEN_LAYERS - 1 ? SQ_FINAL : SQ_LAYER; t.layer = (unsigned char)l; t.wmap = (unsigned char)l;
      t.wrow0 = (unsigned short)(nt * 256); t.wrows = 256; t.oblk = (unsigned char)oblk; t.ocol0 = (unsigned short)(nt * 128);
      t.goff = (unsigned short)(d->coff[l] + nt * 128);
      return t;
    };
    {   // embedding: [zh | zl | zh] x [Bh | Bh | Bl]
      SqTile& t = S.tile[n++];
      t.kind = SQ_EMB; t.wmap = SQ_EMB_MAP; t.wrow0 = 0; t.wrows = (unsigned short)(d->nz / 2); t.nkb = 6;
      sq_k(t, 0, 2, 0, SQ_W_H1); sq_k(t, 1, 3, 64, 0); sq_k(t, 2, 4, 128, 0); sq_k(t, 3, 5, 192, 0); sq_k(t, 4, 2, 256, 0); sq_k(t, 5, 3, 320, 0);
    }
    {   // L0: X0-3 -> Y0,Y1 (+ skip 0)
      SqTile& t = layer_tile(0, 0, 4);
      t.nkb = 4; t.skipsel = tma_store ? 0 : 1;
      for (int j = 0; j < 4; ++j) sq_k(t, j, (!fp16 && j >= 2) ? 4 + j : j, 64 * j, j == 0 ? SQ_W_H0 : 0);   // bf16: z columns sit in Y2,Y3
    }
    for (int nt = 0; nt < 2; ++nt) {   // L1: Y0,Y1 -> X (+ skip 1)
      SqTile& t = layer_tile(1, nt, 2 * nt);
      t.nkb = 2; t.skipsel = tma_store ? 0 : 2;
      for (int j = 0; j < 2; ++j) sq_k(t, j, 4 + j, 64 * j, (nt == 0 && j == 0) ? SQ_W_H1 : 0);
      if (tma_store && nt == 0) t.st[0] = SqSt{0, 0, 4, 0};   // L0's output (Y0,Y1) is complete once hready[1] has been waited for
    }
    for (int nt = 0; nt < 2; ++nt) {   // L2: X -> Y (stays resident: it is also the skip of L4)
      SqTile& t = layer_tile(2, nt, 4 + 2 * nt);
      t.nkb = 4; t.gfence = nt == 0 && !tma_store;   // (thread-stored skip rows: one all-spaces proxy fence, long after the stores)
      if (tma_store && nt == 0) { t.st[0] = SqSt{0, 1, 0, 0}; t.st[1] = SqSt{2, 1, 2, 128}; t.rdwait = 3; }   // L1's halves; Y0,Y1 are rewritten next
      for (int j = 0; j < 4; ++j) sq_k(t, j, j, 64 * j, nt == 0 ? (j == 0 ? SQ_W_H0 : j == 2 ? SQ_W_H1 : 0) : 0);
    }
    for (int nt = 0; nt < 2; ++nt) {   // L3: Y -> X
      SqTile& t = layer_tile(3, nt, 2 * nt);
      t.nkb = 4; t.rdwait = (tma_store && nt == 0) ? 1 : 0;   // X is rewritten by this layer's epilogues
      for (int j = 0; j < 4; ++j) sq_k(t, j, 4 + j, 64 * j, nt == 0 ? (j == 0 ? SQ_W_H0 : j == 2 ? SQ_W_H1 : 0) : 0);
    }
    for (int nt = 0; nt < 2; ++nt) {   // L4: [X | Y] -> X, tile 0's output once tile 1 has finished reading X
      SqTile& t = layer_tile(4, nt, 2 * nt);
      t.nkb = 8; t.defer = nt == 0; t.commit_ldone = nt == 1;
      if (nt == 0) {   // Y (long ready) first, X as L3's epilogues deliver it
        for (int j = 0; j < 4; ++j) sq_k(t, j, 4 + j, 256 + 64 * j, 0);
        for (int j = 0; j < 4; ++j) sq_k(t, 4 + j, j, 64 * j, j == 0 ? SQ_W_H0 : j == 2 ? SQ_W_H1 : 0);
      } else {         // X first: once these four k-blocks are done, tile 0's epilogue may overwrite X0,X1 while the Y part runs
        for (int j = 0; j < 4; ++j) sq_k(t, j, j, 64 * j, 0);
        for (int j = 0; j < 4; ++j) sq_k(t, 4 + j, 4 + j, 256 + 64 * j, 0);
        t.xcommit = 4; t.fullwait = tma_store;   // the skip rows are in L2 before `ldone` lets the TMA loads fetch them back
      }
    }
    {   // L5: [X | skip 1 staged in Y] -> Y0,Y1
      SqTile& t = layer_tile(5, 0, 4);
      t.nkb = 8; t.commit_ldone = 1; t.stage_nblk = 4; t.stage_blk0 = 4; t.stage_map = 1; t.stage_col0 = 0;
      for (int j = 0; j < 4; ++j) sq_k(t, j, 4 + j, 256 + 64 * j, j == 0 ? SQ_W_STG : 0);
      for (int j = 0; j < 4; ++j) sq_k(t, 4 + j, j, 64 * j, j == 0 ? SQ_W_H0 : j == 2 ? SQ_W_H1 : 0);
    }
    {   // L6: [Y0,Y1 | skip 0 staged in X0,X1] -> z
      SqTile& t = layer_tile(6, 0, 0);
      t.nkb = 4; t.stage_nblk = 2; t.stage_blk0 = 0; t.stage_map = 0; t.stage_col0 = 0;
      sq_k(t, 0, 0, 128, SQ_W_STG); sq_k(t, 1, 1, 192, 0); sq_k(t, 2, 4, 0, SQ_W_H0); sq_k(t, 3, 5, 64, 0);
    }
    S.ntiles = n;
  }
  // ---- gate-pass program: per layer dout/128 tiles over the ctx activations in c buffer (layer counter & 1) ----
  {
    int n = 0;
    for (int l = 0; l < DEN_LAYERS; ++l) {
      const int ntl = d->dout[l] / 128;
      for (int nt = 0; nt < ntl; ++nt) {
        SqTile& t = Gp.tile[n++];
        t.kind = SQ_GATE; t.layer = (unsigned char)l; t.wmap = (unsigned char)l; t.wrow0 = (unsigned short)(nt * 256); t.wrows = 256;
        t.ocol0 = (unsigned short)(nt * 128); t.goff = (unsigned short)(d->coff[l] + nt * 128); t.nkb = (unsigned char)(d->dout[l] / 64);
        for (int j = 0; j < t.nkb; ++j) sq_k(t, j, j, 64 * j, (nt == 0 && j == 0) ? SQ_W_C : 0);
        t.layer_last = nt == ntl - 1;
      }
    }
    Gp.ntiles = n;
  }

  const size_t smem = (size_t)SQ_NBLK * SQ_BLK + (size_t)SQ_STAGES_STEP * SQ_WSTAGE + 8 * (2 * SQ_STAGES_STEP + 13) + 16 + 1024;
  const size_t smem_g = (size_t)SQ_NBLK * SQ_BLK + (size_t)SQ_STAGES_GATE * SQ_WSTAGE + 8 * (4 * SQ_STAGES_GATE + 13) + 16 + 1024;
  int dev = 0, sms = 148;
  DAMC_CUDA(cudaGetDevice(&dev));
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const char* epair = getenv("DAMC_SQ_PAIR");   // experiment switch: 0 = the gate pass without CTA pairs
  const bool pair = !(epair && epair[0] == '0') && sms >= 2;
  auto* const step_kernel = fp16 ? den_seq_kernel<0, true, false> : den_seq_kernel<0, false, false>;
  auto* const gate_kernel = pair ? (fp16 ? den_seq_kernel<1, true, true> : den_seq_kernel<1, false, true>)
                                 : (fp16 ? den_seq_kernel<1, true, false> : den_seq_kernel<1, false, false>);
  DAMC_CUDA(cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DAMC_CUDA(cudaFuncSetAttribute(gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
  static unsigned long long* tlog = nullptr;
  const bool dbg = getenv("DAMC_SQ_DBG") != nullptr;
  if (dbg && !tlog) { cudaMalloc(&tlog, 256 * 8); }
  if (dbg) cudaMemsetAsync(tlog, 0, 256 * 8, s);
  const bool dbg_gate = dbg && getenv("DAMC_SQ_DBG")[0] == '2';   // 2: stamps of the gate pass's MMA thread instead
  S.tlog = (dbg && !dbg_gate) ? tlog : nullptr;
  Gp.tlog = dbg_gate ? tlog : nullptr;
  const int win = den_seq_window(B, T, d->csum);
  for (int s0 = 0; s0 < nsteps; s0 += win) {
    const int ns = std::min(win, nsteps - s0);
    Gp.s0 = S.s0 = s0; Gp.nsteps = S.nsteps = ns;
    const int items = ns * (Bpad / 128);
    bool noisy = false;
    for (int k = s0; k < s0 + ns; ++k) noisy = noisy || (host_coef[8 * (size_t)k + 4] != 0.f && host_coef[8 * (size_t)k + 5] == 0.f);
    S.noiseT = nullptr;
    if (noisy && (noise != nullptr || use_philox)) {
      const long long nq = (long long)Bpad * (d->nz / 4);
      den_seq_noise_kernel<<<dim3((unsigned)((nq + 255) / 256), (unsigned)ns), 256, 0, s>>>(reinterpret_cast<float4*>(w.nbuf), noise, B, Bpad, d->nz, T, s0, ns, seed, chain0);
      S.noiseT = reinterpret_cast<const float4*>(w.nbuf);
      count_launch();
    }
    profile_mark(s, true);
    {
      cudaLaunchConfig_t cfg{};
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3((unsigned)(pair ? std::min((items + 1) & ~1, sms & ~1) : std::min(items, sms)));
      cfg.blockDim = dim3(SQ_THREADS_GATE);
      cfg.dynamicSmemBytes = smem_g;
      cfg.stream = s;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      DAMC_CUDA(cudaLaunchKernelEx(&cfg, gate_kernel, Gp));
    }
    profile_mark(s, false);
    profile_mark(s, true);
    step_kernel<<<Bpad / 128, SQ_THREADS, smem, s>>>(S);
    profile_mark(s, false);
    count_launch(2);
  }
  DAMC_CUDA(cudaGetLastError());
  if (dbg) {   // experiment mode: stamps of CTA 0's second step, ns relative to the step's first stamp
    unsigned long long h[256];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, tlog, sizeof(h), cudaMemcpyDeviceToHost);
    const long long t0 = (long long)h[0];
    if (dbg_gate) {
      for (int ti = 0; ti < Gp.ntiles; ++ti) {
        const unsigned long long* e = h + 10 * ti;
        fprintf(stderr, "[sq gate] tile %2d L%d: tma-kb0 %6lld | mma-start %6lld acc-free %6lld ctx-ready+w-wait %6lld w-full %6lld issued %6lld\n", ti,
                Gp.tile[ti].layer, (long long)e[8] - t0, (long long)e[0] - t0, (long long)e[1] - t0, (long long)e[2] - t0, (long long)e[3] - t0, (long long)e[4] - t0);
      }
      return DAMC_OK;
    }
    for (int ti = 0; ti < S.ntiles; ++ti) {
      const unsigned long long* e = h + 10 * ti;
      fprintf(stderr, "[sq] tile %2d kind %d L%d: tma-kb0 %6lld | mma-start %6lld acc-free %6lld w-wait %6lld w-full %6lld issued %6lld | epi: wait-start %6lld acc-full %6lld done %6lld last-warp %6lld\n",
              ti, S.tile[ti].kind, S.tile[ti].layer, (long long)e[8] - t0, (long long)e[0] - t0, (long long)e[1] - t0, (long long)e[2] - t0, (long long)e[3] - t0,
              (long long)e[4] - t0, (long long)e[5] - t0, (long long)e[6] - t0, (long long)e[7] - t0, (long long)e[9] - t0);
    }
  }
  return DAMC_OK;
}

}  // namespace damc
