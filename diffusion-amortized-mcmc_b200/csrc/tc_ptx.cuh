// tc_ptx.cuh -- inline-PTX wrappers for the sm_100a tensor-core path (mbarrier, TMA, tcgen05 / TMEM, clusters) and the
// denoiser's final-update helper, shared by gen_tc.cu (per-layer launches) and denoiser_cluster.cu (cluster-resident loop).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must trap (error reported to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 27)) __trap();
}
// Long waits of many warps (the epilogue warps wait most of a tile's main loop for its accumulator): back off with
// nanosleep after a few polls so that spinning warps stop burning issue slots and power -- the step is power-capped.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > 4) __nanosleep(spin > 64 ? 512 : 128);
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// generic-proxy accesses (st.global of operand rows) <-> async-proxy accesses (TMA loads of those rows), all state spaces
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// the box lands at the same smem offset in every CTA of `mask`, and each destination CTA's mbarrier (same offset) gets the bytes
__device__ __forceinline__ void tma_load_2d_mcast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]   (both operands K-major, bf16, fp32 accumulate)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// same with fp32 operand containers read as tf32 (K = 8 per instruction = the same 32 bytes of a swizzle row)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// arrive (once the issued MMAs have completed) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t v[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t v[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- cta_group::2 (CTA pair) variants ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1,
                                                int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive (once the issued MMAs have completed) on the barrier at the same smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// ---- issue by ONE elected lane of a CONVERGED warp -------------------------------------------------------------------
// Inside `if (lane == 0)` the compiler wraps every tcgen05.mma / tcgen05.commit (uniform-datapath operands) into an
// elect-and-retry loop and re-materialises its uniform registers; when the whole warp runs the issue loop and the election
// happens inside the asm statement, a k-block costs the issuing warp about half the instructions.
#define DAMC_UMMA_ELECT(KIND, GROUP)                                                                       \
  asm volatile(                                                                                            \
      "{\n\t.reg .pred pe, pa;\n\t"                                                                        \
      "elect.sync _|pe, 0xffffffff;\n\t"                                                                   \
      "setp.ne.b32 pa, %4, 0;\n\t"                                                                         \
      "@pe tcgen05.mma.cta_group::" GROUP ".kind::" KIND " [%0], %1, %2, %3, pa;\n\t}"                     \
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory")
template <bool F32, int CG>
__device__ __forceinline__ void umma_elect(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (F32) { if (CG == 2) DAMC_UMMA_ELECT("tf32", "2"); else DAMC_UMMA_ELECT("tf32", "1"); }
  else { if (CG == 2) DAMC_UMMA_ELECT("f16", "2"); else DAMC_UMMA_ELECT("f16", "1"); }
}
template <int CG>
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
  if (CG == 2)
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

// UMMA shared-memory descriptor, K-major SWIZZLE_128B: 8-row groups of 128-byte rows, 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // descriptor version (sm_100), bits [46,48)
  d |= (uint64_t)2 << 61;                        // layout type SWIZZLE_128B, bits [61,64)
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
// fp32 -> nearest tf32 (10-bit mantissa, ties away), kept in an fp32 container: kind::tf32 MMAs ignore the low 13 bits, so
// operands are rounded when they are STORED (zero-mean error) instead of being truncated when they are read
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// two fp32 -> one packed word of the operand type (uniform branch on the launch's precision)
__device__ __forceinline__ uint32_t pack2(bool fp16, float a, float b) { return fp16 ? pack_f16x2(a, b) : pack_bf16x2(a, b); }

// ---- DAMC denoiser epilogues (EPI_DEN_LAYER / EPI_DEN_FINAL) ------------------------------------------------------------
// out = (main + b) * sigmoid(gate + bg) + hyper_bias + skip + bs   (diffusion_net.py:439-445), formed in the kernel's
// denoiser branch from the four column blocks of the accumulator; bias quads (bg, 0, b, bs) per feature sit in smem.
// last layer, 4 features [f0, f0+4) of chain r.b:  eps = z + out (residual, :530-531); x0 prediction and ancestral update
// of z (:610-620), fp32 throughout
// den_final_math: the arithmetic on loaded operands; den_final_vals: loads + arithmetic (res = new z, or eps when d.eps_out is
// set); den_final_quad: the same + the store
__device__ __forceinline__ void den_final_math(const DenEpi& d, const float zt[4], const float nrm[4], bool noisy, const float outv[4],
                                               float res[4]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    // explicit rounding points: every tcgen05 schedule (per-layer launches, cluster kernel, hoisted schedule) inlines this
    // function, and a chain's result must not depend on FMA contractions the compiler would otherwise pick per call site
    const float eps = d.residual ? __fadd_rn(zt[q], outv[q]) : outv[q];
    const float pred = __fmul_rn(d.c_pred, fmaf(-eps, d.c_eps, zt[q]));
    float zn = d.last ? pred : fmaf(d.c_zt, zt[q], __fmul_rn(d.c_x, pred));
    if (noisy) zn = fmaf(d.c_std, nrm[q], zn);
    res[q] = d.eps_out ? eps : zn;
  }
}
__device__ __forceinline__ void den_final_vals(const DenEpi& d, int b, int f0, const float outv[4], float res[4]) {
  const long long zi = (long long)b * d.nz + f0;
  const float4 z4 = *reinterpret_cast<const float4*>(d.z + zi);
  const float zt[4] = {z4.x, z4.y, z4.z, z4.w};
  float nrm[4] = {0.f, 0.f, 0.f, 0.f};
  const bool noisy = d.eps_out == nullptr && !d.last && d.c_std != 0.f;
  if (noisy) {
    if (d.noise) {
      const float4 n4 = *reinterpret_cast<const float4*>(d.noise + zi);
      nrm[0] = n4.x; nrm[1] = n4.y; nrm[2] = n4.z; nrm[3] = n4.w;
    } else if (d.use_philox) {
      philox_normal4(d.seed_ptr ? *d.seed_ptr : d.seed, d.chain0 + (unsigned long long)b, d.step, (uint32_t)(f0 >> 2), nrm);
    }
  }
  den_final_math(d, zt, nrm, noisy, outv, res);
}
__device__ __forceinline__ void den_final_quad(const DenEpi& d, int b, int f0, const float outv[4]) {
  const long long zi = (long long)b * d.nz + f0;
  float res[4];
  den_final_vals(d, b, f0, outv, res);
  float* dstp = d.eps_out ? d.eps_out + zi : d.z + zi;
  *reinterpret_cast<float4*>(dstp) = make_float4(res[0], res[1], res[2], res[3]);
}

}  // namespace damc
