// damc_common.cuh -- shared host/device helpers of libdamc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/damc.h"

namespace damc {

// ---- error plumbing (no C++ exceptions across the ABI) -------------------------------------------------------------
void set_error(const char* fmt, ...);
#define DAMC_FAIL(code, ...)      \
  do {                            \
    damc::set_error(__VA_ARGS__); \
    return (code);                \
  } while (0)
#define DAMC_CUDA(expr)                                                                               \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) DAMC_FAIL(DAMC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                     __FILE__, __LINE__);                                             \
  } while (0)
#define DAMC_TRY(expr)           \
  do {                           \
    int _r = (expr);             \
    if (_r != DAMC_OK) return _r; \
  } while (0)

// ---- measurement hooks ------------------------------------------------------------------------------------------
void count_launch(int n = 1);
bool profiling();
void profile_mark(cudaStream_t s, bool begin);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- handle kinds ------------------------------------------------------------------------------------------------
enum HandleKind { H_MLP = 1, H_GEN = 2, H_TOY = 3, H_DEN = 4, H_ENC = 5 };

// ---- Philox4x32-10 (counter-based; shard-invariant: keyed by GLOBAL chain index and step) -------------------------
__host__ __device__ inline void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c[0], c[1], c[2], c[3], k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// 4 standard normals for (seed, chain, step, quad): element e of the chain's latent lives in quad e/4, lane e%4.
__device__ inline void philox_normal4(uint64_t seed, uint64_t chain, uint64_t step, uint32_t quad, float out[4]) {
  uint32_t c[4] = {quad, (uint32_t)step, (uint32_t)chain, (uint32_t)(chain >> 32) ^ (uint32_t)(step >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float two_m32 = 2.3283064365386963e-10f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float u1 = ((float)c[2 * i] + 0.5f) * two_m32;      // (0,1)
    const float u2 = ((float)c[2 * i + 1] + 0.5f) * two_m32;
    const float r = sqrtf(-2.0f * logf(u1));
    float s, co;
    sincospif(2.0f * u2, &s, &co);
    out[2 * i] = r * co;
    out[2 * i + 1] = r * s;
  }
}
__device__ inline float philox_normal1(uint64_t seed, uint64_t chain, uint64_t step, uint32_t elem) {
  float v[4];
  philox_normal4(seed, chain, step, elem >> 2, v);
  return v[elem & 3];
}

// gated refill kernels: nothing to do when the handle's source tensors are unchanged since the last pack
__device__ __forceinline__ bool gate_clean(const int* dirty) { return dirty != nullptr && *dirty == 0; }

__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace damc
