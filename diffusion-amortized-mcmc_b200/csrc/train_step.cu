// train_step.cu -- the optimiser half of the training step that CALLS the samplers: gradient-norm
// clipping + Adam / AdamW over one flat parameter buffer, as two launches instead of torch's clip_grad_norm_ (a norm per
// tensor, a stack, a norm, a multiply per tensor) + the multi-tensor optimiser.
//
// Replaces, per optimiser step, reference workspace/train_gen_recon.py:218-219 / :229-230 / :239-240
//   torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=100);  optimizer.step()
// for optim.Adam(betas=(0.5, 0.999)) (G, E) and optim.AdamW(weight_decay=1e-4) (Q)  (:152-154), with the data-parallel
// gradient average folded in (grad_scale = 1 / world_size after a SUM all-reduce).
#include <math.h>

#include <algorithm>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

constexpr int TS_PARTIALS = 1024;

// stage 1: block b writes sum (scale * g)^2 over its slice (fixed slices, fixed tree) ; stage 2: one block adds the partials
// in index order -> the norm is bit-reproducible for a given n
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, size_t n, float scale,
                                                            float* __restrict__ partials) {
  const size_t per = (n + gridDim.x - 1) / gridDim.x;
  const size_t b0 = (size_t)blockIdx.x * per, b1 = b0 + per < n ? b0 + per : n;
  float acc = 0.f;
  for (size_t i = b0 + threadIdx.x; i < b1; i += 256) {
    const float v = scale * g[i];
    acc = fmaf(v, v, acc);
  }
  __shared__ float red[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(256) sumsq_final_kernel(float* __restrict__ partials, int nparts) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nparts; i += 256) acc += partials[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += red[i];
    partials[TS_PARTIALS] = sqrtf(t);   // total norm, as clip_grad_norm_ returns it
  }
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
  float bc1, bc2_sqrt;   // 1 - beta1^t ; sqrt(1 - beta2^t)
  int decoupled;
};

// torch.optim.Adam / AdamW single-tensor semantics (no amsgrad, no maximize):
//   g <- clip_coef * grad_scale * g ; AdamW: p *= 1 - lr wd ; Adam: g += wd p
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n, const AdamArgs a,
                                                        const float* __restrict__ total_norm) {
  float coef = a.grad_scale;
  if (a.max_norm > 0.f) {
    const float c = a.max_norm / (*total_norm + 1e-6f);   // clip_grad_norm_: clamp(max_norm / (norm + 1e-6), max = 1)
    coef *= c < 1.f ? c : 1.f;
  }
  const float step_size = a.lr / a.bc1;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    float gi = coef * g[i], pi = p[i];
    if (a.weight_decay != 0.f) {
      if (a.decoupled) pi *= 1.f - a.lr * a.weight_decay; else gi = fmaf(a.weight_decay, pi, gi);
    }
    const float mi = a.beta1 * m[i] + (1.f - a.beta1) * gi;   // lerp form, as torch: m + (g - m)(1 - b1) differs in the last bit only
    const float vi = a.beta2 * v[i] + (1.f - a.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / a.bc2_sqrt + a.eps));
  }
}

// fp32 -> nearest TF32 value (ties away from zero), kept in fp32 containers: the tcgen05 kind::tf32 MMA reads the top 19 bits of
// its operands, i.e. TRUNCATES; rounding the operands first halves the perturbation (what cuBLAS' TF32 path does)
__global__ void __launch_bounds__(256) round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(src[i]));
    dst[i] = __uint_as_float(r);
  }
}

}  // namespace damc

using namespace damc;

extern "C" int damc_round_tf32(const float* src, float* dst, size_t n, void* stream) {
  if (!src || !dst) DAMC_FAIL(DAMC_ERR_INVALID, "damc_round_tf32: null argument");
  if (n == 0) return DAMC_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
  round_tf32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

extern "C" int damc_fused_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n_total,
                                    int nranges, const unsigned long long* range_begin, const unsigned long long* range_count,
                                    const int* range_step, float lr, float beta1, float beta2, float eps, float weight_decay,
                                    int decoupled, float max_norm, float grad_scale, float* scratch, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !scratch || n_total == 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_fused_clip_adam: null argument");
  if (nranges < 0 || (nranges > 0 && (!range_begin || !range_count || !range_step))) DAMC_FAIL(DAMC_ERR_INVALID, "damc_fused_clip_adam: bad ranges");
  cudaStream_t s = (cudaStream_t)stream;
  const int nparts = (int)std::min<size_t>(TS_PARTIALS, (n_total + 4095) / 4096);
  sumsq_partial_kernel<<<nparts, 256, 0, s>>>(grads, n_total, grad_scale, scratch);
  DAMC_CUDA(cudaGetLastError());
  sumsq_final_kernel<<<1, 256, 0, s>>>(scratch, nparts);
  DAMC_CUDA(cudaGetLastError());
  count_launch(2);
  for (int r = 0; r < nranges; ++r) {
    const size_t b = (size_t)range_begin[r], n = (size_t)range_count[r];
    if (n == 0) continue;
    if (b + n > n_total || range_step[r] < 1) DAMC_FAIL(DAMC_ERR_INVALID, "damc_fused_clip_adam: range %d out of bounds / step < 1", r);
    AdamArgs a{};
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
    a.grad_scale = grad_scale; a.decoupled = decoupled;
    a.bc1 = (float)(1.0 - pow((double)beta1, (double)range_step[r]));
    a.bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)range_step[r]));
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    clip_adam_kernel<<<blocks, 256, 0, s>>>(params + b, grads + b, exp_avg + b, exp_avg_sq + b, n, a, scratch + TS_PARTIALS);
    DAMC_CUDA(cudaGetLastError());
    count_launch();
  }
  return DAMC_OK;
}
