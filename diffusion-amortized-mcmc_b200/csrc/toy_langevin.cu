// toy_langevin.cu -- persistent posterior-Langevin loop for the 2-D toy example: all K steps in one launch with the
// whole ReLU-MLP generator (2-128-128-128-2) resident in shared memory.
//
// Replaces the closure sample_langevin_post_z of reference workspace/toy_example/toy_example.py:110-131
// (U = |G(z)-x|^2/(2 sigma^2) + |z|^2/2, no EBM term :118; G defined at :22-47).
// One warp owns CPW chains; hidden units are spread over lanes (4 per lane at nh=128).  The two square weight matrices
// use a row stride of nh+1 floats so that both the forward (lane = output row) and the backward (lane = input column)
// sweeps are bank-conflict free from one copy.
#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

struct ToyPack : damc_handle {
  int nz = 0, nh = 0, nx = 0;
  float* slab = nullptr;  // W1[nh][nz] b1[nh] W2[nh][nh] b2 W3[nh][nh] b3 W4[nx][nh] b4[nx]
  const float* src[8] = {};
  size_t sizes[8] = {};
  ~ToyPack() override { if (slab) cudaFree(slab); }
  int refill(cudaStream_t stream, const int* dirty) override {
    (void)dirty;   // tiny MLP: always copied
    float* p = slab;
    for (int i = 0; i < 8; ++i) {
      DAMC_CUDA(cudaMemcpyAsync(p, src[i], sizes[i] * sizeof(float), cudaMemcpyDeviceToDevice, stream));
      p += sizes[i];
    }
    return DAMC_OK;
  }
};

constexpr int TOY_NH = 128, TOY_MAXD = 8, TOY_WARPS = 8, TOY_CPW = 2, TOY_R = TOY_NH / 32;

struct ToyArgs {
  const float* slab;
  int nz, nx;
  float* z;
  const float* x;
  int B, K;
  float step, inv_sigma2;
  int with_noise;
  const float* noise;
  uint64_t seed, chain0, step0;
};

__global__ void __launch_bounds__(TOY_WARPS * 32, 1) toy_langevin_kernel(const ToyArgs a) {
  constexpr int NH = TOY_NH, S = TOY_NH + 1, R = TOY_R, CPW = TOY_CPW;
  extern __shared__ __align__(16) float smem[];
  const int nz = a.nz, nx = a.nx;
  float* sW1 = smem;                 // [NH][nz]
  float* sb1 = sW1 + NH * nz;        // [NH]
  float* sW2 = sb1 + NH;             // [NH][S]
  float* sb2 = sW2 + NH * S;
  float* sW3 = sb2 + NH;             // [NH][S]
  float* sb3 = sW3 + NH * S;
  float* sW4 = sb3 + NH;             // [nx][NH]
  float* sb4 = sW4 + nx * NH;        // [nx] (padded to TOY_MAXD)
  float* sact = sb4 + TOY_MAXD;      // per warp: [4][CPW][NH]  (a1, a2, a3, d)
  {
    const float* g = a.slab;
    for (int i = threadIdx.x; i < NH * nz; i += blockDim.x) sW1[i] = g[i];
    g += NH * nz;
    for (int i = threadIdx.x; i < NH; i += blockDim.x) sb1[i] = g[i];
    g += NH;
    for (int i = threadIdx.x; i < NH * NH; i += blockDim.x) sW2[(i / NH) * S + (i % NH)] = g[i];
    g += NH * NH;
    for (int i = threadIdx.x; i < NH; i += blockDim.x) sb2[i] = g[i];
    g += NH;
    for (int i = threadIdx.x; i < NH * NH; i += blockDim.x) sW3[(i / NH) * S + (i % NH)] = g[i];
    g += NH * NH;
    for (int i = threadIdx.x; i < NH; i += blockDim.x) sb3[i] = g[i];
    g += NH;
    for (int i = threadIdx.x; i < nx * NH; i += blockDim.x) sW4[i] = g[i];
    g += nx * NH;
    for (int i = threadIdx.x; i < nx; i += blockDim.x) sb4[i] = g[i];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* a1 = sact + (size_t)warp * 4 * CPW * NH;
  float* a2 = a1 + CPW * NH;
  float* a3 = a2 + CPW * NH;
  float* dv = a3 + CPW * NH;
  const int cbase = (blockIdx.x * TOY_WARPS + warp) * CPW;
  if (cbase >= a.B) return;

  float z[CPW][TOY_MAXD], xt[CPW][TOY_MAXD];
  bool live[CPW];
#pragma unroll
  for (int c = 0; c < CPW; ++c) {
    live[c] = cbase + c < a.B;
#pragma unroll
    for (int k = 0; k < TOY_MAXD; ++k) {
      z[c][k] = (live[c] && k < nz) ? a.z[(size_t)(cbase + c) * nz + k] : 0.f;
      xt[c][k] = (live[c] && k < nx) ? a.x[(size_t)(cbase + c) * nx + k] : 0.f;
    }
  }
  const float half_s2 = 0.5f * a.step * a.step;

  for (int it = 0; it < a.K; ++it) {
    // ---- forward ------------------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + 32 * r;
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        float h = sb1[j];
#pragma unroll
        for (int k = 0; k < TOY_MAXD; ++k)
          if (k < nz) h = fmaf(sW1[j * nz + k], z[c][k], h);
        a1[c * NH + j] = fmaxf(h, 0.f);
      }
    }
    __syncwarp();
    auto dense_fwd = [&](const float* W, const float* bias, const float* in, float* out) {
      float acc[R][CPW];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPW; ++c) acc[r][c] = bias[lane + 32 * r];
      for (int i = 0; i < NH; ++i) {
        float av[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) av[c] = in[c * NH + i];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float w = W[(lane + 32 * r) * S + i];
#pragma unroll
          for (int c = 0; c < CPW; ++c) acc[r][c] = fmaf(w, av[c], acc[r][c]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPW; ++c) out[c * NH + lane + 32 * r] = fmaxf(acc[r][c], 0.f);
    };
    dense_fwd(sW2, sb2, a1, a2);
    __syncwarp();
    dense_fwd(sW3, sb3, a2, a3);
    __syncwarp();
    float res[CPW][TOY_MAXD];  // (x_hat - x) / sigma^2
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
      for (int o = 0; o < TOY_MAXD; ++o) {
        if (o < nx) {
          float p = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r) p = fmaf(sW4[o * NH + lane + 32 * r], a3[c * NH + lane + 32 * r], p);
          p = warp_sum(p) + sb4[o];
          res[c][o] = (p - xt[c][o]) * a.inv_sigma2;
        } else {
          res[c][o] = 0.f;
        }
      }
    // ---- backward with respect to z -------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = lane + 32 * r;
#pragma unroll
      for (int c = 0; c < CPW; ++c) {
        float d = 0.f;
#pragma unroll
        for (int o = 0; o < TOY_MAXD; ++o)
          if (o < nx) d = fmaf(sW4[o * NH + j], res[c][o], d);
        dv[c * NH + j] = a3[c * NH + j] > 0.f ? d : 0.f;
      }
    }
    __syncwarp();
    auto dense_bwd = [&](const float* W, const float* dout, const float* act_in, float* din) {
      // din[i] = relu'(act_in[i]) * sum_j W[j][i] dout[j];  din may alias act_in's slot only after the loop
      float acc[R][CPW];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPW; ++c) acc[r][c] = 0.f;
      for (int j = 0; j < NH; ++j) {
        float dj[CPW];
#pragma unroll
        for (int c = 0; c < CPW; ++c) dj[c] = dout[c * NH + j];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float w = W[j * S + lane + 32 * r];
#pragma unroll
          for (int c = 0; c < CPW; ++c) acc[r][c] = fmaf(w, dj[c], acc[r][c]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPW; ++c) {
          const int i = lane + 32 * r;
          din[c * NH + i] = act_in[c * NH + i] > 0.f ? acc[r][c] : 0.f;
        }
      __syncwarp();
    };
    dense_bwd(sW3, dv, a2, a3);  // d2 -> a3 slot (a3 no longer needed)
    dense_bwd(sW2, a3, a1, dv);  // d1 -> dv slot
    float gz[CPW][TOY_MAXD];
#pragma unroll
    for (int c = 0; c < CPW; ++c)
#pragma unroll
      for (int k = 0; k < TOY_MAXD; ++k) {
        if (k < nz) {
          float p = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r) p = fmaf(sW1[(lane + 32 * r) * nz + k], dv[c * NH + lane + 32 * r], p);
          gz[c][k] = warp_sum(p);
        } else {
          gz[c][k] = 0.f;
        }
      }
    __syncwarp();
    // ---- update (every lane keeps an identical copy of z) -----------------------------------------------------------
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
      float nrm[TOY_MAXD];
#pragma unroll
      for (int k = 0; k < TOY_MAXD; ++k) nrm[k] = 0.f;
      if (a.with_noise && live[c]) {
        if (a.noise != nullptr) {
#pragma unroll
          for (int k = 0; k < TOY_MAXD; ++k)
            if (k < nz) nrm[k] = a.noise[((size_t)it * a.B + cbase + c) * nz + k];
        } else {
          philox_normal4(a.seed, a.chain0 + cbase + c, a.step0 + it, 0, nrm);
          if (nz > 4) philox_normal4(a.seed, a.chain0 + cbase + c, a.step0 + it, 1, nrm + 4);
        }
      }
#pragma unroll
      for (int k = 0; k < TOY_MAXD; ++k)
        if (k < nz) z[c][k] = z[c][k] - half_s2 * (gz[c][k] + z[c][k]) + a.step * nrm[k];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < CPW; ++c)
      if (live[c])
        for (int k = 0; k < nz; ++k) a.z[(size_t)(cbase + c) * nz + k] = z[c][k];
  }
}

}  // namespace damc

using namespace damc;

extern "C" int damc_pack_toy_mlp(damc_handle** out, int nz, int nh, int nx, const float* const* host_W,
                                 const float* const* host_b, void* stream) {
  if (!out || !host_W || !host_b) DAMC_FAIL(DAMC_ERR_INVALID, "damc_pack_toy_mlp: null argument");
  if (nh != TOY_NH || nz < 1 || nz > TOY_MAXD || nx < 1 || nx > TOY_MAXD)
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "toy generator: supported nh=128, nz<=8, nx<=8 (nz=%d nh=%d nx=%d)", nz, nh, nx);
  ToyPack* t = new ToyPack();
  t->kind = H_TOY; t->nz = nz; t->nh = nh; t->nx = nx;
  const size_t sizes[8] = {(size_t)nh * nz, (size_t)nh, (size_t)nh * nh, (size_t)nh, (size_t)nh * nh, (size_t)nh,
                           (size_t)nx * nh, (size_t)nx};
  size_t total = 0;
  for (size_t s : sizes) total += s;
  if (cudaMalloc(&t->slab, total * sizeof(float)) != cudaSuccess) { delete t; DAMC_FAIL(DAMC_ERR_CUDA, "cudaMalloc failed"); }
  for (int i = 0; i < 4; ++i) {
    t->src[2 * i] = host_W[i]; t->src[2 * i + 1] = host_b[i];
    t->sizes[2 * i] = sizes[2 * i]; t->sizes[2 * i + 1] = sizes[2 * i + 1];
  }
  const int r = t->refill((cudaStream_t)stream, nullptr);
  if (r != DAMC_OK) { delete t; return r; }
  *out = t;
  return DAMC_OK;
}

extern "C" int damc_toy_posterior_langevin(const damc_handle* mlp, float* z, const float* x, int B, int K,
                                           float step_size, float sigma, int with_noise, const float* noise,
                                           uint64_t seed, uint64_t chain0, uint64_t step0, void* stream) {
  if (!mlp || mlp->kind != H_TOY) DAMC_FAIL(DAMC_ERR_INVALID, "damc_toy_posterior_langevin: not a toy-MLP handle");
  if (!z || !x || B <= 0 || K < 0 || !(sigma > 0.f)) DAMC_FAIL(DAMC_ERR_INVALID, "damc_toy_posterior_langevin: bad arguments");
  if (K == 0) return DAMC_OK;
  const ToyPack* t = static_cast<const ToyPack*>(mlp);
  ToyArgs a{};
  a.slab = t->slab; a.nz = t->nz; a.nx = t->nx; a.z = z; a.x = x; a.B = B; a.K = K; a.step = step_size;
  a.inv_sigma2 = 1.0f / (sigma * sigma); a.with_noise = with_noise; a.noise = noise;
  a.seed = seed; a.chain0 = chain0; a.step0 = step0;
  const int NH = TOY_NH, S = NH + 1;
  const size_t smem = sizeof(float) * ((size_t)NH * t->nz + NH + 2 * ((size_t)NH * S + NH) + (size_t)t->nx * NH +
                                       TOY_MAXD + (size_t)TOY_WARPS * 4 * TOY_CPW * NH);
  DAMC_CUDA(cudaFuncSetAttribute(toy_langevin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int per_cta = TOY_WARPS * TOY_CPW;
  toy_langevin_kernel<<<ceil_div(B, per_cta), TOY_WARPS * 32, smem, (cudaStream_t)stream>>>(a);
  DAMC_CUDA(cudaGetLastError());
  count_launch();
  return DAMC_OK;
}
