// gen_epilogue.cuh -- element loads/stores and the per-element GEMM epilogues shared by the SIMT (gen_simt.cu) and
// tcgen05 (gen_tc.cu) engines.  See damc_internal.h for the meaning of the epilogue kinds.
#pragma once
#include <cuda_fp16.h>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

__device__ __forceinline__ void load8(const float* p, float v[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float v[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load8(const __half* p, float v[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void load4(const __half* p, float v[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
  const float2 a = __half22float2(h[0]), b = __half22float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
__device__ __forceinline__ void store_t(__half* p, float v) { *p = __float2half_rn(v); }
__device__ __forceinline__ void load4(const float* p, float v[4]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float v[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
}
// fp32 container holding a tf32-rounded value (DAMC_PREC_TF32 storage type: rounded on store, plain fp32 on load)
struct tf32_t { float v; };
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void load8(const tf32_t* p, float v[8]) { load8(reinterpret_cast<const float*>(p), v); }
__device__ __forceinline__ void load4(const tf32_t* p, float v[4]) { load4(reinterpret_cast<const float*>(p), v); }
__device__ __forceinline__ float to_f(tf32_t v) { return v.v; }
__device__ __forceinline__ void store_t(tf32_t* p, float v) { p->v = round_tf32(v); }
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_t(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_t(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- epilogue for one accumulator element (shared with the tcgen05 engine) ----------------------------------------
template <typename T>
__device__ __forceinline__ void epilogue_elem(const GemmPlan& p, int split, int m, int b, int y, int x, int n,
                                              float acc, float& loss_acc) {
  const Epilogue& e = p.epi;
  switch (e.kind) {
    case EPI_FWD_ACT: {
      const float h = acc + (e.bias ? e.bias[n % e.bias_mod] : 0.f);
      const long long o = (long long)b * e.o_b + (long long)(y * e.sy + e.py) * e.o_y +
                          (long long)(x * e.sx + e.px) * e.o_x + n;
      store_t(reinterpret_cast<T*>(e.out) + o, h > 0.f ? h : e.slope * h);
    } break;
    case EPI_DGRAD_MASK: {
      const float a = to_f(reinterpret_cast<const T*>(e.act)[(long long)m * p.N + n]);
      const float v = acc * (a > 0.f ? 1.f : e.slope);
      long long o;
      if (e.planar_out) {
        const int Hh = p.Hm >> 1, Wh = p.Wm >> 1;
        const long long plane = (long long)((y & 1) * 2 + (x & 1)) * p.B * Hh * Wh * p.N;
        o = plane + (((long long)b * Hh + (y >> 1)) * Wh + (x >> 1)) * p.N + n;
      } else {
        o = (long long)m * p.N + n;
      }
      store_t(reinterpret_cast<T*>(e.out) + o, v);
    } break;
    case EPI_DGRAD_Z: {
      reinterpret_cast<float*>(e.out)[((long long)split * p.B + b) * e.nz_out + n] = acc;
    } break;
    case EPI_STORE_F32: {
      reinterpret_cast<float*>(e.out)[(long long)m * e.nz_out + n] = acc;
    } break;
    case EPI_STORE_F32_BIAS: {
      reinterpret_cast<float*>(e.out)[(long long)m * e.nz_out + n] = acc + (e.bias ? e.bias[n] : 0.f);
    } break;
    default: break;
    case EPI_FWD_LAST: {
      const float xh = tanhf(acc + (e.bias ? e.bias[n] : 0.f));
      const int oy = y * e.sy + e.py, ox = x * e.sx + e.px;
      const long long xi = (((long long)b * e.nc + n) * e.Ho + oy) * e.Wo + ox;
      if (e.xhat) e.xhat[xi] = xh;
      if (e.x) {
        const float r = xh - e.x[xi];
        const float g = r * e.inv_sigma2 * e.gscale * (1.f - xh * xh);
        loss_acc += 0.5f * e.inv_sigma2 * r * r;
        T* gc = reinterpret_cast<T*>(e.gcol);
        for (int kh = 0; kh < e.k; ++kh) {
          const int ny = oy + e.padding - kh;
          if (ny < 0 || ny % e.stride) continue;
          const int iy = ny / e.stride;
          if (iy >= e.Hi) continue;
          for (int kw = 0; kw < e.k; ++kw) {
            const int nx = ox + e.padding - kw;
            if (nx < 0 || nx % e.stride) continue;
            const int ix = nx / e.stride;
            if (ix >= e.Wi) continue;
            store_t(gc + (((long long)b * e.Hi + iy) * e.Wi + ix) * 64 + (kh * e.k + kw) * 4 + n, g);
          }
        }
      }
    } break;
  }
}


}  // namespace damc
