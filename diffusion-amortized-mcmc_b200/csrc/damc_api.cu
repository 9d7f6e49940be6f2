// damc_api.cu -- extern "C" entry points of libdamc_b200 (declared in include/damc.h).
#include <stdarg.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <utility>
#include <vector>
#include <string.h>

#include "damc_common.cuh"
#include "damc_internal.h"

namespace damc {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_events;
static size_t g_events_used = 0;
static bool g_open = false;

void count_launch(int n) { g_launches += n; }
bool profiling() { return g_profile; }
void profile_mark(cudaStream_t s, bool begin) {
  if (!g_profile) return;
  if (begin) {
    if (g_events_used == g_events.size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      g_events.emplace_back(a, b);
    }
    cudaEventRecord(g_events[g_events_used].first, s);
    g_open = true;
  } else if (g_open) {
    cudaEventRecord(g_events[g_events_used].second, s);
    ++g_events_used;
    g_open = false;
  }
}

int build_generator(GenPack* g, int nlayers, const damc_convt_layer* L, float slope, int precision, cudaStream_t stream);
int dz_splits(const GenPack* g, int B);

static int check_sm100() {
  int dev = 0, major = 0;
  DAMC_CUDA(cudaGetDevice(&dev));
  DAMC_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "libdamc_b200 is built for sm_100a only (device has compute capability %d.x)", major);
  return DAMC_OK;
}

void MlpPack::sources(std::vector<HashSrc>& out) const {
  const size_t n[6] = {(size_t)ndf * nz, (size_t)ndf, (size_t)ndf * ndf, (size_t)ndf, (size_t)ndf, 1};
  for (int i = 0; i < 6; ++i) out.push_back(HashSrc{src[i], (unsigned long long)n[i], 0ull});
}

int MlpPack::refill(cudaStream_t s, const int* dirty) {
  DAMC_TRY(launch_gated_copy(W1, src[0], (size_t)ndf * nz, dirty, s));
  DAMC_TRY(launch_gated_copy(b1, src[1], ndf, dirty, s));
  DAMC_TRY(launch_gated_copy(W2, src[2], (size_t)ndf * ndf, dirty, s));
  DAMC_TRY(launch_gated_copy(b2, src[3], ndf, dirty, s));
  DAMC_TRY(launch_gated_copy(w3, src[4], ndf, dirty, s));
  DAMC_TRY(launch_gated_copy(b3, src[5], 1, dirty, s));
  DAMC_TRY(launch_transpose(src[0], W1T, ndf, nz, s, dirty));   // from the caller's tensors: no dependence on the copies above
  DAMC_TRY(launch_transpose(src[2], W2T, ndf, ndf, s, dirty));
  DAMC_TRY(ebm_tc_refill(this, DAMC_PREC_BF16, s, dirty));   // no-ops until the tensor-core step kernel has asked for them
  DAMC_TRY(ebm_tc_refill(this, DAMC_PREC_FP16, s, dirty));
  return DAMC_OK;
}

// ---- change detection for damc_repack ------------------------------------------------------------------------------------
// h = sum over all source elements of mix(bits, global element index)  (64-bit, order-independent, so blocks add atomically).
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  return x;
}
__global__ void __launch_bounds__(256) weights_hash_kernel(const HashSrc* __restrict__ tab, int ntab,
                                                           unsigned long long* __restrict__ st, int store_only) {
  unsigned long long acc = 0ull;
  const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nth = (unsigned long long)gridDim.x * blockDim.x;
  for (int t = 0; t < ntab; ++t) {
    const HashSrc s = tab[t];
    const unsigned long long n4 = ((reinterpret_cast<unsigned long long>(s.p) & 15ull) == 0ull) ? s.n >> 2 : 0ull;
    const uint4* p4 = reinterpret_cast<const uint4*>(s.p);
    for (unsigned long long i = tid; i < n4; i += nth) {
      const uint4 v = __ldg(p4 + i);
      const unsigned long long k = (s.off + 4ull * i + 1ull) * 0x9E3779B97F4A7C15ull;
      acc += mix64((((unsigned long long)v.y << 32) | v.x) ^ k) + mix64((((unsigned long long)v.w << 32) | v.z) ^ (k + 0x632BE59BD9B4E019ull));
    }
    const unsigned int* p1 = reinterpret_cast<const unsigned int*>(s.p);
    for (unsigned long long i = 4ull * n4 + tid; i < s.n; i += nth)
      acc += mix64((unsigned long long)__ldg(p1 + i) ^ ((s.off + i + 1ull) * 0xC2B2AE3D27D4EB4Full));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ unsigned long long part[8];
  __shared__ bool is_last;
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long b = 0ull;
    for (int w = 0; w < 8; ++w) b += part[w];
    atomicAdd(&st[0], b);
    __threadfence();
    is_last = atomicAdd(&st[3], 1ull) == (unsigned long long)gridDim.x - 1ull;
    if (is_last) {   // every block's sum has landed: compare with the hash the packed buffers were built from
      __threadfence();
      const unsigned long long total = atomicAdd(&st[0], 0ull);
      *reinterpret_cast<int*>(&st[2]) = (store_only || total == st[1]) ? 0 : 1;
      st[1] = total;
      st[0] = 0ull;
      st[3] = 0ull;
    }
  }
}

__global__ void gated_copy_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n, const int* __restrict__ dirty) {
  if (gate_clean(dirty)) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int launch_gated_copy(float* dst, const float* src, size_t n, const int* dirty, cudaStream_t stream) {
  if (n == 0) return DAMC_OK;
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 592);
  gated_copy_kernel<<<blocks, 256, 0, stream>>>(dst, src, n, dirty);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

static int launch_hash(damc_handle* h, int store_only, cudaStream_t stream) {
  weights_hash_kernel<<<296, 256, 0, stream>>>(h->hash_tab, h->hash_n, h->hash_state, store_only);
  DAMC_CUDA(cudaGetLastError());
  return DAMC_OK;
}

int handle_hash_init(damc_handle* h, cudaStream_t stream) {
  std::vector<HashSrc> tab;
  h->sources(tab);
  if (tab.empty() || getenv("DAMC_REPACK_ALWAYS")) return DAMC_OK;
  unsigned long long off = 0;
  for (HashSrc& s : tab) { s.off = off; off += s.n; }
  DAMC_CUDA(cudaMalloc(&h->hash_tab, sizeof(HashSrc) * tab.size()));
  DAMC_CUDA(cudaMalloc(&h->hash_state, 4 * sizeof(unsigned long long)));
  h->hash_n = (int)tab.size();
  // pageable host source: the copy is staged before the call returns, so the local table may go out of scope
  DAMC_CUDA(cudaMemcpyAsync(h->hash_tab, tab.data(), sizeof(HashSrc) * tab.size(), cudaMemcpyHostToDevice, stream));
  DAMC_CUDA(cudaMemsetAsync(h->hash_state, 0, 4 * sizeof(unsigned long long), stream));
  return launch_hash(h, 1, stream);
}

int handle_repack(damc_handle* h, cudaStream_t stream) {
  if (!h->hash_tab) return h->refill(stream, nullptr);
  DAMC_TRY(launch_hash(h, 0, stream));
  return h->refill(stream, reinterpret_cast<const int*>(&h->hash_state[2]));
}

}  // namespace damc

unsigned long long damc_next_uid() {
  static std::atomic<unsigned long long> next{1};
  return next.fetch_add(1);
}

using namespace damc;

extern "C" {

int damc_version(void) { return 100; }
int damc_selftest(void) { return tc_selftest_fastdiv(); }
const char* damc_last_error(void) { return g_err; }

long long damc_launch_count(void) { return g_launches.load(); }
int damc_profile_enable(int on) {
  g_profile = on != 0;
  g_events_used = 0;
  g_open = false;
  return DAMC_OK;
}
int damc_profile_collect(double* gemm_ms, long long* gemm_launches) {
  double ms = 0.0;
  for (size_t i = 0; i < g_events_used; ++i) {
    DAMC_CUDA(cudaEventSynchronize(g_events[i].second));
    float t = 0.f;
    DAMC_CUDA(cudaEventElapsedTime(&t, g_events[i].first, g_events[i].second));
    ms += t;
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_launches) *gemm_launches = (long long)g_events_used;
  g_events_used = 0;
  return DAMC_OK;
}

int damc_repack(damc_handle* h, void* stream) {
  if (!h) DAMC_FAIL(DAMC_ERR_INVALID, "damc_repack: null handle");
  return handle_repack(h, (cudaStream_t)stream);
}

int damc_free(damc_handle* h) {
  delete h;
  return DAMC_OK;
}

int damc_pack_mlp(damc_handle** out, int nz, int ndf, const float* W1, const float* b1, const float* W2,
                  const float* b2, const float* W3, const float* b3, float negative_slope, void* stream) {
  if (!out || !W1 || !b1 || !W2 || !b2 || !W3 || !b3) DAMC_FAIL(DAMC_ERR_INVALID, "damc_pack_mlp: null argument");
  if (nz < 1 || nz > 128 || ndf < 2 || ndf > 256) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "damc_pack_mlp: supported 1<=nz<=128, 2<=ndf<=256 (nz=%d ndf=%d)", nz, ndf);
  DAMC_TRY(check_sm100());
  cudaStream_t s = (cudaStream_t)stream;
  MlpPack* m = new MlpPack();
  m->kind = H_MLP; m->nz = nz; m->ndf = ndf; m->slope = negative_slope;
  const size_t n = 2 * ((size_t)ndf * nz + (size_t)ndf * ndf) + 3 * (size_t)ndf + 1;
  if (cudaMalloc(&m->slab, n * sizeof(float)) != cudaSuccess) { delete m; DAMC_FAIL(DAMC_ERR_CUDA, "damc_pack_mlp: cudaMalloc failed"); }
  float* p = m->slab;
  m->W1 = p; p += (size_t)ndf * nz;
  m->b1 = p; p += ndf;
  m->W2 = p; p += (size_t)ndf * ndf;
  m->b2 = p; p += ndf;
  m->w3 = p; p += ndf;
  m->b3 = p; p += 1;
  m->W1T = p; p += (size_t)ndf * nz;
  m->W2T = p;
  const float* srcs[6] = {W1, b1, W2, b2, W3, b3};
  for (int i = 0; i < 6; ++i) m->src[i] = srcs[i];
  int r = m->refill(s, nullptr);
  if (r == DAMC_OK) r = handle_hash_init(m, s);
  if (r != DAMC_OK) { delete m; return r; }
  *out = m;
  return DAMC_OK;
}

int damc_pack_generator(damc_handle** out, int nlayers, const damc_convt_layer* host_layers, float negative_slope,
                        int precision, void* stream) {
  if (!out || !host_layers) DAMC_FAIL(DAMC_ERR_INVALID, "damc_pack_generator: null argument");
  DAMC_TRY(check_sm100());
  GenPack* g = new GenPack();
  const int r = build_generator(g, nlayers, host_layers, negative_slope, precision, (cudaStream_t)stream);
  if (r != DAMC_OK) { delete g; return r; }
  *out = g;
  return DAMC_OK;
}

int damc_generator_shape(const damc_handle* gen, int* nz, int* nc, int* height, int* width) {
  if (!gen || gen->kind != H_GEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_generator_shape: not a generator handle");
  const GenPack* g = static_cast<const GenPack*>(gen);
  if (nz) *nz = g->nz;
  if (nc) *nc = g->nc;
  if (height) *height = g->H;
  if (width) *width = g->W;
  return DAMC_OK;
}

size_t damc_generator_workspace_bytes(const damc_handle* gen, int B) {
  if (!gen || gen->kind != H_GEN || B <= 0) return 0;
  GenWorkspace ws;
  plan_workspace(static_cast<const GenPack*>(gen), B, nullptr, &ws);
  return ws.bytes;
}

int damc_generator_forward(const damc_handle* gen, const float* z, float* x_hat, int B, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!gen || gen->kind != H_GEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_generator_forward: not a generator handle");
  if (!z || !x_hat || B <= 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_generator_forward: bad arguments");
  const GenPack* g = static_cast<const GenPack*>(gen);
  GenWorkspace ws;
  plan_workspace(g, B, workspace, &ws);
  if (!workspace || workspace_bytes < ws.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
  return generator_forward(g, ws, z, B, nullptr, 1.0f, x_hat, nullptr, (cudaStream_t)stream);
}

// D[m][n] = sum_k A[m][k] W[n][k] (+ bias[n]) on the tcgen05 engine, kind::tf32 (fp32 tensors as they are: the MMA reads the top
// 19 bits of each operand).  The plain-GEMM plan of the generator's first layer: a 1 x 1 pixel grid, M "chains", one tap.
int damc_gemm_tf32(const float* A, const float* W, const float* bias, float* D, int M, int N, int K, int ldd, void* stream) {
  if (!A || !W || !D || M <= 0 || N <= 0 || K <= 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_gemm_tf32: bad arguments");
  if (K % 32 || N % 16 || ldd < N || ldd % 4) DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "damc_gemm_tf32: K %% 32, N %% 16, ldd %% 4 == 0 required (M=%d N=%d K=%d ldd=%d)", M, N, K, ldd);
  if (((uintptr_t)A | (uintptr_t)W | (uintptr_t)D) & 15) DAMC_FAIL(DAMC_ERR_INVALID, "damc_gemm_tf32: pointers must be 16-byte aligned");
  DAMC_TRY(check_sm100());
  GemmPlan p{};
  p.A = A; p.B = M; p.Hm = 1; p.Wm = 1; p.Cs = K;
  p.ntaps = 1; p.taps[0] = Tap{0, 0, 0, 0};
  p.N = p.Np = N; p.ksplit = 1;
  p.Wtc = W;
  p.epi.kind = bias ? EPI_STORE_F32_BIAS : EPI_STORE_F32;
  p.epi.bias = bias;
  p.epi.out = D;
  p.epi.nz_out = ldd;
  const int r = launch_gemm_tc(p, DAMC_PREC_TF32, (cudaStream_t)stream);
  count_launch();
  return r;
}

int damc_posterior_score(const damc_handle* gen, const damc_handle* ebm, const float* z, const float* x, int B, float* score,
                         float* sqerr, void* workspace, size_t workspace_bytes, void* stream) {
  if (!gen || gen->kind != H_GEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_score: not a generator handle");
  if (ebm && ebm->kind != H_MLP) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_score: not an EBM handle");
  if (!z || !x || B <= 0 || (!score && !sqerr)) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_score: bad arguments");
  const GenPack* g = static_cast<const GenPack*>(gen);
  const MlpPack* m = static_cast<const MlpPack*>(ebm);
  if (m && m->nz != g->nz) DAMC_FAIL(DAMC_ERR_INVALID, "EBM nz %d != generator nz %d", m->nz, g->nz);
  GenWorkspace ws;
  plan_workspace(g, B, workspace, &ws);
  if (!workspace || workspace_bytes < ws.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
  cudaStream_t s = (cudaStream_t)stream;
  DAMC_TRY(generator_score_forward(g, ws, z, B, x, s));
  return launch_ebm_score(m, z, B, g->nz, ws.sq_part, score_parts(g), score, sqerr, s);
}

int damc_prior_langevin(const damc_handle* ebm, float* z, int B, int K, float step_size, int with_noise,
                        const float* noise, uint64_t seed, uint64_t chain0, uint64_t step0, float* trace,
                        void* stream) {
  if (!ebm || ebm->kind != H_MLP) DAMC_FAIL(DAMC_ERR_INVALID, "damc_prior_langevin: not an EBM handle");
  if (!z) DAMC_FAIL(DAMC_ERR_INVALID, "damc_prior_langevin: null z");
  cudaStream_t s = (cudaStream_t)stream;
  if (K == 0) return DAMC_OK;
  if (trace) DAMC_CUDA(cudaMemsetAsync(trace, 0, sizeof(float) * 2 * K, s));
  return launch_ebm_langevin(static_cast<const MlpPack*>(ebm), z, B, K, step_size, with_noise, noise, seed, chain0,
                             step0, trace, 2, nullptr, 0, 0, 0, s);
}

int damc_prior_langevin_tc(const damc_handle* ebm, float* z, int B, int K, float step_size, int with_noise, const float* noise,
                           uint64_t seed, uint64_t chain0, uint64_t step0, void* stream) {
  if (!ebm || ebm->kind != H_MLP) DAMC_FAIL(DAMC_ERR_INVALID, "damc_prior_langevin_tc: not an EBM handle");
  if (!z || B <= 0 || K < 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_prior_langevin_tc: bad arguments");
  const MlpPack* m = static_cast<const MlpPack*>(ebm);
  if (!ebm_tc_usable(m, DAMC_PREC_FP16, nullptr))
    DAMC_FAIL(DAMC_ERR_UNSUPPORTED, "damc_prior_langevin_tc: needs nz %% 4 == 0, nz <= 128, ndf <= 256 (nz=%d ndf=%d) and DAMC_EBM_TC != 0", m->nz, m->ndf);
  return launch_ebm_step_tc(m, DAMC_PREC_FP16, z, B, step_size, with_noise, noise, seed, chain0, step0, nullptr, 0, 0, 1.0f,
                            (cudaStream_t)stream, nullptr, K);
}

int damc_posterior_langevin(const damc_handle* gen, const damc_handle* ebm, float* z, const float* x, int B, int K,
                            float step_size, float sigma, int with_noise, const float* noise, uint64_t seed,
                            uint64_t chain0, uint64_t step0, float* trace, float* x_hat_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (!gen || gen->kind != H_GEN) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_langevin: not a generator handle");
  if (ebm && ebm->kind != H_MLP) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_langevin: not an EBM handle");
  if (!z || !x || B <= 0 || K < 0) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_langevin: bad arguments");
  if (!(sigma > 0.f)) DAMC_FAIL(DAMC_ERR_INVALID, "damc_posterior_langevin: sigma must be > 0");
  const GenPack* g = static_cast<const GenPack*>(gen);
  const MlpPack* m = static_cast<const MlpPack*>(ebm);
  if (m && m->nz != g->nz) DAMC_FAIL(DAMC_ERR_INVALID, "EBM nz %d != generator nz %d", m->nz, g->nz);
  cudaStream_t s = (cudaStream_t)stream;
  GenWorkspace ws;
  plan_workspace(g, B, workspace, &ws);
  if (!workspace || workspace_bytes < ws.bytes) DAMC_FAIL(DAMC_ERR_WORKSPACE, "workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
  if (K == 0) return DAMC_OK;
  const GenLayer& last = g->layers[g->nlayers - 1];
  const int S = dz_splits(g, B);
  // the K x (forward, likelihood gradient, dgrad, fused EBM + update) launches on stream st, directly or into a capture
  auto issue = [&](float* zz, const float* xx, const unsigned long long* seed_ptr, cudaStream_t st) -> int {
    // unused im2col slots (image border taps, channel padding) must read as zero; live slots are rewritten every step
    DAMC_CUDA(cudaMemsetAsync(ws.gcol, 0, elem_size(g->precision) * (size_t)B * last.Hin * last.Win * 64, st));
    for (int i = 0; i < K; ++i) {
      float* tr = trace ? trace + 4 * (size_t)i : nullptr;
      DAMC_TRY(generator_forward(g, ws, zz, B, xx, sigma, (i == K - 1) ? x_hat_out : nullptr, tr ? tr + 1 : nullptr, st));
      DAMC_TRY(generator_dgrad(g, ws, B, st));
      const float* nz_i = noise ? noise + (size_t)i * B * g->nz : nullptr;
      if (ebm_tc_usable(m, g->precision, tr))   // 16-bit modes, no trace: the EBM tail as four tcgen05 GEMMs per 128-chain tile
        DAMC_TRY(launch_ebm_step_tc(m, g->precision, zz, B, step_size, with_noise, nz_i, seed, seed_ptr ? 0 : chain0,
                                    (seed_ptr ? 0 : step0) + (uint64_t)i, ws.dz_part, S, g->nz_p, 1.0f / generator_grad_scale(g, sigma),
                                    st, seed_ptr));
      else
        DAMC_TRY(launch_ebm_step(m, zz, B, step_size, with_noise, nz_i, seed,
                                 seed_ptr ? 0 : chain0, (seed_ptr ? 0 : step0) + (uint64_t)i, tr, ws.dz_part, S, g->nz_p,
                                 1.0f / generator_grad_scale(g, sigma), g->nz, st, seed_ptr));
    }
    return DAMC_OK;
  };
  if (trace) DAMC_CUDA(cudaMemsetAsync(trace, 0, sizeof(float) * 4 * K, s));
  // ---- CUDA-graph replay (tensor-core modes, Philox noise, no trace): no per-step host launches from the second call on --
  static const bool use_graph = []{ const char* e = getenv("DAMC_GRAPH"); return !(e && e[0] == '0'); }();
  const bool graphable = use_graph && g->use_tc && !profiling() && noise == nullptr && trace == nullptr && K > 1;
  if (!graphable) return issue(z, x, nullptr, s);
  // (bit 1 of the noise field: which form of the EBM tail the sequence launches -- DAMC_EBM_TC may change between calls)
  const GenPack::GraphKey key = {B, K, with_noise | (ebm_tc_usable(m, g->precision, nullptr) ? 2 : 0), x_hat_out != nullptr ? 1 : 0, step_size, sigma, ws.base, m ? m->uid : 0ull};
  auto same = [&](const GenPack::GraphKey& k0) {
    return k0.B == key.B && k0.K == key.K && k0.with_noise == key.with_noise && k0.want_xhat == key.want_xhat &&
           k0.step == key.step && k0.sigma == key.sigma && k0.ws_base == key.ws_base && k0.ebm == key.ebm;
  };
  GenPack::GraphEntry* ent = nullptr;
  for (GenPack::GraphEntry& e : g->graphs) if (same(e.key)) { ent = &e; break; }
  if (!ent) {   // new configuration: run it directly once; it is captured if it comes back
    if ((int)g->graphs.size() >= GenPack::kMaxGraphs) {   // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < g->graphs.size(); ++i) if (g->graphs[i].last_use < g->graphs[lru].last_use) lru = i;
      if (g->graphs[lru].gexec) cudaGraphExecDestroy(g->graphs[lru].gexec);
      g->graphs.erase(g->graphs.begin() + (long)lru);
    }
    GenPack::GraphEntry e;
    e.key = key;
    e.last_use = ++g->graph_clock;
    g->graphs.push_back(e);
    return issue(z, x, nullptr, s);
  }
  ent->last_use = ++g->graph_clock;
  float* const xhat_user = x_hat_out;
  if (!ent->gexec) {
    if (!g->cap_stream) DAMC_CUDA(cudaStreamCreateWithFlags(&g->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    DAMC_CUDA(cudaStreamBeginCapture(g->cap_stream, cudaStreamCaptureModeThreadLocal));
    const long long n0 = g_launches.load();
    x_hat_out = xhat_user ? ws.xhat_buf : nullptr;   // the captured sequence writes G(z) of the last step into the workspace
    const int r = issue(ws.zbuf, ws.xbuf, ws.seed_dev, g->cap_stream);
    x_hat_out = xhat_user;
    ent->launches = g_launches.load() - n0;
    g_launches -= ent->launches;   // nothing ran yet: replays are counted when they are launched
    const cudaError_t ce = cudaStreamEndCapture(g->cap_stream, &graph);
    if (r != DAMC_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess || !graph) DAMC_FAIL(DAMC_ERR_CUDA, "posterior: stream capture failed: %s", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&ent->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { ent->gexec = nullptr; DAMC_FAIL(DAMC_ERR_CUDA, "posterior: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
    ++g->graph_captures;
  }
  const unsigned long long rng_host[3] = {seed, chain0, step0};
  DAMC_CUDA(cudaMemcpyAsync(ws.seed_dev, rng_host, sizeof(rng_host), cudaMemcpyHostToDevice, s));
  DAMC_CUDA(cudaMemcpyAsync(ws.zbuf, z, sizeof(float) * (size_t)B * g->nz, cudaMemcpyDeviceToDevice, s));
  DAMC_CUDA(cudaMemcpyAsync(ws.xbuf, x, sizeof(float) * (size_t)B * g->nc * g->H * g->W, cudaMemcpyDeviceToDevice, s));
  DAMC_CUDA(cudaGraphLaunch(ent->gexec, s));
  count_launch((int)ent->launches);
  ++g->graph_replays;
  DAMC_CUDA(cudaMemcpyAsync(z, ws.zbuf, sizeof(float) * (size_t)B * g->nz, cudaMemcpyDeviceToDevice, s));
  if (xhat_user)
    DAMC_CUDA(cudaMemcpyAsync(xhat_user, ws.xhat_buf, sizeof(float) * (size_t)B * g->nc * g->H * g->W, cudaMemcpyDeviceToDevice, s));
  return DAMC_OK;
}

long long damc_graph_replays(const damc_handle* gen) {
  if (!gen || gen->kind != H_GEN) return -1;
  return static_cast<const GenPack*>(gen)->graph_replays;
}

}  // extern "C"
