// damc_api.cu -- extern "C" entry points of libdamc_b200 (declared in include/damc.h).
#include <stdarg.h>
#include <stdlib.h>
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

int MlpPack::refill(cudaStream_t s) {
  const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
  DAMC_CUDA(cudaMemcpyAsync(W1, src[0], sizeof(float) * ndf * nz, k, s));
  DAMC_CUDA(cudaMemcpyAsync(b1, src[1], sizeof(float) * ndf, k, s));
  DAMC_CUDA(cudaMemcpyAsync(W2, src[2], sizeof(float) * ndf * ndf, k, s));
  DAMC_CUDA(cudaMemcpyAsync(b2, src[3], sizeof(float) * ndf, k, s));
  DAMC_CUDA(cudaMemcpyAsync(w3, src[4], sizeof(float) * ndf, k, s));
  DAMC_CUDA(cudaMemcpyAsync(b3, src[5], sizeof(float), k, s));
  DAMC_TRY(launch_transpose(W1, W1T, ndf, nz, s));
  DAMC_TRY(launch_transpose(W2, W2T, ndf, ndf, s));
  return DAMC_OK;
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
  return h->refill((cudaStream_t)stream);
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
  const int r = m->refill(s);
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
      DAMC_TRY(launch_ebm_step(m, zz, B, step_size, with_noise, noise ? noise + (size_t)i * B * g->nz : nullptr, seed,
                               chain0, step0 + (uint64_t)i, tr, ws.dz_part, S, g->nz_p,
                               1.0f / generator_grad_scale(g, sigma), g->nz, st, seed_ptr));
    }
    return DAMC_OK;
  };
  if (trace) DAMC_CUDA(cudaMemsetAsync(trace, 0, sizeof(float) * 4 * K, s));
  // ---- CUDA-graph replay (tensor-core modes, Philox noise, no trace): no per-step host launches from the second call on --
  static const bool use_graph = []{ const char* e = getenv("DAMC_GRAPH"); return !(e && e[0] == '0'); }();
  const bool graphable = use_graph && g->use_tc && !profiling() && noise == nullptr && trace == nullptr &&
                         x_hat_out == nullptr && K > 1;
  if (!graphable) return issue(z, x, nullptr, s);
  const GenPack::GraphKey key = {B, K, with_noise, step_size, sigma, (unsigned long long)chain0, (unsigned long long)step0,
                                 ws.base, m ? m->uid : 0ull};
  const GenPack::GraphKey& k0 = g->gkey;
  const bool same = k0.B == key.B && k0.K == key.K && k0.with_noise == key.with_noise && k0.step == key.step &&
                    k0.sigma == key.sigma && k0.chain0 == key.chain0 && k0.step0 == key.step0 && k0.ws_base == key.ws_base &&
                    k0.ebm == key.ebm;
  if (!same) {   // new configuration: run it directly once; capture if it comes back
    if (g->gexec) { cudaGraphExecDestroy(g->gexec); g->gexec = nullptr; }
    g->gkey = key;
    return issue(z, x, nullptr, s);
  }
  if (!g->gexec) {
    if (!g->cap_stream) DAMC_CUDA(cudaStreamCreateWithFlags(&g->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    DAMC_CUDA(cudaStreamBeginCapture(g->cap_stream, cudaStreamCaptureModeThreadLocal));
    const long long n0 = g_launches.load();
    const int r = issue(ws.zbuf, ws.xbuf, ws.seed_dev, g->cap_stream);
    g->graph_launches = g_launches.load() - n0;
    g_launches -= g->graph_launches;   // nothing ran yet: replays are counted when they are launched
    const cudaError_t ce = cudaStreamEndCapture(g->cap_stream, &graph);
    if (r != DAMC_OK) { if (graph) cudaGraphDestroy(graph); return r; }
    if (ce != cudaSuccess || !graph) DAMC_FAIL(DAMC_ERR_CUDA, "posterior: stream capture failed: %s", cudaGetErrorString(ce));
    const cudaError_t ie = cudaGraphInstantiate(&g->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { g->gexec = nullptr; DAMC_FAIL(DAMC_ERR_CUDA, "posterior: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
  }
  const unsigned long long seed_host = seed;
  DAMC_CUDA(cudaMemcpyAsync(ws.seed_dev, &seed_host, sizeof(seed_host), cudaMemcpyHostToDevice, s));
  DAMC_CUDA(cudaMemcpyAsync(ws.zbuf, z, sizeof(float) * (size_t)B * g->nz, cudaMemcpyDeviceToDevice, s));
  DAMC_CUDA(cudaMemcpyAsync(ws.xbuf, x, sizeof(float) * (size_t)B * g->nc * g->H * g->W, cudaMemcpyDeviceToDevice, s));
  DAMC_CUDA(cudaGraphLaunch(g->gexec, s));
  count_launch((int)g->graph_launches);
  DAMC_CUDA(cudaMemcpyAsync(z, ws.zbuf, sizeof(float) * (size_t)B * g->nz, cudaMemcpyDeviceToDevice, s));
  return DAMC_OK;
}

}  // extern "C"
