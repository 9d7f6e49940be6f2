// damc_internal.h -- handle structs and cross-file launch prototypes (not part of the C ABI).
#pragma once
#include "damc_common.cuh"

unsigned long long damc_next_uid();
// one caller-owned weight tensor of a handle (for the change-detection hash of damc_repack)
struct HashSrc { const float* p; unsigned long long n, off; };
struct damc_handle {
  int kind;
  const unsigned long long uid = damc_next_uid();   // never reused: cached graphs key on it, not on the (recyclable) address
  // damc_repack change detection: a 64-bit hash of every source tensor is recomputed on the device (one launch, reads the
  // weights once) and compared with the hash of the values the packed buffers were built from; the packing kernels are
  // always enqueued but return at once when nothing changed (`dirty` flag in device memory) -- no host synchronisation.
  HashSrc* hash_tab = nullptr;            // device copy of the source table
  int hash_n = 0;
  unsigned long long* hash_state = nullptr;   // device: [0] accumulator, [1] hash of the packed values, [2] dirty flag, [3] ticket
  virtual ~damc_handle() {
    if (hash_tab) cudaFree(hash_tab);
    if (hash_state) cudaFree(hash_state);
  }
  // the caller's tensors this handle was packed from; empty = no change detection (refill always repacks)
  virtual void sources(std::vector<HashSrc>& out) const { (void)out; }
  // re-read the caller's weight tensors recorded at pack time into the packed buffers (async on stream).
  // dirty: null = unconditionally; else device flag -- every kernel of the refill returns immediately when *dirty == 0.
  virtual int refill(cudaStream_t stream, const int* dirty) = 0;
};

namespace damc {

struct EbmTcPack;   // 16-bit, zero-padded copies of the EBM weights for the tensor-core step kernel (ebm_tc.cu)
void ebm_tc_free(EbmTcPack* t);

// ---- EBM MLP (weights kept in the reference's Linear layouts; the persistent kernel re-tiles them into smem) -------
struct MlpPack : damc_handle {
  int nz = 0, ndf = 0;
  float slope = 0.2f;
  float *W1 = nullptr, *b1 = nullptr, *W2 = nullptr, *b2 = nullptr, *w3 = nullptr, *b3 = nullptr;
  float *W1T = nullptr, *W2T = nullptr;  // transposed copies [nz][ndf], [ndf][ndf] for the single-step kernel
  const float* src[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // caller's tensors (for damc_repack)
  float* slab = nullptr;
  mutable EbmTcPack* tcp[2] = {nullptr, nullptr};   // [bf16, fp16]: built on first use by launch_ebm_step_tc, kept fresh by refill()
  ~MlpPack() override {
    if (slab) cudaFree(slab);
    ebm_tc_free(tcp[0]);
    ebm_tc_free(tcp[1]);
  }
  int refill(cudaStream_t stream, const int* dirty) override;
  void sources(std::vector<HashSrc>& out) const override;
};

// change detection (damc_api.cu)
int handle_hash_init(damc_handle* h, cudaStream_t stream);   // after the first refill: record the hash of the packed values
int handle_repack(damc_handle* h, cudaStream_t stream);      // hash -> dirty flag -> gated refill
int launch_gated_copy(float* dst, const float* src, size_t n, const int* dirty, cudaStream_t stream);

int launch_ebm_langevin(const MlpPack* m, float* z, int B, int K, float step, int with_noise, const float* noise,
                        uint64_t seed, uint64_t chain0, uint64_t step0, float* trace, int trace_stride,
                        const float* gpart, int nsplit, int gstride, int nz_if_no_ebm, cudaStream_t stream);

// single Langevin step for many chains with the MLP streamed from L2 (the posterior sampler's per-step tail)
int launch_ebm_step(const MlpPack* m, float* z, int B, float step, int with_noise, const float* noise, uint64_t seed,
                    uint64_t chain0, uint64_t step_index, float* trace4, const float* gpart, int nsplit, int gstride,
                    float gpart_scale, int nz_if_no_ebm, cudaStream_t stream,
                    const unsigned long long* seed_ptr = nullptr /* non-null: {seed, chain0, step0} in device memory;
                                                                      chain0 / step_index arguments are then relative */);
int launch_transpose(const float* src, float* dst, int rows, int cols, cudaStream_t stream, const int* dirty = nullptr);
// the same step on the tensor cores (16-bit generator modes, no trace): one CTA per 128 chains, four tcgen05 GEMMs -- ebm_tc.cu
bool ebm_tc_usable(const MlpPack* m, int precision, const float* trace);
int ebm_tc_refill(const MlpPack* m, int precision, cudaStream_t s, const int* dirty);
int launch_ebm_step_tc(const MlpPack* m, int precision, float* z, int B, float step, int with_noise, const float* noise,
                       uint64_t seed, uint64_t chain0, uint64_t step_index, const float* gpart, int nsplit, int gstride,
                       float gpart_scale, cudaStream_t stream, const unsigned long long* seed_ptr, int K = 1);

// ---- generator as a chain of shifted-window GEMMs --------------------------------------------------------------
// Every layer's forward and input-gradient is   D[m,n] = sum_t sum_c A_t[m,c] * W_t[c,n]
// with m = (b,y,x) on an Hm x Wm grid, and A_t[m,:] = src[plane_t][b, y+dy_t, x+dx_t, :] (zero outside the grid).
enum LayerType { L_FIRST = 0, L_UP = 1, L_SAME = 2 };  // 1x1->kxk (s1,p0) | k4,s2,p1 | k3,s1,p1
enum EpiKind { EPI_FWD_ACT = 0, EPI_FWD_LAST = 1, EPI_DGRAD_MASK = 2, EPI_DGRAD_Z = 3, EPI_STORE_F32 = 4,
               EPI_DEN_LAYER = 5, EPI_DEN_FINAL = 6, EPI_STORE_F32_BIAS = 7 };  // 7: raw accumulators + bias[n] (fp32 rows)

struct Tap { signed char plane, dy, dx, pad; };

// DAMC denoiser layer on the tcgen05 engine: every N tile of BN columns holds the four terms of BN/4 output features as
// column blocks [gate | hyper-bias | main | skip]; the K axis is [layer input h (din) | ctx activation c (dout)], and an
// h k-block feeds only (main, skip), a c k-block only (gate, hyper-bias) -- half-width MMAs, no structural zeros.  The
// epilogue forms out = (main+b) sigmoid(gate+bg) + hb + skip + bs  (reference diffusion_net.py:439-445).
struct DenEpi {
  int din;                     // width of the h part of the K axis (multiple of 64)
  int bn;                      // N tile width the weight rows were packed for (256 | 128 | 64); must match the launch
  const float* bias4;          // [4*dout] per feature (bg, 0, b, bs)
  void* dst1; int ld1, off1;   // leaky_relu(out, .01) -> dst1[b*ld1 + off1 + n]   (operand type)
  void* dst2; int ld2, off2;   // second destination (U-net skip), or null
  // EPI_DEN_FINAL: eps = z + out ; reverse update of z (reference diffusion_net.py:610-620)
  float* z; float* eps_out; const float* noise;
  int nz, residual, last, use_philox;
  float c_pred, c_eps, c_zt, c_x, c_std;
  unsigned long long seed, chain0, step;
  const unsigned long long* seed_ptr;   // non-null: the Philox seed is read from device memory (replayed CUDA graphs)
};

struct Epilogue {
  int kind;
  void* out;            // activation / gradient tensor (element type T)
  const float* bias;    // [bias_mod] or null
  int bias_mod;
  long long o_b, o_y, o_x;  // EPI_FWD_ACT: out offset = b*o_b + (y*sy+py)*o_y + (x*sx+px)*o_x + n
  int sy, sx, py, px;
  float slope;
  // EPI_DGRAD_MASK: sign source = activation of the producing layer, NHWC on the M grid, C = N
  const void* act;
  int planar_out;       // 1: scatter (y,x) into 4 parity planes [4][B][Hm/2][Wm/2][N]; 0: flat [M][N]
  // tcgen05 engine only: 1 bit per activation element ("pre-activation > 0"), word index = NHWC element index / 32.
  // Written by EPI_FWD_ACT, read by EPI_DGRAD_MASK instead of the 16x larger activation rows.  null = use `act`.
  uint32_t* maskbits;
  // EPI_FWD_LAST
  const float* x;       // [B,nc,Ho,Wo] or null (forward only)
  float* xhat;          // [B,nc,Ho,Wo] or null
  float inv_sigma2;
  float gscale;         // factor carried by the stored gradient chain (sigma^2 in fp16 mode, else 1); undone in the update
  float* loss;          // null or scalar accumulator: sum (xhat-x)^2 * inv_sigma2/2
  void* gcol;           // im2col'd dL/dh_last for the last layer's dgrad: [B*Hi*Wi][64], slot (kh*k+kw)*4 + c
  int nc, k, stride, padding, Hi, Wi, Ho, Wo;
  DenEpi den;           // EPI_DEN_LAYER / EPI_DEN_FINAL
  // EPI_DGRAD_Z / EPI_STORE_F32
  int nz_out;           // row stride of the fp32 output: [split][B][nz_out] partial sums, or [M][nz_out] raw accumulators
};

struct GemmPlan {
  const void* A;            // element type T
  long long plane_stride;   // elements between source planes
  int B, Hm, Wm, Cs;        // M grid and K per tap (multiple of 64)
  int ntaps;
  Tap taps[16];
  const void* W;            // SIMT engine: [ntaps*Cs][Np] fp32/bf16 (N contiguous)
  const void* Wtc;          // tcgen05 engine: [ntaps][Np][Cs] bf16 (K contiguous), or null
  int N, Np;                // logical / padded (multiple of 16) output columns
  int ksplit;               // >1: grid.z splits the K loop (EPI_DGRAD_Z only)
  int op_fp16;              // tcgen05 engine: operands / stored tensors are fp16 (else bf16); set by launch_gemm_tc
  int op_f32;               // tcgen05 engine: operands / stored tensors are tf32-rounded fp32 (kind::tf32); set by launch_gemm_tc
  int ncls;                 // tcgen05 engine only: 4 = all output-parity classes of a k4-s2-p1 forward in ONE launch
                            // (taps and epilogue parity derived from the class; Wtc holds the 4 class blocks back to back)
  Epilogue epi;
};

struct GenLayer {
  int type, cin, cout, k, stride, pad, Hin, Win, Hout, Wout;
  int cin_p;  // cin padded to a multiple of 64 (first layer: nz)
  float* bias = nullptr;                 // device copy
  // packed weights: fwd (per parity class for L_UP: 4, else 1) and dgrad
  void* w_fwd[4] = {nullptr, nullptr, nullptr, nullptr};
  void* w_dgrad = nullptr;
  void* w_fwd_tc[4] = {nullptr, nullptr, nullptr, nullptr};  // K-major copies for the tcgen05 engine (bf16 mode)
  void* w_dgrad_tc = nullptr;
  // last layer, "scatter form": Y[m_in, (kh,kw,co)] = a[m_in,:] . W[:,co,kh,kw]  (one tap, N = k*k*nc), then col2im
  void* w_scatter = nullptr;
  void* w_scatter_tc = nullptr;
  int n_sc = 0, np_sc = 0;
  int n_fwd = 0, np_fwd = 0;    // N / padded N of the forward GEMM
  int n_dg = 0, np_dg = 0;      // N / padded N of the dgrad GEMM
};

struct GenPack : damc_handle {
  int precision = DAMC_PREC_FP32;
  int nlayers = 0, nz = 0, nz_p = 0, nc = 0, H = 0, W = 0;
  float slope = 0.2f;
  std::vector<GenLayer> layers;
  std::vector<damc_convt_layer> src;  // caller's tensors (for damc_repack)
  std::vector<void*> allocs;
  int dz_splits = 1;
  bool last_fused = false;    // posterior steps run the last layer (forward, likelihood gradient, dgrad) as ONE launch (gen_last.cu)
  bool last_scatter = false;  // last layer runs as scatter-form GEMM + per-image finish kernel (image fits in smem)
  bool use_bits = false;    // tcgen05 engine: LeakyReLU masks travel as 1-bit-per-element words
  bool use_tc = false;      // bf16 mode: tcgen05 engine (default) or the SIMT engine on bf16 storage (DAMC_TC=0)
  // The K-step launch sequence of the posterior sampler (K x ~11 kernels) for the last configuration seen twice, as a
  // CUDA graph: z, x and the Philox seed are staged through the caller's workspace, everything else it references is
  // the workspace or this handle / the EBM handle, so it is replayed while the key matches.
  // A small LRU of such graphs (a training loop alternates e.g. 128 training chains and 500 evaluation chains); the Philox
  // seed, chain offset and step offset are read from device memory, so they are NOT part of the key.
  struct GraphKey { int B, K, with_noise, want_xhat; float step, sigma; void* ws_base; unsigned long long ebm; };
  struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t gexec = nullptr;   // null until the key has been seen twice
    long long launches = 0;            // kernels in the captured sequence (for damc_launch_count on replays)
    unsigned long long last_use = 0;
  };
  static constexpr int kMaxGraphs = 4;
  mutable std::vector<GraphEntry> graphs;
  mutable unsigned long long graph_clock = 0;
  mutable cudaStream_t cap_stream = nullptr;
  mutable long long graph_replays = 0, graph_captures = 0;
  ~GenPack() override {
    for (GraphEntry& e : graphs) if (e.gexec) cudaGraphExecDestroy(e.gexec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (void* p : allocs) cudaFree(p);
  }
  int refill(cudaStream_t stream, const int* dirty) override;
  void sources(std::vector<HashSrc>& out) const override;
};

struct GenWorkspace {   // carved out of the caller's workspace for a given B
  void* zin;                 // [B][nz_p]          T
  std::vector<void*> act;    // a_1..a_{L-1}       T (NHWC)
  std::vector<void*> grad;   // g_1..g_{L-1}       T (planar for L_UP producers, flat for L_FIRST)
  std::vector<uint32_t*> mask;  // sign bits of a_1..a_{L-1} (tcgen05 engine) or null
  void* gcol;                // [B*Hi*Wi][64]      T
  float* ybuf;               // [B*Hi*Wi][np_sc]   fp32 (scatter-form last layer) or null
  float* dz_part;            // [splits][B][nz_p]  fp32
  float* zbuf;               // [B][nz] fp32: staging copy of z for graph replays
  float* xbuf;               // [B][nc][H][W] fp32: staging copy of x for graph replays
  float* xhat_buf;           // [B][nc][H][W] fp32: G(z) of the last step inside a replayed graph (copied out to the caller)
  float* sq_part;            // [B][score_parts] fp32: per-chain partial sums of (G(z) - x)^2 (damc_posterior_score)
  unsigned long long* seed_dev;   // {seed, chain0, step0} for replayed graphs
  void* base;
  size_t bytes;
};

size_t elem_size(int precision);
// 16-bit operand modes (2-byte storage; generator, denoiser and encoder engines)
static inline bool is_tc_precision(int precision) { return precision == DAMC_PREC_BF16 || precision == DAMC_PREC_FP16; }
// modes whose generator GEMMs run on the tcgen05 engine (DAMC_PREC_TF32: fp32 containers, kind::tf32 MMAs)
static inline bool is_gen_tc_precision(int precision) { return is_tc_precision(precision) || precision == DAMC_PREC_TF32; }
int plan_workspace(const GenPack* g, int B, void* base, GenWorkspace* ws);

// SIMT implicit GEMM (fp32 or bf16 storage, fp32 accumulate) -- gen_simt.cu
int launch_gemm_simt(const GemmPlan& p, int precision, cudaStream_t stream);
// tcgen05/TMEM/TMA implicit GEMM (bf16 operands, fp32 accumulate) -- gen_tc.cu
int launch_gemm_tc(const GemmPlan& p, int precision, cudaStream_t stream);
// prepared launches (tensor maps encoded once, replayed many times; the plan's epilogue scalars stay editable)
struct TcLaunch;
int tc_prepare(const GemmPlan& p, int precision, TcLaunch** out);
int tc_encode_2d(void* tensor_map /* CUtensorMap* */, int fp16, const void* base, int cols, int rows, int box_rows);
int tc_encode_act(void* tensor_map /* CUtensorMap* */, int precision, const void* base, int Cs, int Wm, int Hm, int B, int Ht);
int tc_den_tile_width(int B, int Np);   // N tile width convgemm will use for a denoiser layer with Np = 4*dout columns
GemmPlan* tc_plan(TcLaunch* l);
int tc_launch(TcLaunch* l, cudaStream_t stream);
void tc_free(TcLaunch* l);
int tc_available();
int tc_selftest_fastdiv();

// weight packing -- gen_pack.cu
enum PackMode { PK_FIRST_FWD, PK_FIRST_DGRAD, PK_UP_FWD, PK_UP_DGRAD, PK_SAME_FWD, PK_LAST_DGRAD_COL, PK_LAST_FWD_SCATTER };
int launch_pack_convt(const float* w, int cin, int cout, int k, int stride, int pad, int mode, int cls, int ntaps,
                      int Cs, int Np, int nk_layout, int precision, void* dst, cudaStream_t stream,
                      const int* dirty = nullptr);
size_t last_finish_smem(const GenLayer& y);
int launch_last_finish(const GenLayer& y, int precision, const float* Y, int B, const float* x, float* xhat,
                       float inv_sigma2, float gscale, float* loss, void* gcol, cudaStream_t stream);
int launch_stage_z(const float* z, void* zin, int B, int nz, int nz_p, int precision, cudaStream_t stream);

// fused last layer (scatter GEMM + col2im + tanh-likelihood gradient + K = 64 dgrad in one launch) -- gen_last.cu
bool last_fused_supported(const GenPack* g);
int last_fused_parts(const GenPack* g);   // blocks per image = partial sums per chain in score mode
// sq_part == null: posterior step (forward, likelihood gradient, dgrad).  sq_part != null: score mode -- forward and the
// per-block sums of (x_hat - x)^2 only ([B][last_fused_parts] floats)
int launch_last_fused(const GenPack* g, const GenWorkspace& ws, int B, const float* x, float sigma, float* xhat, float* loss,
                      float* sq_part, cudaStream_t stream);

// generator driver -- gen_driver.cu
int generator_forward(const GenPack* g, const GenWorkspace& ws, const float* z, int B, const float* x, float sigma,
                      float* xhat, float* loss, cudaStream_t stream, float* sq_part = nullptr /* fused last layer: score mode */);
float generator_grad_scale(const GenPack* g, float sigma);
int generator_dgrad(const GenPack* g, const GenWorkspace& ws, int B, cudaStream_t stream);
// G(z) with the per-chain squared residual against x reduced on the fly: sq_part [B][score_parts(g)] (eval consumers)
int generator_score_forward(const GenPack* g, const GenWorkspace& ws, const float* z, int B, const float* x, cudaStream_t stream);
int score_parts(const GenPack* g);
int launch_sqerr(const float* xhat, const float* x, int B, int n, float* sq_part, cudaStream_t stream);
int launch_ebm_score(const MlpPack* m, const float* z, int B, int nz, const float* sq_part, int nparts, float* score,
                     float* sqerr, cudaStream_t stream);

// ---- DAMC denoiser (Diffusion_UnetA, reference diffusion_net.py:463-533) ---------------------------------------------
constexpr int DEN_LAYERS = 7;
constexpr int DEN_TM = 16;        // chains per CTA (fp32 streaming kernel)
constexpr int DEN_THREADS = 256;  // = max dout
constexpr int DEN_MAXW = 512;     // widest layer input (concat of two 64*nf halves)

// tcgen05 operands of one precision (bf16 | fp16): per layer one K-major weight matrix [4*dout][din+dout] whose rows are
// grouped per N tile as [gate | hyper-bias | main | skip] blocks of BN/4 features; one copy per tile width in use.
constexpr int DEN_NBN = 3;   // tile widths 256, 128, 64
struct DenSeqPack;           // weights of the hoisted-context schedule (denoiser_seq.cu)
void den_seq_free(DenSeqPack* p);
struct DenTcPack {
  void* slab = nullptr;
  void* Wq[DEN_NBN][DEN_LAYERS];  // [bn index][layer]: [4*dout][din+dout] operand type
  bool live[DEN_NBN] = {false, false, false};   // packed (and kept fresh by refill) once a launch has asked for it
  float* bias4[DEN_LAYERS];       // [4*dout] (bg, 0, b, bs)
  // The T-step launch sequence (T x (1 + 7) kernels) of the last configuration seen twice, captured as a CUDA graph:
  // everything it references lives in the caller's workspace (z is staged through ws.zbuf, the seed through
  // ws.seed_dev) or in this handle, so it can be replayed as long as the key below matches.
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t cap_stream = nullptr;   // private stream the sequence is captured on (the caller's may be the legacy stream)
  struct GraphKey { int B, T, use_philox; void* ws_base; unsigned long long chain0, coef_hash; } gkey = {0, 0, 0, nullptr, 0, 0};
  int gkey_seen = 0;   // calls with gkey so far (the graph is captured on the second one)
  DenSeqPack* seq = nullptr;   // built on first use by den_seq_run, kept fresh by den_tc_refill
};

struct DenPack : damc_handle {
  int nz = 0, nxemb = 0, ntemb = 0, residual = 0, csum = 0;
  int din[DEN_LAYERS], dout[DEN_LAYERS], coff[DEN_LAYERS];
  float* slab = nullptr;
  // device pointers into slab
  float *tw1, *tb1, *tw2, *tb2, *Bp;          // time_mlp, B [nz][nz/2]
  float* WcT_t;                               // [ntemb][csum]   (transposed temb half of all ctx Linears)
  float* WcT_x;                               // [nxemb][csum]   (transposed xemb half)
  float* bc;                                  // [csum]
  float* Wms[DEN_LAYERS];                     // [din][dout][2]  interleaved (main, skip), transposed
  float* Wgb[DEN_LAYERS];                     // [dout][dout][2] interleaved (gate, hyper-bias), transposed
  float* bias3[DEN_LAYERS];                   // [3][dout]: b_main, b_skip, b_gate
  int rows_gb[DEN_LAYERS], rows_ms[DEN_LAYERS], R[DEN_LAYERS];  // padded streamed rows / rows per 32 KB chunk
  float* wstream = nullptr;                   // [Wgb_0 | Wms_0 | Wgb_1 | ...] in consumption order (inside slab)
  size_t stream_floats = 0;
  damc_denoiser_desc src;                     // caller's tensors (for damc_repack)
  mutable DenTcPack* tc[3] = {nullptr, nullptr, nullptr};  // indexed by DAMC_PREC_*; built on first use, refilled by refill()
  ~DenPack() override;
  int refill(cudaStream_t stream, const int* dirty) override;
  void sources(std::vector<HashSrc>& out) const override;
};

struct DenWs {   // carved out of the caller's workspace
  float *cx, *ct, *dlog, *coef;
  void* A[DEN_LAYERS];   // tcgen05 mode: layer operands [B][din+dout] (operand type); null in fp32 mode
  float* zbuf;           // tcgen05 mode: [B][nz] staging copy of z for graph replays
  unsigned long long* seed_dev;
  // hoisted-context schedule (denoiser_seq.cu): gate / hyper-bias words of a window of steps, U-net skips of layers 0 / 1
  void* G;               // [window][B^128][csum] half2
  void* skip[2];         // [B^128][dout] operand type
  void* nbuf;            // [window][B^128][nz] fp32 normals of the window's steps, tile-transposed
  void* zT;              // [B^128][nz] fp32 z between the steps of a window, tile-transposed
  void* xr;              // [B][nxemb] tf32(SiLU(xemb)), the operand of the cx GEMM
  void* base;
  size_t bytes;
};
DenWs den_ws(const DenPack* d, int B, int T, int precision, void* base);

// denoiser_tc.cu: the 7 layers of one reverse step as tcgen05 GEMMs (quad-column epilogue), all T steps
int den_tc_ensure(const DenPack* d, int precision, cudaStream_t stream);
int den_tc_refill(const DenPack* d, int precision, cudaStream_t stream, const int* dirty = nullptr);
void den_tc_free(DenTcPack* t);
// coef: host table [nsteps][8] in execution order (c_pred, c_eps, c_zt, c_x, c_std, last)
// denoiser_cluster.cu: all T reverse steps in ONE launch -- a 4-CTA cluster owns 128 chains, splits every layer's N tiles,
// and synchronises layers with the hardware cluster barrier (operand rows travel through L2)
bool den_cluster_supported(const DenPack* d, int B);
int den_tc_pack_bn(const DenPack* d, int precision, int variant, cudaStream_t stream);   // ensures the row order of tile width 256 / 128 / 64 (variant 0 / 1 / 2) is packed
int den_cluster_run(const DenPack* d, int precision, const DenWs& w, float* z, float* eps_out, int B, int T, int nsteps,
                    const float* host_coef, const float* noise, int use_philox, uint64_t seed, uint64_t chain0,
                    cudaStream_t stream);
int den_tc_run(const DenPack* d, int precision, const DenWs& w, float* z, float* eps_out, int B, int T, int nsteps,
               const float* host_coef, const float* noise, int use_philox, uint64_t seed, uint64_t chain0,
               cudaStream_t stream);
// denoiser_seq.cu: gate / hyper-bias of all layers hoisted out of the step loop (one parallel pass per window of steps), then ONE
// CTA per 128 chains runs every step of the window with the activations resident in shared memory
bool den_seq_shape_ok(const DenPack* d);    // the U-net widths the kernel's buffer plan is written for (sizes the workspace)
bool den_seq_supported(const DenPack* d);   // shape ok and not switched off (DAMC_DEN_SEQ=0)
int den_seq_window(int B, int T, int csum); // steps per window
bool den_seq_hoist_usable(const DenPack* d, int precision, int B);
int den_seq_hoist(const DenPack* d, int precision, const DenWs& w, const float* xemb, int B, cudaStream_t stream);   // cx on the tensor cores
int den_seq_refill(const DenPack* d, int precision, cudaStream_t stream, const int* dirty);
int den_seq_run(const DenPack* d, int precision, const DenWs& w, float* z, int B, int T, int nsteps, const float* host_coef,
                const float* noise, int use_philox, uint64_t seed, uint64_t chain0, cudaStream_t stream);

}  // namespace damc
