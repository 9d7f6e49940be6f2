/* damc.h -- C ABI of the B200-native sampling library (libdamc_b200.so).
 *
 * This is the drop-in boundary for the hot path named in BASELINE.json:north_star.  The reference has no FFI of its
 * own: its boundary is the Python call surface of workspace/src/MCMC.py and workspace/src/diffusion_net.py, so every
 * entry point below cites the reference interface it replaces.  The Python mirror of that surface
 * (diffusion-amortized-mcmc_b200/damc_b200/MCMC.py) binds these symbols with ctypes; INTEGRATION.md shows the stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless named host_*.
 *   - all tensors are contiguous fp32 in the reference's own layouts: z [B,nz] row-major, x [B,nc,H,W] (NCHW),
 *     ConvTranspose2d weight [Cin,Cout,kH,kW], Linear weight [out,in].
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and performs no host
 *     synchronisation.  Calls are re-entrant for distinct (handle, stream, workspace) triples; one handle must not be
 *     used by two host threads at the same time (it caches packed weights, prepared launches and a CUDA graph).
 *   - the caller owns every buffer including the workspace; handles own only their packed weights, until damc_free.
 *   - return value: 0 = ok, non-zero = error; damc_last_error() returns a thread-local message.
 *   - there is no CPU fallback and no backend dispatch: unsupported configurations are hard errors.
 */
#ifndef DAMC_H_
#define DAMC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct damc_handle damc_handle; /* opaque packed-weight handle */

enum { DAMC_OK = 0, DAMC_ERR_INVALID = 1, DAMC_ERR_UNSUPPORTED = 2, DAMC_ERR_CUDA = 3, DAMC_ERR_WORKSPACE = 4 };

/* arithmetic of the GEMM operands -- generator (chosen at pack time) and DAMC denoiser (chosen per call).  The EBM, the
 * Langevin / reverse-step updates, z itself and every accumulator are always fp32. */
enum {
  DAMC_PREC_FP32 = 0, /* fp32 CUDA-core implicit GEMM: exact fp32 FMA accumulation (cross-check mode)     */
  DAMC_PREC_BF16 = 1, /* bf16 operands, fp32 accumulate on tcgen05/TMEM fed by TMA: the throughput mode */
  DAMC_PREC_FP16 = 2, /* same engine with fp16 operands (8x finer operand rounding than bf16, same speed); the
                         gradient chain is carried scaled by sigma^2 so that it stays inside the fp16 range */
  DAMC_PREC_TF32 = 3  /* generator only: the same tcgen05 engine with fp32 tensors rounded to tf32 (10-bit mantissa,
                         fp32 range) and kind::tf32 MMAs -- the arithmetic torch/cuDNN use by default for the reference's
                         convolutions on a GPU (src/MCMC.py:55-60); the rel-1e-3 parity mode and the library default */
};

int damc_version(void);
const char* damc_last_error(void);
int damc_free(damc_handle* h);
/* Re-read the caller's weight tensors (the device pointers given at pack time, which must still be alive) into the
 * packed buffers -- asynchronous on `stream`, no allocation.  The Python mirror calls it before every sampler call, so
 * in-place parameter updates (optimizer steps, EMA copies through .data as in train_gen_recon.py:258-261) are seen. */
int damc_repack(damc_handle* h, void* stream);

/* ---- EBM prior  (replaces netE.ebm : nn.Sequential(Linear,LeakyReLU(.2),Linear,LeakyReLU(.2),Linear),
 *                  reference src/diffusion_net.py:207-223) ------------------------------------------------------- */
int damc_pack_mlp(damc_handle** out, int nz, int ndf, const float* W1, const float* b1, const float* W2,
                  const float* b2, const float* W3, const float* b3, float negative_slope, void* stream);

/* ---- generator  (replaces netG.gen : ConvTranspose2d / LeakyReLU(.2) / Tanh stack,
 *                  reference src/diffusion_net.py:20-203) -------------------------------------------------------- */
typedef struct {
  int cin, cout, k, stride, pad;
  const float* weight; /* [cin,cout,k,k] */
  const float* bias;   /* [cout] or NULL */
} damc_convt_layer;

int damc_pack_generator(damc_handle** out, int nlayers, const damc_convt_layer* host_layers, float negative_slope,
                        int precision, void* stream);
/* output geometry of a packed generator */
int damc_generator_shape(const damc_handle* gen, int* nz, int* nc, int* height, int* width);
size_t damc_generator_workspace_bytes(const damc_handle* gen, int B);

/* x_hat = G(z)  (reference: netG.forward, src/diffusion_net.py:49-51; used by gen_samples, src/MCMC.py:126-127) */
int damc_generator_forward(const damc_handle* gen, const float* z, float* x_hat, int B, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- eval consumers right after the posterior sampler (reference eval_anomaly_det.py:114-119, eval_gen_recon.py:192-194) --
 * One pass: x_hat = G(z) is formed tile by tile and reduced against x where it is produced (it never reaches HBM in the
 * 16-bit tensor-core modes); E(z) and |z|^2/2 come from one small kernel -- no separate netG / netE forward.
 *   sqerr[b] = sum_{c,h,w} (G(z_b) - x_b)^2                 (recon MSE = sqerr / (nc H W), eval_gen_recon.py:193)
 *   score[b] = sqerr[b] + E(z_b) + |z_b|^2 / 2              (anomaly score, eval_anomaly_det.py:117; E = 0 when ebm is NULL)
 * score / sqerr: [B] device pointers, either may be NULL.  Same workspace as damc_posterior_langevin.                 */
int damc_posterior_score(const damc_handle* gen, const damc_handle* ebm, const float* z, const float* x, int B,
                         float* score, float* sqerr, void* workspace, size_t workspace_bytes, void* stream);

/* ---- prior Langevin  (replaces sample_langevin_prior_z, reference src/MCMC.py:27-46) --------------------------
 * z [B,nz] is updated IN PLACE:  z <- z - step^2/2 (dE/dz + z) + step * eps  for K steps in ONE launch.
 * noise : NULL -> Philox4x32-10 keyed by (seed, chain0 + row, step0 + i); else injected normals [K,B,nz].
 * trace : NULL, or [K,2] receiving (sum_b E_b, |z|^2/2) evaluated at the pre-update z of every step (:40-41).   */
int damc_prior_langevin(const damc_handle* ebm, float* z, int B, int K, float step_size, int with_noise,
                        const float* noise, uint64_t seed, uint64_t chain0, uint64_t step0, float* trace,
                        void* stream);

/* The same K-step loop with the MLP's four mat-mat products on the tensor cores (fp16 operands, fp32 accumulation and update),
 * one CTA per 128 chains, all K steps in one launch: the large-batch form (16 384 chains x 60 steps: 1.x ms vs 15.7 ms), at
 * 16-bit accuracy of dE/dz (csrc/ebm_tc.cu).  No trace.  nz % 4 == 0, nz <= 128, ndf <= 256.                                    */
int damc_prior_langevin_tc(const damc_handle* ebm, float* z, int B, int K, float step_size, int with_noise, const float* noise,
                           uint64_t seed, uint64_t chain0, uint64_t step0, void* stream);

/* ---- posterior Langevin  (replaces sample_langevin_post_z_with_prior, reference src/MCMC.py:48-74) -------------
 * U(z) = |G(z)-x|^2/(2 sigma^2) + E(z) + |z|^2/2 ;  ebm may be NULL (toy-style target without the EBM term).
 * trace : NULL, or [K,4] receiving (sum E, |G(z)-x|^2/(2 sigma^2), |z|^2/2, mean(grad)) per step (:65-67).
 * x_hat_out : NULL, or [B,nc,H,W] receiving G(z) of the LAST step's pre-update z.                                */
int damc_posterior_langevin(const damc_handle* gen, const damc_handle* ebm, float* z, const float* x, int B, int K,
                            float step_size, float sigma, int with_noise, const float* noise, uint64_t seed,
                            uint64_t chain0, uint64_t step0, float* trace, float* x_hat_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- toy posterior  (replaces the closure sample_langevin_post_z, reference toy_example/toy_example.py:110-131;
 *                      G = ReLU MLP nz-nh-nh-nh-nx, :22-47) ----------------------------------------------------- */
int damc_pack_toy_mlp(damc_handle** out, int nz, int nh, int nx, const float* const* host_W /*4*/,
                      const float* const* host_b /*4*/, void* stream);
int damc_toy_posterior_langevin(const damc_handle* mlp, float* z, const float* x, int B, int K, float step_size,
                                float sigma, int with_noise, const float* noise, uint64_t seed, uint64_t chain0,
                                uint64_t step0, void* stream);

/* ---- DAMC denoiser  (replaces the reverse loop of _netQ_U.forward, reference src/diffusion_net.py:595-622, with
 *                      Q.p = Diffusion_UnetA :463-533 and the helpers src/diffusion_helper_func.py:36-70) ---------- */
typedef struct {
  int nz, nxemb, ntemb, nf, residual;
  const float* time_w1; const float* time_b1; const float* time_w2; const float* time_b2; /* time_mlp.{1,3} */
  const float* Bproj;                                                                      /* p.B [nz,nz/2]  */
  /* 7 ConcatSquashLinearSkipCtx layers in execution order: in0,in1,in2,mid0,out0,out1,out2 */
  int dim_in[7], dim_out[7];
  const float* W[7];  const float* b[7];   /* _layer.0         [out,in]          */
  const float* Wc[7]; const float* bc[7];  /* _layer_ctx.1     [out,ntemb+nxemb] */
  const float* Wg[7]; const float* bg[7];  /* _hyper_gate      [out,out]         */
  const float* Wb[7];                      /* _hyper_bias      [out,out]         */
  const float* Ws[7]; const float* bs[7];  /* _skip            [out,in]          */
} damc_denoiser_desc;

int damc_pack_denoiser(damc_handle** out, const damc_denoiser_desc* host_desc, void* stream);
size_t damc_denoise_workspace_bytes(const damc_handle* den, int B, int T, int precision);

/* z [B,nz] holds z_T on entry and z_0 on return; xemb [B,nxemb] is encoder(x) or prior_emb(randn).
 * host_logsnr[T+1]: lambda(t_i) for i = 0..T-1 then unused; computed by the caller exactly as the reference's
 *   logsnr_schedule_fn (fp32) so the time embedding sees the same argument.
 * var_type: 0 = 'small', 1 = 'large'.  noise: NULL -> Philox, else [T-1,B,nz] consumed in execution order.
 * The reference draws eps even when with_noise is false (:616); with injected noise that is immaterial.
 * precision: DAMC_PREC_FP32 = all T steps in ONE persistent fp32 kernel (z resident in shared memory, weights streamed
 *   by cp.async.bulk); DAMC_PREC_BF16 / DAMC_PREC_FP16 = per step one operand-preparation kernel + 7 tcgen05/TMEM GEMMs
 *   fed by TMA (each ConcatSquashLinearSkipCtx layer is one GEMM with a fused gate/bias/skip epilogue; the last one also
 *   applies the reverse update).  Layer widths must be multiples of 64 in the tensor-core modes.                    */
int damc_denoise(const damc_handle* den, float* z, const float* xemb, int B, int T, const float* host_logsnr,
                 int var_type, int with_noise, const float* noise, uint64_t seed, uint64_t chain0, int precision,
                 void* workspace, size_t workspace_bytes, void* stream);

/* single eps-prediction of Q.p (reference src/diffusion_net.py:501-533) -- used by the per-step parity tests */
int damc_denoiser_eps(const damc_handle* den, const float* z, const float* xemb, float logsnr, float* eps_out, int B,
                      int precision, void* workspace, size_t workspace_bytes, void* stream);

/* ---- image encoder of the amortizer  (replaces Q.encoder(x), reference src/diffusion_net.py:590 with Encoder_cifar10
 *      :227-266 / Encoder_celeba64 :268-313 / Encoder_celebaHQ :315-372: Conv2d(nc,nif,3,1,1), Conv2d(.,.,4,2,1) x n,
 *      Conv2d(.,nemb,k,1,0) on the final k x k map; InstanceNorm2d(affine) + LeakyReLU after every layer but the last).
 *      The k4-s2-p1 layers run as the generator engines' stride-2 dgrad GEMMs (tcgen05 in the bf16/fp16 modes, CUDA-core
 *      fp32 in DAMC_PREC_FP32).  Odd-sized maps (the 28x28 MNIST encoder) are not supported: hard error. ---------------- */
typedef struct {
  int cin, cout, k, stride, pad;
  const float* weight;    /* Conv2d weight [cout,cin,k,k]                     */
  const float* bias;      /* [cout] or NULL                                   */
  const float* in_weight; /* InstanceNorm2d weight [cout]; NULL on the last layer */
  const float* in_bias;   /* InstanceNorm2d bias   [cout]; NULL on the last layer */
} damc_conv_layer;

int damc_pack_encoder(damc_handle** out, int nlayers, const damc_conv_layer* host_layers, int height, int width,
                      float negative_slope, float eps, int precision, void* stream);
size_t damc_encoder_workspace_bytes(const damc_handle* enc, int B);
/* xemb [B,nemb] = encoder(x [B,nc,H,W]) */
int damc_encoder_forward(const damc_handle* enc, const float* x, float* xemb, int B, void* workspace,
                         size_t workspace_bytes, void* stream);

/* ---- measurement hooks (bench.py; no reference counterpart) ------------------------------------------------------
 * damc_launch_count : cumulative number of kernels this library has launched in the calling process.
 * damc_profile_enable(1) : bracket every generator GEMM launch with a cudaEvent pair on its own stream;
 * damc_profile_collect : synchronise those events, return their summed duration (ms) and count, and reset.        */
/* damc_selftest : host-only consistency checks of the library's index arithmetic (no GPU needed); 0 = ok. */
int damc_selftest(void);
long long damc_launch_count(void);
/* number of posterior calls on this generator handle served by replaying a captured CUDA graph (-1: not a generator) */
long long damc_graph_replays(const damc_handle* gen);
int damc_profile_enable(int on);
int damc_profile_collect(double* gemm_ms, long long* gemm_launches);

/* ---- optimiser half of the training step that calls the samplers (reference train_gen_recon.py:218-219, :229-230, :239-240:
 * clip_grad_norm_(max_norm) followed by Adam / AdamW .step(), optimisers of :152-154) over ONE flat fp32 buffer ---------
 * grads are read as grad_scale * g (grad_scale = 1 / world_size after a SUM all-reduce of the flat buffer), their global L2
 * norm is reduced in a fixed order, and the update is applied to the listed element ranges only (parameters that received
 * no gradient this step are skipped, as torch skips .grad = None), each with its own 1-based step count.
 *   decoupled = 1: AdamW (p *= 1 - lr wd), 0: Adam (g += wd p).  max_norm <= 0: no clipping.
 *   scratch: >= 1025 floats of device memory; scratch[1024] receives the total gradient norm (what clip_grad_norm_ returns). */
int damc_fused_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n_total, int nranges,
                         const unsigned long long* range_begin, const unsigned long long* range_count, const int* range_step,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, float max_norm,
                         float grad_scale, float* scratch, void* stream);

/* ---- plain GEMM on the tcgen05 engine, TF32 operands / fp32 accumulate: D[m][n] = sum_k A[m][k] W[n][k] (+ bias[n]) ------------
 * A [M][K] and W [N][K] row-major fp32 (W is nn.Linear's weight layout), D [M][ldd].  K % 32 == 0, N % 16 == 0.
 * Used by damc_b200.denoiser_train for the Linear layers of Q.p in Q.calculate_loss (reference diffusion_net.py:417-445,
 * :624-646): forward X W^T, input-gradient dY W, weight-gradient dY^T X are all this one form on (transposed) copies.     */
int damc_gemm_tf32(const float* A, const float* W, const float* bias, float* D, int M, int N, int K, int ldd, void* stream);
/* dst[i] = src[i] rounded to the nearest TF32 value (cvt.rna), fp32 container; dst may equal src.  The MMA above truncates its
 * operands; rounding them first (as cuBLAS' TF32 path does) halves the perturbation of damc_gemm_tf32.                          */
int damc_round_tf32(const float* src, float* dst, size_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAMC_H_ */
