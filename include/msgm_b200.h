/* msgm_b200.h -- C ABI of libmsgm_b200.so: the B200 (sm_100a) replacement for the data-parallel hot path of
 * vressegu/sdeflow-light (MSGM): reverse/forward SDE sampling loops and the sliced-score-matching train step.
 *
 * The reference has no FFI of its own (it is pure Python); its "plugin" boundary is the duck-typed Python
 * surface of sde_scheme.py / SDEs.py / NN.py.  Each entry point below names the reference interface it
 * replaces (file:line under the reference root).  The Python shims in sdeflow_light_b200/ bind these with
 * ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every pointer marked "device" is a CUDA device pointer owned by the caller; the library owns
 * only the opaque msgm_ctx (small device workspace).  All calls are stream-ordered on `stream` (a cudaStream_t
 * passed as void*), never synchronise the host, and are re-entrant per ctx.  Return 0 on success, a negative
 * msgm_status otherwise; msgm_last_error() gives a thread-local message.  No C++ exception crosses the ABI.
 */
#ifndef MSGM_B200_H
#define MSGM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSGM_ABI_VERSION 1
#define MSGM_HIDDEN 128   /* NN.py:77 hidden_dim (the only width the reference driver uses) */
#define MSGM_MAX_DIM_MLP 32

typedef enum {
  MSGM_OK = 0,
  MSGM_ERR_INVALID = -1,     /* bad argument (maps to ValueError in the Python shim) */
  MSGM_ERR_UNSUPPORTED = -2, /* valid in the reference, not built here (maps to NotImplementedError) */
  MSGM_ERR_CUDA = -3,        /* a CUDA runtime call failed */
  MSGM_ERR_NO_DEVICE = -4    /* no sm_100 device: there is NO CPU fallback */
} msgm_status;

typedef enum { MSGM_SDE_SGM = 0, MSGM_SDE_MSGM_DENSE = 1, MSGM_SDE_MSGM_SPARSE = 2 } msgm_sde_kind;
/* Hutchinson probe distributions of sample_v (SDEs.py:514-536): 'rademacher', 'normal'/'gaussian', 'uniform' (sphere) */
typedef enum { MSGM_V_RADEMACHER = 0, MSGM_V_GAUSSIAN = 1, MSGM_V_SPHERE = 2 } msgm_vtype;
typedef enum { MSGM_SCHEME_EM = 0, MSGM_SCHEME_HEUN = 1, MSGM_SCHEME_RK4 = 2 } msgm_scheme;
typedef enum {
  MSGM_PREC_FP32 = 0, /* CUDA-core fp32 everywhere: the parity mode (reference is fp32 end to end) */
  MSGM_PREC_F16TC = 1 /* tcgen05 tensor cores, fp16 operands / fp32 accumulate for the 128-wide layers */
} msgm_precision;

/* Base SDE coefficients.  Replaces SGMsde / MSGMsde.{beta,f,f_strato,div_Sigma,g,IJK}
 * (SDEs.py:72-73,183-194,401-432).  The sparse cyclic tensor (SDEs.py:369-399) needs no arrays: it is the
 * 3-point stencil  (g w)_i = c sqrt(beta) (y_{i+1} w_i - y_{i-1} w_{i-1}),  c = sqrt(2)/2. */
typedef struct {
  int32_t kind;       /* msgm_sde_kind */
  int32_t dim;        /* state dimension d */
  float beta_min;     /* SDEs.py:58 */
  float beta_delta;   /* (float)(beta_max - beta_min), difference taken in double like the reference */
  float T;            /* horizon, SDEs.py:57 */
  const float* G;     /* device (d,d,d) row-major [i][j][k]; MSGM dense only (SDEs.py:315-341) */
  const float* L_G;   /* device (d,d); MSGM dense only; Ito correction (SDEs.py:246) */
} msgm_sde_desc;

/* NN.MLP weights in torch.nn.Linear layout (NN.py:98-106): W[l] is (out,in) row-major, b[l] is (out,).
 * Layer sizes: (d + premodule + 1) -> 128 -> 128 -> 128 -> d.  All device pointers, fp32. */
typedef struct {
  int32_t input_dim;
  int32_t premodule;  /* 1 = NormalizeLogRadius (NN.py:56-70), 0 = none */
  const float* W[4];
  const float* b[4];
} msgm_mlp_desc;

/* One sampler call.  Replaces euler_maruyama_sampler / heun_sampler / rk4_stratonovich_sampler
 * (sde_scheme.py:43-99,101-172,174-269) driving PluginReverseSDE.{mu,mu_Strato,sigma} (SDEs.py:556-588) or,
 * with forward_only=1, the forward_SDE adapter (SDEs.py:30-47; `mlp` may then be NULL). */
typedef struct {
  int32_t scheme;          /* msgm_scheme */
  int32_t num_steps;       /* N */
  float lmbd;              /* lambda family, SDEs.py:561,584,588 */
  int32_t norm_correction; /* re-pin |x| to its initial value after every step (sde_scheme.py:254-255) */
  int32_t include_t0;      /* trajectory slot 0 holds x_0 (sde_scheme.py:207-212) */
  int32_t forward_only;    /* integrate the forward (noising) SDE instead of the reverse one */
  int32_t precision;       /* msgm_precision */
  float T_;                /* integration horizon override (sde_scheme.py:196-199); < 0 = sde.T */
  const float* ts;         /* device (N+1,) time grid = linspace(0,1,N+1)*T_ as the caller computed it
                              (sde_scheme.py:201); NULL = computed in-kernel as i*(T_/N) */
  const float* noise;      /* device (N,B,d) standard normals, one draw per step shared by all stages
                              (sde_scheme.py:227); NULL = in-kernel Philox4x32-10 keyed
                              (seed, particle_offset + row, step) so results do not depend on sharding */
  uint64_t seed;
  uint64_t particle_offset;
  float* traj;             /* device (N + include_t0, B, d) or NULL (keep_all_samples, sde_scheme.py:257) */
  const int32_t* keep_step;/* device (B,) or NULL: samplesToKeep (sde_scheme.py:259-262) */
  float* keep_out;         /* device (B,d), caller-zeroed; row k written when step index == keep_step[k] */
  const float* T_rows;     /* device (B,) or NULL: per-particle horizon T_ (batched form of the reference's
                              one-sample calls with T_=t[k], SDEs.py:114-116); `ts` then holds the UNIT grid
                              linspace(0,1,N+1) and row k uses ts[i]*T_rows[k] */
} msgm_sample_args;

typedef struct msgm_ctx msgm_ctx;

int msgm_abi_version(void);
const char* msgm_last_error(void);
/* Create / destroy the per-GPU context.  Fails with MSGM_ERR_NO_DEVICE when `device` is not sm_100. */
int msgm_create(msgm_ctx** out, int device);
int msgm_destroy(msgm_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
int64_t msgm_launch_count(const msgm_ctx* ctx);
/* Error word of the tensor-core kernels (sampler, conv, attention): 1 = a bounded mbarrier wait timed out and the
 * launch gave up (its output is undefined), 2 = shared-memory / TMEM base assumption violated; 0 in a healthy run.  The
 * kernels write the word into mapped pinned host memory, so msgm_async_error reads it WITHOUT synchronising: it reports
 * whatever has been raised by launches that already ran (the Python shims call it on entry and after every device->host
 * copy and raise RuntimeError).  msgm_debug_flags synchronises the device first.  Both are read-and-clear.
 * (The reference has no counterpart: a CUDA fault in its ATen ops surfaces as a RuntimeError of the next call.) */
int msgm_debug_flags(msgm_ctx* ctx, int32_t* out_host);
int msgm_async_error(msgm_ctx* ctx, int32_t* out_host);
/* Debug: cycle counters of the tensor-core sampler (CTA 0), filled only while the environment variable
 * MSGM_TC_PROF is set; n <= 24.  [0..7] owner thread, [8..15] MMA thread, [16..23] helper thread. */
int msgm_debug_counters(msgm_ctx* ctx, int64_t* out_host, int n);

/* Whole sampling loop for an MLP score net, all N steps in one persistent launch; x_inout (B,d) device fp32
 * holds x_0 on entry and x_N on exit. */
int msgm_sample_mlp(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp,
                    const msgm_sample_args* args, float* x_inout, int64_t B, void* stream);

/* Forward noising for training, whole batch in ONE launch.  Replaces SDE.sample_scheme / MSGMsde.sample
 * (SDEs.py:78-122,434-436): row k is integrated (RK4-Stratonovich, forward SDE, no radius correction) for
 * n_k = trunc(N t_k / T) steps of size T/N and stops there; a row with n_k == 0 takes one step of size t_k instead.
 * Dense tensor: d <= 32; sparse tensor: d <= 4096 (one CTA per row: the U-Net configurations, NNUnet1D / NNUnet sizes).
 * t (B,) device; y_inout (B,d) holds x on entry and y_t on exit.  noise (N,B,d) / noise_single (B,d) inject the normals
 * of the common grid / of the one-step rows; NULL = in-kernel Philox keyed by (seed, particle_offset + row). */
int msgm_noise_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* t, float* y_inout, int32_t num_steps_forward,
                       const float* ts /* (N+1,) grid linspace(0,1,N+1)*T as the caller computed it, or NULL */,
                       const float* noise, const float* noise_single, uint64_t seed, uint64_t particle_offset, int64_t B,
                       void* stream);

/* Training prologue in ONE launch.  Replaces PluginReverseSDE.sample_txy + sample_t + sample_v (SDEs.py:648-693,
 * 514-536) and the forward noising they call: for every row draws t ~ U(0,T) floored at t_epsilon, the probe v, and
 * y_t | x (MSGM: the simulation of msgm_noise_forward; SGM: the closed-form marginal of SDEs.py:134-146).
 * x (B,d) in; t_out (B,), v_out (B,d), y_out (B,d) out; all device fp32.  Every draw is Philox keyed by
 * (seed + *seed_offset_dev, sample_offset + row): seed_offset_dev (device uint64, may be NULL) lets a replayed CUDA
 * graph see a fresh stream per iteration; sample_offset makes a batch sharded over ranks draw what one rank would. */
int msgm_ssm_prepare(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* x, float* t_out, float* v_out, float* y_out,
                     int32_t num_steps_forward, const float* ts, float t_epsilon, int32_t vtype, uint64_t seed,
                     const uint64_t* seed_offset_dev, uint64_t sample_offset, int64_t B, void* stream);

/* Stand-alone score-net forward a(y, s) -> (B,d) for NN.MLP.forward (NN.py:108-120); s is (B,). */
int msgm_mlp_forward(msgm_ctx* ctx, const msgm_mlp_desc* mlp, const float* y, const float* s, float* out,
                     int64_t B, void* stream);

/* Sliced-score-matching train step for the MLP score net.  Replaces PluginReverseSDE.ssm_loss (SDEs.py:616-646):
 *   loss[b] = v_b^T d/dy[ mu_to_div(y_b) ] v_b + |a(y_b, t_b)|^2 / 2
 * evaluated in forward mode (primal + tangent through the net) instead of autograd's VJP, and differentiated by hand
 * instead of double backward.  y (B,d) are the noised samples, v (B,d) the Hutchinson probes (SDEs.py:514-536), t (B,)
 * the noise times.  `scratch` is caller-owned device memory of msgm_ssm_scratch_bytes(B) bytes that carries the
 * activations from the forward to the backward call.  The backward call takes the upstream gradient of the per-sample
 * loss, grad_out (B,) (1/B for `.mean().backward()`), and writes d(sum_b grad_out[b] loss[b])/d(theta) into grad_flat in
 * torch parameter order [W0 (128,d+1+pre), b0, W1, b1, W2, b2, W3 (d,128), b3]. */
uint64_t msgm_ssm_scratch_bytes(int64_t B);
int msgm_ssm_mlp_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                         const float* v, const float* t, float* loss_out, void* scratch, int64_t B, void* stream);
int msgm_ssm_mlp_backward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                          const float* v, const float* t, const float* grad_out, void* scratch, float* grad_flat,
                          int64_t B, void* stream);

/* The same training step on tcgen05 tensor cores (fp16 operands, fp32 accumulation), ONE launch for loss + activation
 * backward + every weight gradient (csrc/ssm_tc.cu): loss_out[b] as msgm_ssm_mlp_forward, grad_flat_out = gradient of
 * sum_b grad_out[b] loss[b] in torch parameter order as msgm_ssm_mlp_backward.  d <= 16.  scratch: device memory of
 * msgm_ssm_tc_scratch_bytes(ctx, d, premodule, B) bytes (per-CTA partial gradients, summed by a second small launch).
 * cot_scale: a power of two that brings grad_out to O(1) (B for the 1/B of a batch mean): the cotangents are fp16
 * tensor-core operands; the scale is applied inside the kernel and removed from the result.
 * Stated tolerance vs the fp32 kernels / the reference: loss 2e-3 of max|loss|, gradients 3e-3 of max|g| per tensor.
 * Replaces PluginReverseSDE.ssm + loss.backward() (SDEs.py:607-646; MSGM_higherDim.py:803-809). */
uint64_t msgm_ssm_tc_scratch_bytes(const msgm_ctx* ctx, int32_t d, int32_t premodule, int64_t B);
int msgm_ssm_mlp_fwd_bwd_tc(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                            const float* v, const float* t, const float* grad_out, float* loss_out, float* grad_flat_out,
                            void* scratch, float cot_scale, int64_t B, void* stream);

/* Per-stage SDE update for score nets that are not fused into the sampler (U-Nets, d up to 4096; SGM and sparse MSGM).
 * One call per Runge-Kutta stage replaces EMstep + PluginReverseSDE.{mu_Strato,sigma} + the stage bookkeeping
 * (sde_scheme.py:18-40,223-255; SDEs.py:556-588).  `a` is the score-net output at (stage input, s); x (B,d) is the state
 * at the start of the step and receives the new state on the last stage (with the radius re-pinned to r0 when
 * norm_correction); y (B,d) is a separate buffer holding the stage input of stages > 0; ks (B,d) the running RK sum. */
int msgm_stage_update(msgm_ctx* ctx, const msgm_sde_desc* sde, int32_t scheme, int32_t stage, float lmbd,
                      int32_t norm_correction, int32_t forward_only, float s, float delta, const float* a,
                      const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B, void* stream);
/* Device-side step clock of the per-stage sampler.  With it every step-dependent value of the loop sde_scheme.py:223-255 is
 * read from device memory, so ONE captured CUDA graph (noise draw, nstage net evaluations, nstage stage updates, clock advance)
 * is replayed for all N steps instead of ~130 host launches per step: clock[0] = index i of the current step,
 * s_table[i * nstage + st] = noise time of stage st of step i (N * nstage + 1 entries, the last one unused). */
typedef struct msgm_step_clock {
  const int32_t* clock;      /* device (1,) */
  const float* s_table;      /* device (N * nstage + 1,) */
  float* s_next;             /* device (B,) or NULL: receives, per row, the noise time of the NEXT net evaluation */
  float* traj;               /* device (N + include_t0, B, d) or NULL: the new state lands at [clock + include_t0] (last stage) */
  const int32_t* keep_step;  /* device (B,) per-row step to keep, with keep_out (samplesToKeep, sde_scheme.py:101-103,258-260) */
  float* keep_out;           /* device (B, d) or NULL */
  int32_t include_t0;
  int32_t reserved;
} msgm_step_clock;
/* msgm_stage_update with the stage time taken from clk (stage input / output as above);
 * msgm_philox_normal_clocked: out = scale * N(0,1) keyed by the clock's step, or scale * noise[clock] when injected noise
 * (N,B,d) is given; also writes the time of the step's first net evaluation to clk->s_next;
 * msgm_clock_advance: clock[0] += 1 (last node of the step graph). */
int msgm_stage_update_clocked(msgm_ctx* ctx, const msgm_sde_desc* sde, int32_t scheme, int32_t stage, float lmbd,
                              int32_t norm_correction, int32_t forward_only, const msgm_step_clock* clk, float delta,
                              const float* a, const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B,
                              void* stream);
int msgm_philox_normal_clocked(msgm_ctx* ctx, float* out, int32_t d, int64_t B, float scale, uint64_t seed,
                               uint64_t particle_offset, const msgm_step_clock* clk, int32_t nstage,
                               const float* noise_or_null, void* stream);
int msgm_clock_advance(msgm_ctx* ctx, int32_t* clock, void* stream);
/* r[b] = |x[b,:]|  (torch.norm(x_t, dim=1), sde_scheme.py:66,124,205). */
int msgm_row_norm(msgm_ctx* ctx, const float* x, float* r, int32_t d, int64_t B, void* stream);
/* out (B,d) = scale * N(0,1) with the samplers' Philox keying (seed, particle_offset + row, step). */
int msgm_philox_normal(msgm_ctx* ctx, float* out, int32_t d, int64_t B, float scale, uint64_t seed,
                       uint64_t particle_offset, uint32_t step, void* stream);

/* Prior sampling.  Replaces MSGMsde.latent_sample / gen_radial_distribution / randu_on_sphere (SDEs.py:438-493,520-526;
 * msgm=1: x0 = quantile(r_T, U) [exp(.) - 1e-6 if log_map] * z/|z|) and SGMsde.latent_sample (SDEs.py:201-203; msgm=0:
 * x0 = z).  rT_sorted: device (n_r,) ascending.  U (B,) / Z (B,d) inject the uniform / normal draws (parity tests); NULL
 * = in-kernel Philox keyed by (seed, particle_offset + row). */
int msgm_latent_sample(msgm_ctx* ctx, const float* rT_sorted, int32_t n_r, int32_t log_map, int32_t msgm, const float* U,
                       const float* Z, float* out, int32_t d, int64_t B, uint64_t seed, uint64_t particle_offset,
                       void* stream);
/* MMD metric.  Replaces compute_kernel / compute_mmd (quantitative_comparison.py:22-47): sums_out (device, 3 doubles)
 * receives the pair sums of exp(-|a-b|^2/d^2) over (x,x), (y,y), (x,y); mmd = s0/N^2 + s1/M^2 - 2 s2/(N M). */
int msgm_mmd_sums(msgm_ctx* ctx, const float* x, int64_t N, const float* y, int64_t M, int32_t d, double* sums_out,
                  void* stream);

/* Sample-quality metrics of the reference's evaluation step, on device-resident particles.
 * msgm_row_norm_stats: norms_out[b] = |x_b * scale| (scale_opt: (d,) per-dimension factor or NULL;
 *   own_plotting.py:646-653,729-736) and minpos_max_out = {smallest positive norm, largest norm} (two floats), the span of
 *   the shared radius grid (_compute_common_R_grid, own_plotting.py:616-632).
 * msgm_survival_counts: counts_out[g] = #{b : norms[b] > R_grid[g]} for an ascending float64 grid (n_grid <= 4096), i.e.
 *   norms.size - searchsorted(sort(norms), R_grid, side='right') of _empirical_survival_from_norms (own_plotting.py:635-640);
 *   scratch: 8 (n_grid + 1) bytes of device memory.  The tail exponent of _tail_fit_loglog (own_plotting.py:656-700) needs
 *   no sort either: R_g >= the (n-k-1)-th order statistic  <=>  counts[g] <= k.
 * msgm_moments: colsum_out (d doubles) = sum_b x_b, gram_out (d x d doubles, upper 32x32 tiles filled) = sum_b x_b x_b^T:
 *   mean, torch.cov, torch.var and the energy mean|x|^2 of preprocessing() (own_plotting.py:339-394). */
int msgm_row_norm_stats(msgm_ctx* ctx, const float* x, const float* scale_opt, float* norms_out, float* minpos_max_out,
                        int32_t d, int64_t n, void* stream);
int msgm_survival_counts(msgm_ctx* ctx, const float* norms, int64_t n, const double* R_grid, int32_t n_grid,
                         int64_t* counts_out, void* scratch, void* stream);
int msgm_moments(msgm_ctx* ctx, const float* x, int64_t n, int32_t d, double* colsum_out, double* gram_out, void* stream);

/* Adam update (torch.optim.Adam(lr) as the reference driver builds it, MSGM_higherDim.py:792: betas (0.9, 0.999),
 * eps 1e-8, no weight decay) over the trainer's flat gradient buffer in one launch.  seg_table: device array of n_tensors
 * {float* param; int64 begin} records (16 bytes each, ascending begin) mapping flat ranges to the parameter tensors;
 * exp_avg / exp_avg_sq: flat moment buffers (total floats); lr_dev, step_dev: device scalars (the kernel uses step+1 and
 * stores it back), so the launch replays inside a CUDA graph; grad_scale multiplies the gradient (1/world after a summed
 * all-reduce). */
int msgm_adam_step(msgm_ctx* ctx, const void* seg_table, int32_t n_tensors, int64_t total, const float* grad_flat,
                   float* exp_avg, float* exp_avg_sq, const float* lr_dev, int64_t* step_dev, float beta1, float beta2,
                   float eps, float grad_scale, void* stream);

/* ---- U-Net score nets: building blocks of the hand-written training path (csrc/unet_train.cu) ---------------------------------
 * The SSM loss of a U-Net (PluginReverseSDE.ssm_loss, SDEs.py:616-646, over NNUnet1D.py:110-179 / model/unet.py:101-250) is
 * evaluated in forward mode: every activation is a PAIR stacked along the batch axis, samples [0,B) primal, [B,2B) tangent.
 * Convolutions / Linear layers are linear, so the inference entry points above (msgm_conv1d[_tc], msgm_conv2d_tc, ...) run
 * them on the 2B-sample tensors (and, with flipped / transposed weights, their data gradients); these are the remaining ops.
 * They replace what torch autograd + cuDNN did for the reference (F.gelu / SiLU and their double backward, conv weight
 * gradients, nn.Linear, NormalizeLogRadius NN.py:56-70) on this path.
 *   msgm_pair_act       act 0 = exact GELU, 1 = SiLU.  grad_h NULL: out = (phi(z); phi'(z) zdot).  grad_h given: out =
 *                       (hbar phi' + hdotbar phi'' zdot; hdotbar phi').  z, out, grad_h: 2 * half_elems floats.
 *   msgm_amax / msgm_pow2_scale  max|x| into a device float; y = x * 2^k with k such that max|x| lands at 2^target_exp
 *                       (inverse: 2^-k).  Brackets the tensor-core data-gradient convs: deep-layer cotangents (~1e-7) sit in
 *                       the fp16 subnormal range where the hi/lo operand split of msgm_conv*_tc would lose its bits.
 *   msgm_rows_bias_add  x[n][c][p] += bias[c] for n < nrows (bias acts on the primal half only).
 *   msgm_channel_sums   out[c] += sum_{n < nrows, p} x[n][c][p] (bias gradients).
 *   msgm_tap_sums_1d    E[n][co][k] = sum of cot[n][co][p] over the output positions p whose tap k lands inside [0, Lin): the
 *                       cotangent of the folded embedding table of msgm_emb_fold.
 *   msgm_conv_wgrad     gW[co][coff + ci][ky][kx] += sum_{n,oy,ox} cot[n][co][oy][ox] in[n][ci][oy s + ky - pad][ox s + kx - pad]
 *                       (in = channel concat [in1, in2], optionally nearest-upsampled x2; 1-D: Hi = Ho = KH = 1), all N samples.
 *   msgm_gemm_f32       C = [C +] op(A) op(B) (row-major, fp32): the nn.Linear layers of the embedding MLPs and table folds.
 *   msgm_premodule_pair xn (2B,d) = scale * (x / (|x| + 1e-6); its tangent along v), logn (2B) = (log(|x| + 1e-6); tangent).
 *   msgm_sparse_ssm_loss  gout NULL: out[b] = q . adot + |a|^2/2 (+ beta |v|^2/2, SGM) from the net output pair a_pair
 *                       (2B,d), q from the cyclic sparse tensor (SDEs.py:369-399) or sqrt(beta) v (SGM); gout given: out =
 *                       the output cotangent pair (gout a; gout q), (2B,d). */
/*   msgm_bgemm_f32     msgm_gemm_f32 over `batch` problems (pointer strides) with a scalar factor: the products of
 *                       QKVAttention (model/unet.py:236-250) on a pair and of its backward.
 *   msgm_gn_pair        GroupNorm32 (model/nn_utils.py:39-46,107-114) on a pair x (2B,C,HW): grad_y NULL: out = (y; ydot),
 *                       stats (B,G,4) written; grad_y given: out = cotangents of (x; xdot), ggamma / gbeta accumulated.
 *   msgm_softmax_pair   A NULL: out1 = softmax(S) rows, out2 = Pdot = P (Sdot - sum P Sdot); A given (first argument = P):
 *                       out1 = Sbar, out2 = Sdotbar from the direct cotangents A = dL/dP, Pdotbar = dL/dPdot.
 *   msgm_sincos_pair    timestep_embedding (model/nn_utils.py:130-148) of a pair of scalars (value; tangent) -> (2B, dim).
 *   msgm_resample2      mode 0: zero-stuffing to twice the size (adjoint of a stride-2 conv's sub-sampling); mode 1: 2x2 block
 *                       sums (adjoint of the nearest-neighbour upsampling of Upsample, model/unet.py:40-60). */
/* A group of small batched products of ONE shape in one launch: problem i computes, for every matrix b < batch,
 *   C_i[b] = [C_i[b] +] alpha_i * sum_{s < nseg_i} op(A_i,s[b]) op(B_i,s[b])      (M x N, contraction length K)
 * -- the attention of the U-Net training path (QKVAttention on primal / tangent pairs, model/unet.py:236-250, and its adjoint)
 * is four such groups instead of 21 separate launches.  Problems of a group must not write overlapping outputs. */
#define MSGM_GEMM_MAX_PROBLEMS 6
typedef struct msgm_gemm_problem {
  const float* A[2];
  const float* B[2];
  float* C;
  int64_t stride_a[2], stride_b[2], stride_c;  /* element strides between the matrices of the batch */
  int32_t lda[2], ldb[2], ldc;
  int32_t trans_a[2], trans_b[2];              /* 1: stored transposed (A as K x M, B as N x K) */
  int32_t nseg;                                /* 1 or 2 products summed */
  int32_t accumulate;                          /* 1: add to C */
  float alpha;
  int32_t reserved;
} msgm_gemm_problem;
int msgm_gemm_group_f32(msgm_ctx* ctx, const msgm_gemm_problem* problems, int32_t n_problems, int32_t M, int32_t N, int32_t K,
                        int32_t batch, void* stream);
int msgm_bgemm_f32(msgm_ctx* ctx, const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t lda,
                   int32_t ldb, int32_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c, int32_t batch, int32_t trans_a,
                   int32_t trans_b, float alpha, int32_t accumulate, void* stream);
int msgm_gn_pair(msgm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* stats, const float* grad_y_or_null,
                 float* out, float* ggamma, float* gbeta, int32_t B, int32_t C, int32_t G, int32_t HW, void* stream);
int msgm_softmax_pair(msgm_ctx* ctx, const float* S_or_P, const float* Sdot, const float* A_or_null, const float* Pdotbar_or_null,
                      float* out1, float* out2, int64_t nrows, int32_t T, void* stream);
int msgm_sincos_pair(msgm_ctx* ctx, const float* val_pair, float* emb_pair, int32_t B, int32_t dim, void* stream);
int msgm_resample2(msgm_ctx* ctx, const float* x, float* out, int64_t NC, int32_t H, int32_t W, int32_t mode, void* stream);
/* Weight image of the DATA-GRADIENT conv of a stride-1 "same" convolution, packed straight from the forward weight W
 * (Cout, Cw, taps), taps = 9 (3x3) / 3 (1-D k3) / 1, first Cin input channels used: the adjoint is the same conv with the channel
 * roles swapped and the taps flipped, Wd[ci][co][taps-1-t] = W[co][ci][t], so the image has Cin output and Cout input channels and
 * is consumed by msgm_conv2d_tc / msgm_conv1d_tc like any other (size: msgm_conv2d_tc_pack_bytes / msgm_conv1d_tc_pack_bytes of
 * the swapped shape).  Replaces the transposed / flipped weight copy torch's convolution backward makes. */
int msgm_conv_tc_pack_dgrad(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cw, int32_t Cin, int32_t taps, void* wimg,
                            void* stream);
/* Range scaling of the NEXT msgm_conv2d_tc / msgm_conv1d_tc / msgm_convt1d_tc launch of this context (one-shot: the launch
 * consumes it): the kernel stages its input times the power of two that brings amax (device word, max|input| as written by
 * msgm_amax) into [2^14, 2^15) and scales its accumulators back before bias / residual terms.  Used by the data gradients of the U-Net
 * training path, whose cotangents (~1e-7 in the deep layers) would otherwise fall into the fp16 subnormal range of the split
 * operands.  NULL clears a pending request. */
int msgm_tc_range_scale(msgm_ctx* ctx, const float* amax_or_null);
/* msgm_conv_wgrad on tcgen05 (csrc/conv_wgrad_tc.cu): the weight gradient as a product over positions with both operands read
 * MN-major from the forward conv's staged tile layout, split fp16 x 3 (fp32-level parity), cotangent range-scaled by the power of
 * two derived from amax (device word written by msgm_amax; NULL = no scaling; amax_in does the same for the input operand, which
 * is the small one when the call computes a ConvTranspose weight gradient: there the "input" is the cotangent).  Takes stride-1 convolutions with "same" padding
 * -- 3x3 (KH = KW = 3, pad 1), 1-D k3 (KH = 1, KW = 3, pad 1) and 1x1, up in {1, 2} -- and the stride-2 convolutions of the
 * U-Nets -- 1-D k4 pad 1 (KH = 1, KW = 4; also ConvTranspose1d(k4, s2, p1) with the two tensors' roles swapped) and 3x3 pad 1 on
 * even sizes -- with channel counts % 16 == 0 (C1 % 16 == 0);
 * msgm_conv_wgrad_tc_ok tells (1 / 0) whether a shape is taken; otherwise the call returns MSGM_ERR_UNSUPPORTED and the caller
 * uses msgm_conv_wgrad.  (Hs, Ws) is the stored input size; cot is (N, Cout, Hs up, Ws up) for stride 1, (N, Cout, Hs / 2 [or 1], Ws / 2)
 * for stride 2.  scratch: device buffer of
 * msgm_conv_wgrad_tc_scratch_bytes(...) bytes (Cin = C1 + C2) for the per-slice partial tiles, summed by a second launch in a
 * fixed order (no atomics); accumulate = 0 overwrites the addressed block of gW instead of adding to it. */
int msgm_conv_wgrad_tc_ok(int32_t N, int32_t Cout, int32_t C1, int32_t C2, int32_t KH, int32_t KW, int32_t stride, int32_t pad,
                          int32_t up, int32_t Hs, int32_t Ws);
uint64_t msgm_conv_wgrad_tc_scratch_bytes(const msgm_ctx* ctx, int32_t N, int32_t Cout, int32_t Cin, int32_t KH, int32_t KW,
                                          int32_t stride, int32_t pad, int32_t up, int32_t Hs, int32_t Ws);
int msgm_conv_wgrad_tc(msgm_ctx* ctx, const float* cot, const float* in1, const float* in2, float* gW_accumulate,
                       const float* amax_or_null, const float* amax_in_or_null, void* scratch, int32_t N, int32_t Cout, int32_t C1, int32_t C2, int32_t Cw, int32_t coff,
                       int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t up, int32_t Hs, int32_t Ws, int32_t accumulate,
                       void* stream);
int msgm_pair_act(msgm_ctx* ctx, const float* z, const float* grad_h_or_null, float* out, int64_t half_elems, int32_t act,
                  void* stream);
int msgm_amax(msgm_ctx* ctx, const float* x, int64_t n, float* amax_out, void* stream);
/* Both range words of a convolution's backward in one launch: amax_out2[0] = max |cot|, amax_out2[1] = max |[x1, x2]| (x2 may be
 * NULL): the data gradient scales its input by the first, the weight gradient scales both operands. */
int msgm_amax2(msgm_ctx* ctx, const float* cot, int64_t n_cot, const float* x1, int64_t n1, const float* x2_or_null, int64_t n2,
               float* amax_out2, void* stream);
int msgm_pow2_scale(msgm_ctx* ctx, const float* x, float* y, int64_t n, const float* amax_dev, int32_t target_exp,
                    int32_t inverse, void* stream);
int msgm_rows_bias_add(msgm_ctx* ctx, float* x, const float* bias, int64_t nrows, int32_t C, int64_t P, void* stream);
int msgm_channel_sums(msgm_ctx* ctx, const float* x, float* out_accumulate, int64_t nrows, int32_t C, int64_t P, void* stream);
int msgm_tap_sums_1d(msgm_ctx* ctx, const float* cot, float* E, int64_t N, int32_t Cout, int32_t K, int32_t stride, int32_t pad,
                     int32_t Lin, int32_t Lout, void* stream);
int msgm_conv_wgrad(msgm_ctx* ctx, const float* cot, const float* in1, const float* in2, float* gW_accumulate, int32_t N,
                    int32_t Cout, int32_t C1, int32_t C2, int32_t Cw, int32_t coff, int32_t KH, int32_t KW, int32_t stride,
                    int32_t pad, int32_t up, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, void* stream);
int msgm_gemm_f32(msgm_ctx* ctx, const float* A, const float* B, float* C, int32_t M, int32_t N, int32_t K, int32_t lda,
                  int32_t ldb, int32_t ldc, int32_t trans_a, int32_t trans_b, int32_t accumulate, void* stream);
int msgm_premodule_pair(msgm_ctx* ctx, const float* x, const float* v, float* xn_pair, float* logn_pair, int64_t B, int32_t d,
                        float scale, void* stream);
int msgm_sparse_ssm_loss(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* a_pair, const float* y, const float* v,
                         const float* t, const float* gout_or_null, float* out, int64_t B, void* stream);

/* Gradient all-reduce fused with the Adam update over NVLink peer memory (csrc/p2p.cu; one node, one process per GPU):
 * what `dist.all_reduce(flat); flat /= world; optim.step()` does in a data-parallel run of the reference's loop
 * (MSGM_higherDim.py:803-809), as two launches without a collective-library call.  Every rank pushes its flat gradient into
 * a receive slot of every rank, raises a sequence flag there, and each rank's Adam kernel waits for the world's flags, sums
 * the slots in rank order (bit-identical on all ranks) and updates.
 *   msgm_p2p_create   allocates this rank's receive buffer (cudaMalloc) for gradients of up to nfloats floats and writes its
 *                     64-byte CUDA IPC handle to handle_out; the caller all-gathers the handles (torch.distributed);
 *   msgm_p2p_connect  all_handles: world x 64 bytes in rank order; opens the peers' buffers;
 *   msgm_p2p_disconnect / msgm_p2p_destroy  teardown: every rank closes its peer mappings, the ranks synchronise, then
 *                     each frees its own buffer (freeing memory a peer still maps can block);
 *   msgm_p2p_allreduce_adam  arguments as msgm_adam_step (grad_scale is 1/world); grad_flat must be readable up to the
 *                     next multiple of 4 floats.  Stream-ordered and CUDA-graph capturable; every rank must issue the same
 *                     sequence of calls.  A peer that never arrives raises error code 3 (msgm_async_error) after ~2 s. */
typedef struct msgm_p2p msgm_p2p;
int msgm_p2p_create(msgm_ctx* ctx, int64_t nfloats, int32_t world, int32_t rank, msgm_p2p** out, unsigned char* handle_out);
int msgm_p2p_connect(msgm_ctx* ctx, msgm_p2p* h, const unsigned char* all_handles);
int msgm_p2p_disconnect(msgm_p2p* h);  /* closes the peers' buffers; call on every rank, then synchronise the ranks, then destroy */
int msgm_p2p_destroy(msgm_p2p* h);
int msgm_p2p_allreduce_adam(msgm_ctx* ctx, msgm_p2p* h, const void* seg_table, int32_t n_tensors, int64_t total,
                            const float* grad_flat, float* exp_avg, float* exp_avg_sq, const float* lr_dev, int64_t* step_dev,
                            float beta1, float beta2, float eps, void* stream);

/* Log density of a 1-D Gaussian kernel density estimate at m query points: what MSGMsde.log_latent_pdf and the
 * normalising-constant estimate of the constructor ask sklearn's KernelDensity.score_samples for (SDEs.py:240,261,509;
 * kernel='gaussian', exact sum).  samples (n,), queries (m,), out (m,): device fp32. */
int msgm_kde_logpdf(msgm_ctx* ctx, const float* samples, int32_t n, float bandwidth, const float* queries, float* out,
                    int32_t m, void* stream);

/* ---- 1-D U-Net score net layers (NNUnet1D.py:13-179), fp32, NCL layout -----------------------------------------------
 * One msgm_conv1d call = nn.Conv1d (+ optional exact GELU) over the channel concatenation [x1, x2, emb] WITHOUT building
 * it: x2 (decoder skip) may be NULL; the Cemb embedding channels, constant along the signal, enter through the folded
 * table E (B,Cout,K) produced by msgm_emb_fold (NULL when Cemb == 0).  W is the module's weight (Cout, C1+C2+Cemb, K). */
typedef struct {
  const float* x1; const float* x2; const float* W; const float* bias; const float* E; float* out;
  int32_t B, C1, C2, Cemb, Cout, K, stride, pad, Lin, Lout, gelu;
} msgm_conv1d_desc;
int msgm_conv1d(msgm_ctx* ctx, const msgm_conv1d_desc* desc, void* stream);
/* E[b,co,k] = sum_ci W[co, Coff+ci, k] emb[b,ci] : the embedding channels of a conv block folded into a bias table. */
int msgm_emb_fold(msgm_ctx* ctx, const float* W, const float* emb, float* E, int32_t Cw, int32_t Coff, int32_t Cemb,
                  int32_t Cout, int32_t K, int32_t B, void* stream);
/* nn.ConvTranspose1d(Cin, Cout, kernel_size=4, stride=2, padding=1) followed by right zero padding to Lout
 * (NNUnet1D.py:98,165-169).  W is (Cin, Cout, 4). */
int msgm_convt1d_k4s2(msgm_ctx* ctx, const float* x, const float* W, const float* bias, float* out, int32_t B, int32_t Cin,
                      int32_t Cout, int32_t Lin, int32_t Lout, void* stream);
/* out (B,E) (+)= Linear(E,E)(GELU(Linear(1,E)(t))) : time_mlp / scale_embed (NNUnet1D.py:52-68); E <= 256. */
int msgm_embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2,
                   float* out, int32_t B, int32_t E, int32_t accumulate, void* stream);
/* xn = x / (|x| + 1e-6) * sqrt(L), lognorm = log(|x| + 1e-6) per row (NN.py:64-70, NNUnet1D.py:136-139). */
int msgm_normalize_log_radius(msgm_ctx* ctx, const float* x, float* xn, float* lognorm, int32_t B, int32_t L, void* stream);

/* ---- 2-D U-Net score net layers (NNUnet.py:80-245, model/unet.py:40-517), fp32, NCHW ---------------------------------
 * msgm_conv2d = nn.Conv2d 3x3 (padding 1, stride 1|2) or 1x1 over the channel concat [x1, x2] read in place, with
 *   prologue 0: none | 1: GroupNorm (stats from msgm_gn_stats, affine gamma/beta) | 2: GroupNorm + SiLU applied while
 *   the input tile is staged; up = 2 folds Upsample's nearest interpolation (model/unet.py:57-64) into the indexing;
 *   epilogue adds bias[co], ebias[b,co] (ResBlock's embedding term, model/unet.py:181-192) and the residual tensor res. */
typedef struct {
  const float* x1; const float* x2; const float* W; const float* bias; const float* ebias; const float* res;
  const float* stats; const float* gamma; const float* beta; float* out;
  int32_t B, C1, C2, Cout, K, stride, up, Hs, Ws, G, prologue;
} msgm_conv2d_desc;
int msgm_conv2d(msgm_ctx* ctx, const msgm_conv2d_desc* desc, void* stream);
/* Tensor-core form of msgm_conv2d (tcgen05, fp16 x 3 split operands, fp32 accumulate: fp32-level parity): 3x3 (padding 1)
 * or 1x1, stride 1|2, C1 % 16 == 0, (C1 + C2) % 16 == 0, Cout % 32 == 0.  `wimg` is the packed weight image written by
 * msgm_conv2d_tc_pack (msgm_conv2d_tc_pack_bytes bytes; pack once per weight tensor).  `ss` (B, C1 + C2, 2) is the
 * GroupNorm of the input folded into a per-(sample, channel) scale / shift by msgm_gn_scale_shift (NULL: no norm);
 * prologue 1 = normalise, 2 = normalise + SiLU, applied while the input tile is staged. */
typedef struct {
  const float* x1; const float* x2; const void* wimg; const float* bias; const float* ebias; const float* res;
  const float* ss; float* out;
  int32_t B, C1, C2, Cout, K, stride, up, Hs, Ws, prologue;
  int32_t fast; /* 0: three split products (fp32-level parity); 1: one fp16 product, ~1e-3 relative (sampling only) */
} msgm_conv2d_tc_desc;
int msgm_conv2d_tc(msgm_ctx* ctx, const msgm_conv2d_tc_desc* desc, void* stream);
int64_t msgm_conv2d_tc_pack_bytes(int32_t Cout, int32_t Cin, int32_t K);
int msgm_conv2d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cin, int32_t K, void* wimg, void* stream);
/* Tensor-core form of msgm_conv1d for the 1-D U-Net (NNUnet1D.py:13-33,81-102): nn.Conv1d k3 (stride 1, padding 1),
 * k4 (stride 2, padding 1: the down-sampling conv, evaluated at every position and kept at the even ones) or k1, over
 * the concat [x1, x2]; (C1 + C2) % 16 == 0, C1 % 16 == 0, Cout % 32 == 0.  W (Cout, Cw, K) may carry Cw - (C1 + C2)
 * trailing embedding channels: they are not packed, their contribution comes in as the folded table E (B, Cout, K) of
 * msgm_emb_fold and is added per tap where that tap reads inside the signal.  gelu = 1 applies the exact GELU. */
typedef struct {
  const float* x1; const float* x2; const void* wimg; const float* bias; const float* E; float* out;
  int32_t B, C1, C2, Cout, K, stride, Lin, gelu;
  int32_t fast; /* as in msgm_conv2d_tc_desc */
} msgm_conv1d_tc_desc;
int msgm_conv1d_tc(msgm_ctx* ctx, const msgm_conv1d_tc_desc* desc, void* stream);
int64_t msgm_conv1d_tc_pack_bytes(int32_t Cout, int32_t Cin, int32_t K);
int msgm_conv1d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cw, int32_t Cin, int32_t K, void* wimg,
                        void* stream);
/* Tensor-core form of msgm_convt1d_k4s2: the transposed conv is a 3-tap conv with 2 Cout output columns (even | odd
 * outputs).  Cin % 16 == 0, Cout % 16 == 0.  W (Cin, Cout, 4) is packed by msgm_convt1d_tc_pack into 24 Cin Cout bytes.
 * `out` (B, Cout, Lout >= 2 Lin) must be zero-filled by the caller when Lout > 2 Lin (the reference's right padding). */
int msgm_convt1d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cin, void* wimg, void* stream);
int msgm_convt1d_tc(msgm_ctx* ctx, const float* x, const void* wimg, const float* bias, float* out, int32_t B, int32_t Cin,
                    int32_t Cout, int32_t Lin, int32_t Lout, int32_t fast, void* stream);
/* msgm_emb_fold for up to 16 convs of one forward in ONE launch (the tables depend on the embedding vector only): entry i
 * is msgm_emb_fold(W[i], emb, E[i], Cw[i], Coff[i], Cemb, Cout[i], K[i], B); results are bit-identical. */
typedef struct {
  const float* W[16]; float* E[16];
  int32_t Cw[16], Coff[16], Cout[16], K[16];
  int32_t n, Cemb, B;
  const float* emb;
} msgm_emb_fold_multi_desc;
int msgm_emb_fold_multi(msgm_ctx* ctx, const msgm_emb_fold_multi_desc* desc, void* stream);
/* The time MLP and (u != NULL) the log-radius MLP of UNet1D in one launch: out (B,E) = MLP_a(t) [+ MLP_b(u)], as
 * msgm_embed_mlp(.., accumulate = 0) followed by msgm_embed_mlp(.., accumulate = 1), bit-identical; E <= 224 (the whole
 * second-layer weight sits in shared memory). */
int msgm_embed_mlp2(msgm_ctx* ctx, const float* t, const float* W1a, const float* b1a, const float* W2a, const float* b2a,
                    const float* u, const float* W1b, const float* b1b, const float* W2b, const float* b2b, float* out,
                    int32_t B, int32_t E, void* stream);
/* ---- TMA-fed 1-D convs on activation "planes" (csrc/conv1d_tcp.cu; NNUnet1D.py:13-33,81-102,165-169) ------------------
 * Between the convs of the 1-D U-Net (conv -> GELU -> conv -> GELU, no normalisation) an activation (B, C, L), C % 8 == 0,
 * is kept in the tensor cores' operand format instead of fp32 NCL:
 *     planes = fp16 [hi | lo][C / 8][R][8],  R = 2 * 640 + B (L + 3) rows of 16 bytes,
 *     row of (b, l) = 640 + b (L + 3) + 1 + l,  value = hi + lo (fp16 split of the fp32 value, ~22 mantissa bits).
 * Rows that hold no position (one left / two right of every signal, 640 guard rows at both ends) must be ZERO: allocate a
 * planes buffer zero-filled (msgm_planes_bytes) -- the kernels below only ever write rows of real positions.
 * msgm_conv1d_tcp is msgm_conv1d_tc / msgm_convt1d_tc with planes in and planes and / or fp32 NCL out: the operand tiles
 * are fetched by cp.async.bulk, products and epilogues of consecutive tiles overlap (persistent CTAs, double-buffered
 * TMEM accumulators).  K = 3 (stride 1, padding 1), K = 4 (stride 2, padding 1) or transposed = 1 (ConvTranspose1d k4 s2
 * p1, image of msgm_convt1d_tc_pack, K = 3, Lout >= 2 Lin given; rows beyond 2 Lin stay zero = the reference's right
 * padding).  x2 / E / bias / out_planes / out_f32 may be NULL (at least one output).  Shapes as msgm_conv1d_tc.
 * A planes buffer must not be the input and the output of one call. */
typedef struct {
  const void* x1; const void* x2; const void* wimg; const float* bias; const float* E; void* out_planes; float* out_f32;
  int32_t B, C1, C2, Cout, K, Lin, Lout, gelu, transposed;
  int32_t fast; /* as in msgm_conv2d_tc_desc */
} msgm_conv1d_tcp_desc;
int64_t msgm_planes_bytes(int64_t B, int32_t C, int32_t L);
int msgm_conv1d_tcp(msgm_ctx* ctx, const msgm_conv1d_tcp_desc* desc, void* stream);
/* fp32 (B, C, L) -> planes (rows of real positions only) and back (hi + lo). */
int msgm_planes_pack(msgm_ctx* ctx, const float* x, void* planes, int32_t B, int32_t C, int32_t L, void* stream);
int msgm_planes_unpack(msgm_ctx* ctx, const void* planes, float* x, int32_t B, int32_t C, int32_t L, void* stream);
/* First conv of the 1-D U-Net (one real input channel + folded embedding table E (B, Cout, 3), k3 p1, W (Cout, Cw, 3)
 * of which input channel 0 is used, optional exact GELU), written as planes.  Cout % 8 == 0, Cout <= 128. */
int msgm_conv1d_first_planes(msgm_ctx* ctx, const float* x, const float* W, int32_t Cw, const float* bias, const float* E,
                             void* planes, int32_t B, int32_t Cout, int32_t L, int32_t gelu, void* stream);
/* ss[b, c] = (rstd gamma_c, beta_c - mean rstd gamma_c) with the GroupNorm32 statistics of [x1, x2] (eps 1e-5). */
int msgm_gn_scale_shift(msgm_ctx* ctx, const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t HW, int32_t G,
                        int32_t B, const float* gamma, const float* beta, float* ss, void* stream);
/* GroupNorm32 statistics (model/nn_utils.py:39-41,107-114): stats (B,G,2) = mean, 1/sqrt(var + 1e-5) of [x1, x2]. */
int msgm_gn_stats(msgm_ctx* ctx, const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t HW, int32_t G, int32_t B,
                  float* stats, void* stream);
/* out (B,Cout) = Linear(SiLU(emb)) : ResBlock.emb_layers. */
int msgm_emb_proj(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int32_t E, int32_t Cout,
                  int32_t B, void* stream);
/* Every ResBlock's emb_layers of one forward in one launch: W (Ctot,E) / bias (Ctot) are the blocks' Linear layers stacked
 * along the output dimension, seg_start[0..nseg] (host array, nseg <= 64) the row offsets of the blocks; block i's result
 * is the contiguous (B, len_i) matrix at out + B * seg_start[i] (the `ebias` operand of that block's conv). */
int msgm_emb_proj_multi(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int32_t E,
                        int32_t Ctot, int32_t B, int32_t nseg, const int32_t* seg_start, void* stream);
/* out (B,E) (+)= Linear(E,E)(SiLU(Linear(dim,E)(timestep_embedding(t, dim)))) : time_embed / scale_embed. */
int msgm_sincos_embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2,
                          float* out, int32_t B, int32_t dim, int32_t E, int32_t accumulate, void* stream);
/* QKVAttention, single head: qkv (B,3C,T) -> out (B,C,T) = v softmax(q^T k / sqrt(C))^T (model/unet.py:236-250). */
int msgm_attention(msgm_ctx* ctx, const float* qkv, float* out, int32_t B, int32_t C, int32_t T, void* stream);
/* Tensor-core form of msgm_attention (tcgen05, fp16 x 3 split operands: fp32-level parity) for C in {32,64,96,128},
 * T in {64,128,192,256} within the shared-memory budget; msgm_attention_tc_supported says whether a shape is covered. */
int msgm_attention_tc_supported(int32_t C, int32_t T);
int msgm_attention_tc(msgm_ctx* ctx, const float* qkv, float* out, int32_t B, int32_t C, int32_t T, void* stream);
/* msgm_attention_tc with the AttentionBlock's output projection and residual in the same launch (model/unet.py:228-234:
 * `x + proj_out(attention(qkv(norm(x))))`): out (B,C,T) = W a + bias + res, a = the attention output, W the 1x1 conv's
 * weights packed by msgm_conv2d_tc_pack(W, C, C, 1, wimg); bias / res may be NULL; out must not alias res.
 * C in {32, 64, 128} within msgm_attention_tc's shapes (msgm_attention_proj_tc_supported). */
int msgm_attention_proj_tc_supported(int32_t C, int32_t T);
int msgm_attention_proj_tc(msgm_ctx* ctx, const float* qkv, const void* wimg, const float* bias, const float* res, float* out,
                           int32_t B, int32_t C, int32_t T, void* stream);
/* VorticityUNet wrapper: flat (B,H*W) -> image (B,1,H,W) / 5 [after x/(|x|+eps)*sqrt(d) when pre] and back (x5). */
int msgm_vort_pre(msgm_ctx* ctx, const float* x, float* img, float* lognorm, int32_t B, int32_t H, int32_t W, int32_t forder,
                  int32_t pre, void* stream);
int msgm_vort_post(msgm_ctx* ctx, const float* img, float* y, int32_t B, int32_t H, int32_t W, int32_t forder, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSGM_B200_H */
