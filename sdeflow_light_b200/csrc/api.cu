// C-ABI entry points of libmsgm_b200.so (see include/msgm_b200.h).  Validation + dispatch only.
#include <cstdio>
#include <cstring>

#include "msgm_common.cuh"

namespace msgm {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return MSGM_ERR_CUDA;
}

int sample_mlp_fp32(msgm_ctx*, const msgm_sde_desc*, const msgm_mlp_desc*, const msgm_sample_args*, float*, int64_t,
                    cudaStream_t, const float* t_noise = nullptr, const float* noise_single = nullptr);
int noise_forward(msgm_ctx*, const msgm_sde_desc*, const float*, float*, int, const float*, const float*, const float*,
                  uint64_t, uint64_t, int64_t, cudaStream_t);
int ssm_prepare(msgm_ctx*, const msgm_sde_desc*, const float*, float*, float*, float*, int, const float*, float, int, uint64_t,
                const uint64_t*, uint64_t, int64_t, cudaStream_t);
int sample_mlp_tc(msgm_ctx*, const msgm_sde_desc*, const msgm_mlp_desc*, const msgm_sample_args*, float*, int64_t,
                  cudaStream_t);
int mlp_forward_fp32(msgm_ctx*, const msgm_mlp_desc*, const float*, const float*, float*, int64_t, cudaStream_t);

size_t ssm_scratch_floats(long long B);
int ssm_forward(msgm_ctx*, const msgm_sde_desc*, const msgm_mlp_desc*, const float*, const float*, const float*, float*,
                float*, int64_t, cudaStream_t);
int ssm_backward(msgm_ctx*, const msgm_sde_desc*, const msgm_mlp_desc*, const float*, const float*, const float*,
                 const float*, float*, float*, int64_t, cudaStream_t);

int ssm_tc_grid(const msgm_ctx*, long long);
int ssm_fwd_bwd_tc(msgm_ctx*, const msgm_sde_desc*, const msgm_mlp_desc*, const float*, const float*, const float*,
                   const float*, float*, float*, float*, float, int64_t, cudaStream_t);

int stage_update(msgm_ctx*, const msgm_sde_desc*, int, int, float, int, int, float, float, const float*, const float*,
                 const float*, float*, float*, float*, int64_t, cudaStream_t, const msgm_step_clock*);
int row_norm(msgm_ctx*, const float*, float*, int, int64_t, cudaStream_t);
int philox_normal(msgm_ctx*, float*, int, int64_t, float, uint64_t, uint64_t, uint32_t, cudaStream_t, const msgm_step_clock*,
                  const float*, int);
int clock_advance(msgm_ctx*, int32_t*, cudaStream_t);

int latent_sample(msgm_ctx*, const float*, int, int, int, const float*, const float*, float*, int, int64_t, uint64_t,
                  uint64_t, cudaStream_t);
int mmd_sums(msgm_ctx*, const float*, int64_t, const float*, int64_t, int, double*, cudaStream_t);
int kde_logpdf(msgm_ctx*, const float*, int, float, const float*, float*, int, cudaStream_t);

int row_norm_stats(msgm_ctx*, const float*, const float*, float*, float*, int, int64_t, cudaStream_t);
int survival_counts(msgm_ctx*, const float*, int64_t, const double*, int, int64_t*, void*, cudaStream_t);
int moments(msgm_ctx*, const float*, int64_t, int, double*, double*, cudaStream_t);
int adam_step(msgm_ctx*, const void*, int, int64_t, const float*, float*, float*, const float*, int64_t*, float, float, float,
              float, cudaStream_t);

int conv1d(msgm_ctx*, const msgm_conv1d_desc*, cudaStream_t);
int emb_fold(msgm_ctx*, const float*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
int convt1d(msgm_ctx*, const float*, const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
int embed_mlp(msgm_ctx*, const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
              cudaStream_t);
int normalize_log_radius(msgm_ctx*, const float*, float*, float*, int, int, cudaStream_t);

int gn_stats(msgm_ctx*, const float*, int, const float*, int, int, int, int, float*, cudaStream_t);
int conv2d(msgm_ctx*, const msgm_conv2d_desc*, cudaStream_t);
int conv2d_tc(msgm_ctx*, const msgm_conv2d_tc_desc*, cudaStream_t);
size_t conv2d_tc_pack_bytes(int, int, int);
int conv2d_tc_pack(msgm_ctx*, const float*, int, int, int, int, void*, cudaStream_t, int dgrad = 0);
int conv1d_tc(msgm_ctx*, const msgm_conv1d_tc_desc*, cudaStream_t);
int conv1d_tcp(msgm_ctx*, const msgm_conv1d_tcp_desc*, cudaStream_t);
int emb_fold_multi(msgm_ctx*, const msgm_emb_fold_multi_desc*, cudaStream_t);
int embed_mlp2(msgm_ctx*, const float*, const float*, const float*, const float*, const float*, const float*, const float*,
               const float*, const float*, const float*, float*, int, int, cudaStream_t);
int64_t planes_bytes(int64_t, int, int);
int planes_pack(msgm_ctx*, const float*, void*, int, int, int, cudaStream_t);
int planes_unpack(msgm_ctx*, const void*, float*, int, int, int, cudaStream_t);
int conv1d_first_planes(msgm_ctx*, const float*, const float*, int, const float*, const float*, void*, int, int, int, int,
                        cudaStream_t);
int convt1d_tc_pack(msgm_ctx*, const float*, int, int, void*, cudaStream_t);
int convt1d_tc(msgm_ctx*, const float*, const void*, const float*, float*, int, int, int, int, int, int, cudaStream_t);
int gn_scale_shift(msgm_ctx*, const float*, int, const float*, int, int, int, int, const float*, const float*, float*, cudaStream_t);
int emb_proj(msgm_ctx*, const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
int sincos_embed_mlp(msgm_ctx*, const float*, const float*, const float*, const float*, const float*, float*, int, int, int,
                     int, cudaStream_t);
int attention(msgm_ctx*, const float*, float*, int, int, int, cudaStream_t);
bool attention_tc_ok(int, int);
int emb_proj_multi(msgm_ctx*, const float*, const float*, const float*, float*, int, int, int, int, const int*, cudaStream_t);
int attention_tc(msgm_ctx*, const float*, float*, int, int, int, cudaStream_t, const void* wimg = nullptr,
                 const float* pbias = nullptr, const float* res = nullptr);
bool attention_proj_tc_ok(int, int);
int vort_pre(msgm_ctx*, const float*, float*, float*, int, int, int, int, int, cudaStream_t);
int vort_post(msgm_ctx*, const float*, float*, int, int, int, int, cudaStream_t);

static int invalid(const char* msg) {
  set_error(msg);
  return MSGM_ERR_INVALID;
}

}  // namespace msgm

using namespace msgm;

extern "C" {

int msgm_abi_version(void) { return MSGM_ABI_VERSION; }
const char* msgm_last_error(void) { return g_err.c_str(); }

int msgm_create(msgm_ctx** out, int device) {
  if (!out) return invalid("msgm_create: out is NULL");
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
    cudaGetLastError();
    set_error("msgm_create: no CUDA device (this library has no CPU fallback)");
    return MSGM_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  MSGM_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof buf, "msgm_create: device %d is sm_%d%d; libmsgm_b200 is built for sm_100a only", device,
             prop.major, prop.minor);
    set_error(buf);
    return MSGM_ERR_NO_DEVICE;
  }
  msgm_ctx* c = new msgm_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  c->launches = 0;
  c->ws = nullptr;
  c->ws_bytes = 0;
  c->host_flag = nullptr;
  c->host_flag_dev = nullptr;
  c->launch_seq = 0;
  // fixed device workspace: [0,256) debug flags, [256,256K) packed fp16 weight image, [256K,512K) padded G
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaMalloc(&c->ws, 1 << 19);
  if (e == cudaSuccess) e = cudaMemset(c->ws, 0, 1 << 19);
  // error word of the tensor-core kernels: mapped pinned host memory, written by the device, polled by the host
  if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&c->host_flag), sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable);
  if (e == cudaSuccess) {
    *c->host_flag = 0;
    e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->host_flag_dev), c->host_flag, 0);
  }
  if (e != cudaSuccess) {
    if (c->ws) cudaFree(c->ws);
    if (c->host_flag) cudaFreeHost(c->host_flag);
    delete c;
    return cuda_fail(e, "msgm_create workspace");
  }
  c->ws_bytes = 1 << 19;
  *out = c;
  return MSGM_OK;
}

int msgm_destroy(msgm_ctx* ctx) {
  if (!ctx) return MSGM_OK;
  if (ctx->ws) cudaFree(ctx->ws);
  if (ctx->host_flag) cudaFreeHost(ctx->host_flag);
  delete ctx;
  return MSGM_OK;
}

int64_t msgm_launch_count(const msgm_ctx* ctx) { return ctx ? ctx->launches : 0; }

int msgm_debug_flags(msgm_ctx* ctx, int32_t* out_host) {
  if (!ctx || !out_host) return invalid("msgm_debug_flags: NULL argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  MSGM_CUDA_TRY(cudaDeviceSynchronize());
  return msgm_async_error(ctx, out_host);
}

int msgm_async_error(msgm_ctx* ctx, int32_t* out_host) {
  if (!ctx || !out_host) return invalid("msgm_async_error: NULL argument");
  volatile int* f = ctx->host_flag;
  *out_host = *f;
  if (*out_host != 0) *f = 0;  // read-and-clear
  return MSGM_OK;
}

static int check_mlp(const msgm_mlp_desc* m, int d) {
  if (!m) return invalid("mlp descriptor is NULL");
  if (m->input_dim != d) return invalid("mlp.input_dim != sde.dim");
  for (int l = 0; l < 4; ++l)
    if (!m->W[l] || !m->b[l]) return invalid("mlp weight pointer is NULL");
  return MSGM_OK;
}

int msgm_sample_mlp(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const msgm_sample_args* a,
                    float* x, int64_t B, void* stream) {
  if (!ctx || !sde || !a || !x) return invalid("msgm_sample_mlp: NULL argument");
  if (B < 0) return invalid("msgm_sample_mlp: B < 0");
  if (sde->dim < 1) return invalid("msgm_sample_mlp: dim < 1");
  if (sde->dim > MSGM_MAX_DIM_MLP) {
    set_error("msgm_sample_mlp: dim > 32 is not built for the MLP path (use the stage-update kernels)");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (sde->kind < MSGM_SDE_SGM || sde->kind > MSGM_SDE_MSGM_SPARSE) return invalid("unknown sde kind");
  if (sde->kind == MSGM_SDE_MSGM_DENSE && (!sde->G || !sde->L_G)) return invalid("dense MSGM needs G and L_G");
  if (a->scheme < MSGM_SCHEME_EM || a->scheme > MSGM_SCHEME_RK4) return invalid("unknown scheme");
  if (a->num_steps < 1) return invalid("num_steps < 1");
  if (a->lmbd < 0.0f || a->lmbd > 1.0f) return invalid("lmbd must be in [0,1]");
  if ((a->keep_step == nullptr) != (a->keep_out == nullptr)) return invalid("keep_step and keep_out go together");
  if (a->T_rows && !a->ts) return invalid("T_rows needs the unit time grid in ts");
  if (!a->forward_only) {
    int rc = check_mlp(mlp, sde->dim);
    if (rc) return rc;
  }
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  if (a->precision == MSGM_PREC_FP32) return sample_mlp_fp32(ctx, sde, mlp, a, x, B, (cudaStream_t)stream);
  if (a->precision == MSGM_PREC_F16TC) return sample_mlp_tc(ctx, sde, mlp, a, x, B, (cudaStream_t)stream);
  return invalid("msgm_sample_mlp: unknown precision");
}

int msgm_mlp_forward(msgm_ctx* ctx, const msgm_mlp_desc* mlp, const float* y, const float* s, float* out, int64_t B,
                     void* stream) {
  if (!ctx || !mlp || !y || !s || !out) return invalid("msgm_mlp_forward: NULL argument");
  if (mlp->input_dim < 1 || mlp->input_dim > MSGM_MAX_DIM_MLP) {
    set_error("msgm_mlp_forward: input_dim must be in [1,32]");
    return MSGM_ERR_UNSUPPORTED;
  }
  int rc = check_mlp(mlp, mlp->input_dim);
  if (rc) return rc;
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return mlp_forward_fp32(ctx, mlp, y, s, out, B, (cudaStream_t)stream);
}

static int check_ssm(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, int64_t B) {
  if (!ctx || !sde || !mlp) return invalid("msgm_ssm: NULL argument");
  if (B < 0) return invalid("msgm_ssm: B < 0");
  if (sde->dim < 1 || sde->dim > MSGM_MAX_DIM_MLP) {
    set_error("msgm_ssm: dim must be in [1,32] for the MLP path");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (sde->kind < MSGM_SDE_SGM || sde->kind > MSGM_SDE_MSGM_SPARSE) return invalid("unknown sde kind");
  if (sde->kind == MSGM_SDE_MSGM_DENSE && !sde->G) return invalid("dense MSGM needs G");
  return check_mlp(mlp, sde->dim);
}

uint64_t msgm_ssm_scratch_bytes(int64_t B) { return sizeof(float) * (uint64_t)ssm_scratch_floats(B < 1 ? 1 : B); }

int msgm_ssm_mlp_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                         const float* v, const float* t, float* loss_out, void* scratch, int64_t B, void* stream) {
  int rc = check_ssm(ctx, sde, mlp, B);
  if (rc) return rc;
  if (B == 0) return MSGM_OK;
  if (!y || !v || !t || !loss_out || !scratch) return invalid("msgm_ssm_mlp_forward: NULL buffer");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return ssm_forward(ctx, sde, mlp, y, v, t, loss_out, (float*)scratch, B, (cudaStream_t)stream);
}

int msgm_ssm_mlp_backward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                          const float* v, const float* t, const float* grad_out, void* scratch, float* grad_flat,
                          int64_t B, void* stream) {
  int rc = check_ssm(ctx, sde, mlp, B);
  if (rc) return rc;
  if (!grad_flat) return invalid("msgm_ssm_mlp_backward: NULL buffer");
  if (B == 0) return MSGM_OK;
  if (!y || !v || !t || !grad_out || !scratch) return invalid("msgm_ssm_mlp_backward: NULL buffer");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return ssm_backward(ctx, sde, mlp, y, v, t, grad_out, (float*)scratch, grad_flat, B, (cudaStream_t)stream);
}

uint64_t msgm_ssm_tc_scratch_bytes(const msgm_ctx* ctx, int32_t d, int32_t premodule, int64_t B) {
  if (!ctx || d < 1) return 0;
  const uint64_t nparam = 128ull * (d + 1 + (premodule ? 1 : 0)) + 128 + 2 * (128 * 128 + 128) + 128ull * d + d;
  return sizeof(float) * ((nparam + 3) & ~3ull) * (uint64_t)ssm_tc_grid(ctx, B < 1 ? 1 : B);
}

int msgm_ssm_mlp_fwd_bwd_tc(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y,
                            const float* v, const float* t, const float* grad_out, float* loss_out, float* grad_flat,
                            void* scratch, float cot_scale, int64_t B, void* stream) {
  int rc = check_ssm(ctx, sde, mlp, B);
  if (rc) return rc;
  if (!loss_out || !grad_flat) return invalid("msgm_ssm_mlp_fwd_bwd_tc: NULL buffer");
  if (B == 0) return MSGM_OK;
  if (!y || !v || !t || !grad_out || !scratch) return invalid("msgm_ssm_mlp_fwd_bwd_tc: NULL buffer");
  if (!(cot_scale > 0.0f)) return invalid("msgm_ssm_mlp_fwd_bwd_tc: cot_scale must be positive");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return ssm_fwd_bwd_tc(ctx, sde, mlp, y, v, t, grad_out, loss_out, grad_flat, (float*)scratch, cot_scale, B,
                        (cudaStream_t)stream);
}

static int stage_update_checked(const char* who, msgm_ctx* ctx, const msgm_sde_desc* sde, int32_t scheme, int32_t stage, float lmbd,
                                int32_t norm_correction, int32_t forward_only, float s, float delta, const float* a,
                                const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B, void* stream,
                                const msgm_step_clock* clk) {
  (void)who;
  if (!ctx || !sde || !dW || !x || !y || !ks) return invalid("msgm_stage_update: NULL argument");
  if (!forward_only && !a) return invalid("msgm_stage_update: score-net output missing");
  if (norm_correction && !r0) return invalid("msgm_stage_update: r0 missing");
  if (sde->kind != MSGM_SDE_SGM && sde->kind != MSGM_SDE_MSGM_SPARSE) {
    set_error("msgm_stage_update: built for SGM and sparse MSGM (dense G with d > 32 is O(d^3) per particle)");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (sde->dim < 1 || sde->dim > 4096) {
    set_error("msgm_stage_update: d must be in [1,4096]");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (scheme < MSGM_SCHEME_EM || scheme > MSGM_SCHEME_RK4) return invalid("unknown scheme");
  const int nstage = scheme == MSGM_SCHEME_RK4 ? 4 : (scheme == MSGM_SCHEME_HEUN ? 2 : 1);
  if (stage < 0 || stage >= nstage) return invalid("stage out of range");
  if (clk && (!clk->clock || !clk->s_table)) return invalid("msgm_stage_update_clocked: clock / s_table missing");
  if (clk && clk->keep_out && !clk->keep_step) return invalid("msgm_stage_update_clocked: keep_step missing");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return stage_update(ctx, sde, scheme, stage, lmbd, norm_correction, forward_only, s, delta, a, dW, r0, x, y, ks, B,
                      (cudaStream_t)stream, clk);
}

int msgm_stage_update(msgm_ctx* ctx, const msgm_sde_desc* sde, int32_t scheme, int32_t stage, float lmbd,
                      int32_t norm_correction, int32_t forward_only, float s, float delta, const float* a,
                      const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B, void* stream) {
  return stage_update_checked("msgm_stage_update", ctx, sde, scheme, stage, lmbd, norm_correction, forward_only, s, delta, a, dW, r0,
                              x, y, ks, B, stream, nullptr);
}

int msgm_stage_update_clocked(msgm_ctx* ctx, const msgm_sde_desc* sde, int32_t scheme, int32_t stage, float lmbd,
                              int32_t norm_correction, int32_t forward_only, const msgm_step_clock* clk, float delta,
                              const float* a, const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B,
                              void* stream) {
  if (!clk) return invalid("msgm_stage_update_clocked: clock missing");
  return stage_update_checked("msgm_stage_update_clocked", ctx, sde, scheme, stage, lmbd, norm_correction, forward_only, 0.0f, delta,
                              a, dW, r0, x, y, ks, B, stream, clk);
}

int msgm_tc_range_scale(msgm_ctx* ctx, const float* amax_or_null) {
  if (!ctx) return invalid("msgm_tc_range_scale: NULL context");
  ctx->tc_in_amax = reinterpret_cast<const unsigned int*>(amax_or_null);
  return MSGM_OK;
}

int msgm_row_norm(msgm_ctx* ctx, const float* x, float* r, int32_t d, int64_t B, void* stream) {
  if (!ctx || !x || !r || d < 1) return invalid("msgm_row_norm: bad argument");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return row_norm(ctx, x, r, d, B, (cudaStream_t)stream);
}

int msgm_philox_normal(msgm_ctx* ctx, float* out, int32_t d, int64_t B, float scale, uint64_t seed,
                       uint64_t particle_offset, uint32_t step, void* stream) {
  if (!ctx || !out || d < 1) return invalid("msgm_philox_normal: bad argument");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return philox_normal(ctx, out, d, B, scale, seed, particle_offset, step, (cudaStream_t)stream, nullptr, nullptr, 1);
}

int msgm_philox_normal_clocked(msgm_ctx* ctx, float* out, int32_t d, int64_t B, float scale, uint64_t seed,
                               uint64_t particle_offset, const msgm_step_clock* clk, int32_t nstage,
                               const float* noise_or_null, void* stream) {
  if (!ctx || !out || d < 1 || !clk || !clk->clock || !clk->s_table || nstage < 1)
    return invalid("msgm_philox_normal_clocked: bad argument");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return philox_normal(ctx, out, d, B, scale, seed, particle_offset, 0u, (cudaStream_t)stream, clk, noise_or_null, nstage);
}

int msgm_clock_advance(msgm_ctx* ctx, int32_t* clock, void* stream) {
  if (!ctx || !clock) return invalid("msgm_clock_advance: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return clock_advance(ctx, clock, (cudaStream_t)stream);
}

int msgm_latent_sample(msgm_ctx* ctx, const float* rT_sorted, int32_t n_r, int32_t log_map, int32_t msgm, const float* U,
                       const float* Z, float* out, int32_t d, int64_t B, uint64_t seed, uint64_t particle_offset,
                       void* stream) {
  if (!ctx || !out || d < 1) return invalid("msgm_latent_sample: bad argument");
  if (msgm && (!rT_sorted || n_r < 1)) return invalid("msgm_latent_sample: radius table missing");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return latent_sample(ctx, rT_sorted, n_r, log_map, msgm, U, Z, out, d, B, seed, particle_offset, (cudaStream_t)stream);
}

int msgm_mmd_sums(msgm_ctx* ctx, const float* x, int64_t N, const float* y, int64_t M, int32_t d, double* sums_out,
                  void* stream) {
  if (!ctx || !x || !y || !sums_out || d < 1 || N < 1 || M < 1) return invalid("msgm_mmd_sums: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return mmd_sums(ctx, x, N, y, M, d, sums_out, (cudaStream_t)stream);
}

int msgm_row_norm_stats(msgm_ctx* ctx, const float* x, const float* scale_opt, float* norms_out, float* minpos_max_out,
                        int32_t d, int64_t n, void* stream) {
  if (!ctx || !x || !norms_out || !minpos_max_out || d < 1 || n < 1) return invalid("msgm_row_norm_stats: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return row_norm_stats(ctx, x, scale_opt, norms_out, minpos_max_out, d, n, (cudaStream_t)stream);
}

int msgm_survival_counts(msgm_ctx* ctx, const float* norms, int64_t n, const double* R_grid, int32_t n_grid,
                         int64_t* counts_out, void* scratch, void* stream) {
  if (!ctx || !norms || !R_grid || !counts_out || !scratch || n < 1 || n_grid < 1)
    return invalid("msgm_survival_counts: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return survival_counts(ctx, norms, n, R_grid, n_grid, counts_out, scratch, (cudaStream_t)stream);
}

int msgm_moments(msgm_ctx* ctx, const float* x, int64_t n, int32_t d, double* colsum_out, double* gram_out, void* stream) {
  if (!ctx || !x || !colsum_out || !gram_out || n < 1 || d < 1) return invalid("msgm_moments: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return moments(ctx, x, n, d, colsum_out, gram_out, (cudaStream_t)stream);
}

int msgm_adam_step(msgm_ctx* ctx, const void* seg_table, int32_t n_tensors, int64_t total, const float* grad_flat,
                   float* exp_avg, float* exp_avg_sq, const float* lr_dev, int64_t* step_dev, float beta1, float beta2,
                   float eps, float grad_scale, void* stream) {
  if (!ctx || !seg_table || !grad_flat || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || n_tensors < 1 || total < 1)
    return invalid("msgm_adam_step: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return adam_step(ctx, seg_table, n_tensors, total, grad_flat, exp_avg, exp_avg_sq, lr_dev, step_dev, beta1, beta2, eps,
                   grad_scale, (cudaStream_t)stream);
}

int msgm_kde_logpdf(msgm_ctx* ctx, const float* samples, int32_t n, float bandwidth, const float* queries, float* out,
                    int32_t m, void* stream) {
  if (!ctx || !samples || !queries || !out || n < 1 || m < 0 || !(bandwidth > 0.0f))
    return invalid("msgm_kde_logpdf: bad argument");
  if (m == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return kde_logpdf(ctx, samples, n, bandwidth, queries, out, m, (cudaStream_t)stream);
}

int msgm_conv1d(msgm_ctx* ctx, const msgm_conv1d_desc* D, void* stream) {
  if (!ctx || !D || !D->x1 || !D->W || !D->out) return invalid("msgm_conv1d: NULL argument");
  if (D->K < 1 || D->K > 4 || D->stride < 1 || D->stride > 2 || D->B < 0 || D->Cout < 1 || D->C1 < 1 || D->Lin < 1 ||
      D->Lout < 1)
    return invalid("msgm_conv1d: unsupported shape (k <= 4, stride <= 2)");
  if ((D->Cemb > 0) != (D->E != nullptr)) return invalid("msgm_conv1d: Cemb and E go together");
  if (D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv1d(ctx, D, (cudaStream_t)stream);
}

int msgm_emb_fold(msgm_ctx* ctx, const float* W, const float* emb, float* E, int32_t Cw, int32_t Coff, int32_t Cemb,
                  int32_t Cout, int32_t K, int32_t B, void* stream) {
  if (!ctx || !W || !emb || !E || Cemb < 1 || Cemb > 1024 || Cout < 1 || K < 1 || B < 0) return invalid("msgm_emb_fold: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return emb_fold(ctx, W, emb, E, Cw, Coff, Cemb, Cout, K, B, (cudaStream_t)stream);
}

int msgm_convt1d_k4s2(msgm_ctx* ctx, const float* x, const float* W, const float* bias, float* out, int32_t B, int32_t Cin,
                      int32_t Cout, int32_t Lin, int32_t Lout, void* stream) {
  if (!ctx || !x || !W || !bias || !out || Cin < 1 || Cout < 1 || Lin < 1 || Lout < 2 * Lin || B < 0)
    return invalid("msgm_convt1d_k4s2: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return convt1d(ctx, x, W, bias, out, B, Cin, Cout, Lin, Lout, (cudaStream_t)stream);
}

int msgm_embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2,
                   float* out, int32_t B, int32_t E, int32_t accumulate, void* stream) {
  if (!ctx || !t || !W1 || !b1 || !W2 || !b2 || !out || E < 1 || E > 256 || B < 0) return invalid("msgm_embed_mlp: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return embed_mlp(ctx, t, W1, b1, W2, b2, out, B, E, accumulate, (cudaStream_t)stream);
}

int msgm_normalize_log_radius(msgm_ctx* ctx, const float* x, float* xn, float* lognorm, int32_t B, int32_t L, void* stream) {
  if (!ctx || !x || !xn || !lognorm || L < 1 || B < 0) return invalid("msgm_normalize_log_radius: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return normalize_log_radius(ctx, x, xn, lognorm, B, L, (cudaStream_t)stream);
}

int msgm_conv2d(msgm_ctx* ctx, const msgm_conv2d_desc* D, void* stream) {
  if (!ctx || !D || !D->x1 || !D->W || !D->out) return invalid("msgm_conv2d: NULL argument");
  if ((D->K != 1 && D->K != 3) || D->stride < 1 || D->stride > 2 || (D->up != 1 && D->up != 2) || D->Cout < 1 || D->C1 < 1 ||
      D->Hs < 1 || D->Ws < 1 || D->B < 0)
    return invalid("msgm_conv2d: unsupported shape (k in {1,3}, stride <= 2, up in {1,2})");
  {  // the staged input tile (rows touched by 128 consecutive output positions x padded width) must fit the kernel's buffer
    const int pad = D->K == 3 ? 1 : 0, wcols = D->Ws * D->up + 2 * pad;
    const int Ho = (D->Hs * D->up + 2 * pad - D->K) / D->stride + 1, Wo = (D->Ws * D->up + 2 * pad - D->K) / D->stride + 1;
    int span = (128 % Wo == 0) ? 128 / Wo - 1 : 127 / Wo + 1;
    span = span < Ho - 1 ? span : Ho - 1;
    if ((long long)(span * D->stride + D->K) * wcols > (D->K == 3 ? 17 * 34 : 160)) {
      set_error("msgm_conv2d: image too wide for the staged tile (built for the reference's 32x32 / 16x16 / 8x8 levels)");
      return MSGM_ERR_UNSUPPORTED;
    }
  }
  if (D->prologue && (!D->stats || !D->gamma || !D->beta || D->G < 1)) return invalid("msgm_conv2d: GroupNorm inputs missing");
  if (D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv2d(ctx, D, (cudaStream_t)stream);
}

static bool conv2d_tc_shape_ok(int Cout, int Cin, int C1, int K) {
  return (K == 1 || K == 3) && Cout >= 32 && Cout % 32 == 0 && Cin >= 16 && Cin % 16 == 0 && C1 % 16 == 0;
}

int msgm_conv2d_tc(msgm_ctx* ctx, const msgm_conv2d_tc_desc* D, void* stream) {
  if (!ctx || !D || !D->x1 || !D->wimg || !D->out) return invalid("msgm_conv2d_tc: NULL argument");
  const int Cin = D->C1 + (D->x2 ? D->C2 : 0);
  if (!conv2d_tc_shape_ok(D->Cout, Cin, D->C1, D->K) || D->stride < 1 || D->stride > 2 || (D->up != 1 && D->up != 2) ||
      D->Hs < 1 || D->Ws < 1 || D->B < 0 || (D->stride == 2 && ((D->Hs * D->up) % 2 || (D->Ws * D->up) % 2 || D->K != 3)))
    return invalid("msgm_conv2d_tc: unsupported shape (k in {1,3}, Cin % 16 == 0, C1 % 16 == 0, Cout % 32 == 0, stride <= 2)");
  if (D->prologue && !D->ss) return invalid("msgm_conv2d_tc: prologue without a scale/shift table");
  if (D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv2d_tc(ctx, D, (cudaStream_t)stream);
}

int64_t msgm_conv2d_tc_pack_bytes(int32_t Cout, int32_t Cin, int32_t K) {
  if (!conv2d_tc_shape_ok(Cout, Cin, 0, K)) return -1;
  return (int64_t)conv2d_tc_pack_bytes(Cout, Cin, K * K);
}

int msgm_conv2d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cin, int32_t K, void* wimg, void* stream) {
  if (!ctx || !W || !wimg) return invalid("msgm_conv2d_tc_pack: NULL argument");
  if (!conv2d_tc_shape_ok(Cout, Cin, 0, K)) return invalid("msgm_conv2d_tc_pack: unsupported shape");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv2d_tc_pack(ctx, W, Cout, Cin, Cin, K * K, wimg, (cudaStream_t)stream);
}

static bool conv1d_tc_shape_ok(int Cout, int Cin, int C1, int K) {
  return (K == 1 || K == 3 || K == 4) && Cout >= 32 && Cout % 32 == 0 && Cin >= 16 && Cin % 16 == 0 && C1 % 16 == 0;
}

int msgm_conv1d_tc(msgm_ctx* ctx, const msgm_conv1d_tc_desc* D, void* stream) {
  if (!ctx || !D || !D->x1 || !D->wimg || !D->out) return invalid("msgm_conv1d_tc: NULL argument");
  const int Cin = D->C1 + (D->x2 ? D->C2 : 0);
  if (!conv1d_tc_shape_ok(D->Cout, Cin, D->C1, D->K) || D->Lin < 1 || D->B < 0 ||
      !((D->K == 4 && D->stride == 2 && D->Lin >= 2) || (D->K != 4 && D->stride == 1)))
    return invalid("msgm_conv1d_tc: unsupported shape (k3/k1 stride 1 or k4 stride 2, Cin % 16 == 0, C1 % 16 == 0, Cout % 32 == 0)");
  if (D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv1d_tc(ctx, D, (cudaStream_t)stream);
}

int msgm_emb_fold_multi(msgm_ctx* ctx, const msgm_emb_fold_multi_desc* D, void* stream) {
  if (!ctx || !D || !D->emb) return invalid("msgm_emb_fold_multi: NULL argument");
  if (D->n < 0 || D->n > 16 || D->B < 0 || D->Cemb < 1 || D->Cemb > 1024) return invalid("msgm_emb_fold_multi: n <= 16, Cemb in 1..1024");
  for (int i = 0; i < D->n; ++i)
    if (!D->W[i] || !D->E[i] || D->Cout[i] < 1 || D->K[i] < 1 || D->Coff[i] < 0 || D->Cw[i] < D->Coff[i] + D->Cemb)
      return invalid("msgm_emb_fold_multi: bad layer entry");
  if (D->n == 0 || D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return emb_fold_multi(ctx, D, (cudaStream_t)stream);
}

int msgm_embed_mlp2(msgm_ctx* ctx, const float* t, const float* W1a, const float* b1a, const float* W2a, const float* b2a,
                    const float* u, const float* W1b, const float* b1b, const float* W2b, const float* b2b, float* out,
                    int32_t B, int32_t E, void* stream) {
  if (!ctx || !t || !W1a || !b1a || !W2a || !b2a || !out || (u && (!W1b || !b1b || !W2b || !b2b)))
    return invalid("msgm_embed_mlp2: NULL argument");
  if (B < 0 || E < 1 || E > 224) return invalid("msgm_embed_mlp2: E in 1..224");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return embed_mlp2(ctx, t, W1a, b1a, W2a, b2a, u, W1b, b1b, W2b, b2b, out, B, E, (cudaStream_t)stream);
}

int64_t msgm_planes_bytes(int64_t B, int32_t C, int32_t L) {
  if (B < 0 || C < 8 || C % 8 != 0 || L < 1) return -1;
  return planes_bytes(B, C, L);
}

int msgm_conv1d_tcp(msgm_ctx* ctx, const msgm_conv1d_tcp_desc* D, void* stream) {
  if (!ctx || !D || !D->x1 || !D->wimg || (!D->out_planes && !D->out_f32)) return invalid("msgm_conv1d_tcp: NULL argument");
  const int Cin = D->C1 + (D->x2 ? D->C2 : 0);
  const bool shape_ok = D->transposed
                            ? (D->K == 3 && Cin % 16 == 0 && D->C1 % 16 == 0 && D->Cout % 16 == 0 && D->Cout > 0 && D->Lout >= 2 * D->Lin)
                            : (conv1d_tc_shape_ok(D->Cout, Cin, D->C1, D->K) && (D->K == 3 || (D->K == 4 && D->Lin >= 2)));
  if (!shape_ok || D->Lin < 1 || D->B < 0 || Cin <= 0)
    return invalid("msgm_conv1d_tcp: unsupported shape (k3 stride 1, k4 stride 2 or transposed k4 s2; Cin % 16 == 0, C1 % 16 == 0, Cout % 32 == 0)");
  if (D->out_planes && (D->out_planes == D->x1 || (D->x2 && D->out_planes == D->x2)))
    return invalid("msgm_conv1d_tcp: the output planes alias an input");
  if (D->B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv1d_tcp(ctx, D, (cudaStream_t)stream);
}

int msgm_planes_pack(msgm_ctx* ctx, const float* x, void* planes, int32_t B, int32_t C, int32_t L, void* stream) {
  if (!ctx || !x || !planes) return invalid("msgm_planes_pack: NULL argument");
  if (B < 0 || C < 8 || C % 8 != 0 || L < 1) return invalid("msgm_planes_pack: C % 8 == 0, L >= 1");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return planes_pack(ctx, x, planes, B, C, L, (cudaStream_t)stream);
}

int msgm_planes_unpack(msgm_ctx* ctx, const void* planes, float* x, int32_t B, int32_t C, int32_t L, void* stream) {
  if (!ctx || !x || !planes) return invalid("msgm_planes_unpack: NULL argument");
  if (B < 0 || C < 8 || C % 8 != 0 || L < 1) return invalid("msgm_planes_unpack: C % 8 == 0, L >= 1");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return planes_unpack(ctx, planes, x, B, C, L, (cudaStream_t)stream);
}

int msgm_conv1d_first_planes(msgm_ctx* ctx, const float* x, const float* W, int32_t Cw, const float* bias, const float* E,
                             void* planes, int32_t B, int32_t Cout, int32_t L, int32_t gelu, void* stream) {
  if (!ctx || !x || !W || !planes) return invalid("msgm_conv1d_first_planes: NULL argument");
  if (B < 0 || Cout < 8 || Cout % 8 != 0 || Cout > 128 || L < 1 || Cw < 1)
    return invalid("msgm_conv1d_first_planes: Cout % 8 == 0, Cout <= 128");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv1d_first_planes(ctx, x, W, Cw, bias, E, planes, B, Cout, L, gelu, (cudaStream_t)stream);
}

int64_t msgm_conv1d_tc_pack_bytes(int32_t Cout, int32_t Cin, int32_t K) {
  if (!conv1d_tc_shape_ok(Cout, Cin, 0, K)) return -1;
  return (int64_t)conv2d_tc_pack_bytes(Cout, Cin, K);
}

int msgm_conv1d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cw, int32_t Cin, int32_t K, void* wimg,
                        void* stream) {
  if (!ctx || !W || !wimg) return invalid("msgm_conv1d_tc_pack: NULL argument");
  if (!conv1d_tc_shape_ok(Cout, Cin, 0, K) || Cw < Cin) return invalid("msgm_conv1d_tc_pack: unsupported shape");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv2d_tc_pack(ctx, W, Cout, Cw, Cin, K, wimg, (cudaStream_t)stream);
}

int msgm_conv_tc_pack_dgrad(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cw, int32_t Cin, int32_t taps, void* wimg,
                            void* stream) {
  if (!ctx || !W || !wimg) return invalid("msgm_conv_tc_pack_dgrad: NULL argument");
  // the data-gradient conv has Cin output and Cout input channels
  if (!(taps == 1 || taps == 3 || taps == 9) || Cin < 32 || Cin % 32 || Cout < 16 || Cout % 16 || Cw < Cin)
    return invalid("msgm_conv_tc_pack_dgrad: unsupported shape (taps 1 / 3 / 9, Cin % 32 == 0, Cout % 16 == 0)");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv2d_tc_pack(ctx, W, Cin, Cw, Cout, taps, wimg, (cudaStream_t)stream, 1);
}

int msgm_convt1d_tc_pack(msgm_ctx* ctx, const float* W, int32_t Cout, int32_t Cin, void* wimg, void* stream) {
  if (!ctx || !W || !wimg) return invalid("msgm_convt1d_tc_pack: NULL argument");
  if (Cout < 16 || Cout % 16 || Cin < 16 || Cin % 16) return invalid("msgm_convt1d_tc_pack: Cin % 16 == 0 and Cout % 16 == 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return convt1d_tc_pack(ctx, W, Cout, Cin, wimg, (cudaStream_t)stream);
}

int msgm_convt1d_tc(msgm_ctx* ctx, const float* x, const void* wimg, const float* bias, float* out, int32_t B, int32_t Cin,
                    int32_t Cout, int32_t Lin, int32_t Lout, int32_t fast, void* stream) {
  if (!ctx || !x || !wimg || !out || B < 0 || Lin < 1 || Lout < 2 * Lin) return invalid("msgm_convt1d_tc: bad argument");
  if (Cout < 16 || Cout % 16 || Cin < 16 || Cin % 16) return invalid("msgm_convt1d_tc: Cin % 16 == 0 and Cout % 16 == 0");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return convt1d_tc(ctx, x, wimg, bias, out, B, Cin, Cout, Lin, Lout, fast, (cudaStream_t)stream);
}

int msgm_gn_scale_shift(msgm_ctx* ctx, const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t HW, int32_t G,
                        int32_t B, const float* gamma, const float* beta, float* ss, void* stream) {
  if (!ctx || !x1 || !gamma || !beta || !ss || C1 < 1 || G < 1 || HW < 1 || B < 0) return invalid("msgm_gn_scale_shift: bad argument");
  const int C = C1 + (x2 ? C2 : 0);
  if (C % G || C / G > 256) return invalid("msgm_gn_scale_shift: channels not divisible by groups");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return gn_scale_shift(ctx, x1, C1, x2, C2, HW, G, B, gamma, beta, ss, (cudaStream_t)stream);
}

int msgm_gn_stats(msgm_ctx* ctx, const float* x1, int32_t C1, const float* x2, int32_t C2, int32_t HW, int32_t G, int32_t B,
                  float* stats, void* stream) {
  if (!ctx || !x1 || !stats || C1 < 1 || G < 1 || HW < 1 || B < 0) return invalid("msgm_gn_stats: bad argument");
  if ((C1 + (x2 ? C2 : 0)) % G) return invalid("msgm_gn_stats: channels not divisible by groups");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return gn_stats(ctx, x1, C1, x2, C2, HW, G, B, stats, (cudaStream_t)stream);
}

int msgm_emb_proj(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int32_t E, int32_t Cout,
                  int32_t B, void* stream) {
  if (!ctx || !emb || !W || !bias || !out || E < 1 || E > 8192 || Cout < 1 || B < 0) return invalid("msgm_emb_proj: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return emb_proj(ctx, emb, W, bias, out, E, Cout, B, (cudaStream_t)stream);
}

int msgm_emb_proj_multi(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int32_t E,
                        int32_t Ctot, int32_t B, int32_t nseg, const int32_t* seg_start, void* stream) {
  if (!ctx || !emb || !W || !bias || !out || !seg_start || E < 1 || E > 8192 || Ctot < 1 || B < 0 || nseg < 1 || nseg > 64)
    return invalid("msgm_emb_proj_multi: bad argument");
  if (seg_start[0] != 0 || seg_start[nseg] != Ctot) return invalid("msgm_emb_proj_multi: segments must cover [0, Ctot)");
  for (int i = 0; i < nseg; ++i)
    if (seg_start[i + 1] <= seg_start[i]) return invalid("msgm_emb_proj_multi: segment offsets must increase");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return emb_proj_multi(ctx, emb, W, bias, out, E, Ctot, B, nseg, seg_start, (cudaStream_t)stream);
}

int msgm_sincos_embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2,
                          float* out, int32_t B, int32_t dim, int32_t E, int32_t accumulate, void* stream) {
  if (!ctx || !t || !W1 || !b1 || !W2 || !b2 || !out || dim < 2 || dim > 64 || (dim & 1) || E < 1 || E > 256 || B < 0)
    return invalid("msgm_sincos_embed_mlp: bad argument (even dim <= 64, E <= 256)");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return sincos_embed_mlp(ctx, t, W1, b1, W2, b2, out, B, dim, E, accumulate, (cudaStream_t)stream);
}

int msgm_attention_tc_supported(int32_t C, int32_t T) { return attention_tc_ok(C, T) ? 1 : 0; }

int msgm_attention_tc(msgm_ctx* ctx, const float* qkv, float* out, int32_t B, int32_t C, int32_t T, void* stream) {
  if (!ctx || !qkv || !out || B < 0) return invalid("msgm_attention_tc: bad argument");
  if (!attention_tc_ok(C, T)) {
    set_error("msgm_attention_tc: shape not covered (C % 32 == 0, C <= 128, T % 64 == 0, T <= 256); use msgm_attention");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return attention_tc(ctx, qkv, out, B, C, T, (cudaStream_t)stream);
}

int msgm_attention_proj_tc_supported(int32_t C, int32_t T) { return attention_proj_tc_ok(C, T) ? 1 : 0; }

int msgm_attention_proj_tc(msgm_ctx* ctx, const float* qkv, const void* wimg, const float* bias, const float* res, float* out,
                           int32_t B, int32_t C, int32_t T, void* stream) {
  if (!ctx || !qkv || !wimg || !out || B < 0) return invalid("msgm_attention_proj_tc: bad argument");
  if (out == res) return invalid("msgm_attention_proj_tc: out aliases res");
  if (!attention_proj_tc_ok(C, T)) {
    set_error("msgm_attention_proj_tc: shape not covered (C in {32, 64, 128}, T % 64 == 0, T <= 256, C <= 2 T)");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return attention_tc(ctx, qkv, out, B, C, T, (cudaStream_t)stream, wimg, bias, res);
}

int msgm_attention(msgm_ctx* ctx, const float* qkv, float* out, int32_t B, int32_t C, int32_t T, void* stream) {
  if (!ctx || !qkv || !out || C < 1 || T < 1 || T > 4096 || B < 0) return invalid("msgm_attention: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return attention(ctx, qkv, out, B, C, T, (cudaStream_t)stream);
}

int msgm_vort_pre(msgm_ctx* ctx, const float* x, float* img, float* lognorm, int32_t B, int32_t H, int32_t W, int32_t forder,
                  int32_t pre, void* stream) {
  if (!ctx || !x || !img || (pre && !lognorm) || H < 1 || W < 1 || B < 0) return invalid("msgm_vort_pre: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return vort_pre(ctx, x, img, lognorm, B, H, W, forder, pre, (cudaStream_t)stream);
}

int msgm_vort_post(msgm_ctx* ctx, const float* img, float* y, int32_t B, int32_t H, int32_t W, int32_t forder, void* stream) {
  if (!ctx || !img || !y || H < 1 || W < 1 || B < 0) return invalid("msgm_vort_post: bad argument");
  if (B == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return vort_post(ctx, img, y, B, H, W, forder, (cudaStream_t)stream);
}

int msgm_noise_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* t, float* y_inout, int32_t num_steps_forward,
                       const float* ts, const float* noise, const float* noise_single, uint64_t seed,
                       uint64_t particle_offset, int64_t B, void* stream) {
  if (!ctx || !sde || !t || !y_inout) return invalid("msgm_noise_forward: NULL argument");
  if (sde->kind != MSGM_SDE_MSGM_DENSE && sde->kind != MSGM_SDE_MSGM_SPARSE)
    return invalid("msgm_noise_forward: the additive SDE has a closed-form marginal (SDEs.py:134-146)");
  if (sde->dim < 1 || sde->dim > 4096 || (sde->dim > MSGM_MAX_DIM_MLP && sde->kind == MSGM_SDE_MSGM_DENSE)) {
    set_error("msgm_noise_forward: dim must be in [1,32] for the dense tensor, [1,4096] for the sparse tensor / SGM");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (sde->kind == MSGM_SDE_MSGM_DENSE && (!sde->G || !sde->L_G)) return invalid("dense MSGM needs G and L_G");
  if (num_steps_forward < 1) return invalid("num_steps_forward < 1");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return noise_forward(ctx, sde, t, y_inout, num_steps_forward, ts, noise, noise_single, seed, particle_offset, B,
                       (cudaStream_t)stream);
}

int msgm_ssm_prepare(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* x, float* t_out, float* v_out, float* y_out,
                     int32_t num_steps_forward, const float* ts, float t_epsilon, int32_t vtype, uint64_t seed,
                     const uint64_t* seed_offset_dev, uint64_t sample_offset, int64_t B, void* stream) {
  if (!ctx || !sde || !x || !t_out || !v_out || !y_out) return invalid("msgm_ssm_prepare: NULL argument");
  if (sde->dim < 1 || sde->dim > 4096 || (sde->dim > MSGM_MAX_DIM_MLP && sde->kind == MSGM_SDE_MSGM_DENSE)) {
    set_error("msgm_ssm_prepare: dim must be in [1,32] for the dense tensor, [1,4096] for the sparse tensor / SGM");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (sde->kind != MSGM_SDE_SGM && sde->kind != MSGM_SDE_MSGM_DENSE && sde->kind != MSGM_SDE_MSGM_SPARSE)
    return invalid("msgm_ssm_prepare: unknown sde kind");
  if (sde->kind == MSGM_SDE_MSGM_DENSE && !sde->G) return invalid("dense MSGM needs G");
  if (vtype < MSGM_V_RADEMACHER || vtype > MSGM_V_SPHERE) return invalid("msgm_ssm_prepare: unknown vtype");
  if (num_steps_forward < 1) return invalid("num_steps_forward < 1");
  if (B <= 0) return B == 0 ? MSGM_OK : invalid("B < 0");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return ssm_prepare(ctx, sde, x, t_out, v_out, y_out, num_steps_forward, ts, t_epsilon, vtype, seed, seed_offset_dev,
                     sample_offset, B, (cudaStream_t)stream);
}

int msgm_debug_counters(msgm_ctx* ctx, int64_t* out_host, int n) {
  if (!ctx || !out_host || n < 0 || n > 24) return invalid("msgm_debug_counters: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  MSGM_CUDA_TRY(cudaDeviceSynchronize());
  MSGM_CUDA_TRY(cudaMemcpy(out_host, reinterpret_cast<unsigned char*>(ctx->ws) + 64, sizeof(int64_t) * n,
                           cudaMemcpyDeviceToHost));
  return MSGM_OK;
}

}  // extern "C"
