// Per-stage SDE update kernels for score nets that are NOT fused into the sampler (U-Nets: d = 1000 / 1024).
//
// The reference evaluates, per Runge-Kutta stage, g(s,y) three times, f twice, a gather/scatter_add with atomics and
// ~15 elementwise ATen kernels (sde_scheme.py:18-40,223-255; SDEs.py:556-588).  Here one kernel per stage does
//   w = delta (1 - lambda/2) a + sqrt(1-lambda) dW ;  K = g(s,y) . w + delta c_f f(s,y)      (3-point cyclic stencil)
//   Runge-Kutta bookkeeping (running sum, next stage input or new state) ;  radius re-pin on the last stage
// with one read of (x, y, a, dW, ksum) and one write of (ksum, y_next | x) per element: 20-28 B per element-stage,
// coalesced, one CTA per particle row so that the stencil neighbours come from L1 and the row norm is a block reduce.
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

struct StageParams {
  int kind;        // SGM or MSGM_SPARSE
  int d;
  int scheme;      // EM / HEUN / RK4
  int stage;       // 0 .. nstage-1
  int nc;          // re-pin the radius on the last stage
  int fwd;         // forward adapter: no net, time runs forward
  float s;         // noise time of this stage (T - t for the reverse SDE)
  float bmin, bdel, delta, lmbd;
  const float* a;  // (B,d) score-net output at (y, s); NULL when fwd
  const float* dW; // (B,d) Wiener increment of the step
  const float* r0; // (B,) initial radii (nc)
  float* x;        // (B,d) state at the start of the step; overwritten with the new state on the last stage
  float* y;        // (B,d) stage input (in) / next stage input (out); aliases x on stage 0
  float* ks;       // (B,d) running Runge-Kutta sum
  long long B;
  // device-side step clock (msgm_step_clock): all NULL / 0 for the host-clocked call
  const int* clock;
  const float* s_tab;
  float* s_next;
  float* traj;
  const int* keep_step;
  float* keep_out;
  int inc_t0;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (l < (blockDim.x >> 5)) ? red[l] : 0.0f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (l == 0) red[0] = v;
  }
  __syncthreads();
  v = red[0];
  __syncthreads();
  return v;
}

__global__ void __launch_bounds__(256) stage_update_kernel(const __grid_constant__ StageParams P) {
  __shared__ float red[32];
  const int d = P.d;
  const int nstage = P.scheme == MSGM_SCHEME_RK4 ? 4 : (P.scheme == MSGM_SCHEME_HEUN ? 2 : 1);
  const bool last = P.stage == nstage - 1;
  const bool ito = P.scheme == MSGM_SCHEME_EM;
  const float lm = P.lmbd, delta = P.delta;
  const int step = P.clock ? *P.clock : 0;
  const float s_cur = P.clock ? P.s_tab[step * nstage + P.stage] : P.s;
  const float bt = beta_of(P.bmin, P.bdel, s_cur), sb = sqrtf(bt);
  const float c_a = P.fwd ? 0.0f : delta * (1.0f - 0.5f * lm);
  const float c_w = P.fwd ? 1.0f : sqrtf(1.0f - lm);
  const float c_f = P.fwd ? (ito ? 1.0f : 0.0f) : (ito ? (1.0f - 2.0f * lm) : -lm);
  constexpr int MAXE = 16;  // elements per thread: d <= 4096
  for (long long row = blockIdx.x; row < P.B; row += gridDim.x) {
    const float* yr = (P.stage == 0 ? P.x : P.y) + row * d;
    const float* ar = P.a ? P.a + row * d : nullptr;
    const float* wr = P.dW + row * d;
    float* xr = P.x + row * d;
    float* yo = P.y + row * d;
    float* kr = P.ks + row * d;
    float Kv[MAXE];
    // phase 1: every thread reads its stencil neighbourhood and forms K in registers
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      const int c = threadIdx.x + e * 256;
      float K = 0.0f;
      if (c < d) {
        const float yc = yr[c];
        if (P.kind == MSGM_SDE_SGM) {
          if (P.fwd) K = delta * (-0.5f * bt * yc) + sb * wr[c];
          else K = delta * ((1.0f - 0.5f * lm) * (sb * ar[c]) + 0.5f * bt * yc) + (c_w * sb) * wr[c];
        } else {
          const int cn = (c + 1 == d) ? 0 : c + 1, cp = (c == 0) ? d - 1 : c - 1;
          const float wc = (ar ? c_a * ar[c] : 0.0f) + c_w * wr[c];
          const float wp = (ar ? c_a * ar[cp] : 0.0f) + c_w * wr[cp];
          K = (SQRT_HALF * (sb * yr[cn])) * wc + (-SQRT_HALF * (sb * yr[cp])) * wp;
          K = fmaf(delta * c_f, 0.5f * bt * yc, K);  // sparse f = +beta y / 2 (SDEs.py:412-413)
        }
      }
      Kv[e] = K;
    }
    __syncthreads();  // all reads of the stage input are done before y is overwritten in place
    // phase 2: Runge-Kutta bookkeeping (sde_scheme.py:86 | 147,156 | 232-253), radius re-pin on the last stage
    float sq = 0.0f;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
      const int c = threadIdx.x + e * 256;
      if (c < d) {
        const float K = Kv[e], xc = xr[c];
        float xn = 0.0f;
        if (nstage == 1) {
          xn = xc + K;
        } else if (nstage == 2) {
          if (P.stage == 0) { kr[c] = K; yo[c] = xc + K; }
          else xn = xc + (kr[c] + K) / 2.0f;
        } else {
          if (P.stage == 0) { kr[c] = K; yo[c] = xc + K / 2.0f; }
          else if (P.stage == 1) { kr[c] = kr[c] + 2.0f * K; yo[c] = xc + K / 2.0f; }
          else if (P.stage == 2) { kr[c] = kr[c] + 2.0f * K; yo[c] = xc + K; }
          else xn = xc + (kr[c] + K) / 6.0f;
        }
        Kv[e] = xn;
        sq = fmaf(xn, xn, sq);
      }
    }
    if (last) {
      float sc = 1.0f;
      if (P.nc) sc = P.r0[row] / sqrtf(block_sum(sq, red));
      float* tr = P.traj ? P.traj + ((long long)(step + P.inc_t0) * P.B + row) * d : nullptr;
      float* ko = (P.keep_step && P.keep_step[row] == step + P.inc_t0) ? P.keep_out + row * d : nullptr;
#pragma unroll
      for (int e = 0; e < MAXE; ++e) {
        const int c = threadIdx.x + e * 256;
        if (c < d) {
          const float xn = Kv[e] * sc;
          xr[c] = xn;
          if (tr) tr[c] = xn;
          if (ko) ko[c] = xn;
        }
      }
    }
    if (P.s_next && threadIdx.x == 0) P.s_next[row] = P.s_tab[step * nstage + P.stage + 1];  // time of the next net evaluation
    __syncthreads();
  }
}

// NOTE on aliasing: on stage 0 the stage input is x itself and y (a separate buffer) is only written; on later stages
// y is read in phase 1 and overwritten in place in phase 2, x is only read until the last stage writes the new state.

__global__ void __launch_bounds__(256) row_norm_kernel(const float* __restrict__ x, float* __restrict__ r, int d, long long B) {
  __shared__ float red[32];
  for (long long row = blockIdx.x; row < B; row += gridDim.x) {
    float sq = 0.0f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) sq = fmaf(x[row * d + c], x[row * d + c], sq);
    const float tot = block_sum(sq, red);
    if (threadIdx.x == 0) r[row] = sqrtf(tot);
  }
}

// dW = scale * N(0,1), same Philox keying as the fused samplers: (seed, global particle, step, component block)
// With a step clock the step index comes from device memory, injected noise (N,B,d) replaces Philox when given, and the
// noise time of the step's first net evaluation is written per row (s_first).
__global__ void __launch_bounds__(256) philox_normal_kernel(float* __restrict__ out, int d, long long B, float scale,
                                                            unsigned long long seed, unsigned long long poff, unsigned step_host,
                                                            const int* __restrict__ clock, const float* __restrict__ noise,
                                                            const float* __restrict__ s_tab, float* __restrict__ s_first,
                                                            int nstage) {
  const unsigned step = clock ? (unsigned)*clock : step_host;
  const long long nblk = (long long)((d + 3) / 4);
  const long long total = B * nblk;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / nblk;
    const int blk = (int)(i % nblk);
    float zz[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (blk * 4 + c < d) zz[c] = noise[((long long)step * B + row) * d + blk * 4 + c];
    } else {
      const float4 z = philox_normal4(seed, poff + (unsigned long long)row, step, (unsigned)blk);
      zz[0] = z.x; zz[1] = z.y; zz[2] = z.z; zz[3] = z.w;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (blk * 4 + c < d) out[row * d + blk * 4 + c] = scale * zz[c];
    if (s_first && blk == 0) s_first[row] = s_tab[step * nstage];
  }
}

__global__ void clock_advance_kernel(int* clock) { *clock += 1; }

int stage_update(msgm_ctx* ctx, const msgm_sde_desc* sde, int scheme, int stage, float lmbd, int nc, int fwd, float s,
                 float delta, const float* a, const float* dW, const float* r0, float* x, float* y, float* ks, int64_t B,
                 cudaStream_t stream, const msgm_step_clock* clk) {
  StageParams P{};
  P.kind = sde->kind;
  P.d = sde->dim;
  P.scheme = scheme;
  P.stage = stage;
  P.nc = nc;
  P.fwd = fwd;
  P.s = s;
  P.bmin = sde->beta_min;
  P.bdel = sde->beta_delta;
  P.delta = delta;
  P.lmbd = lmbd;
  P.a = a; P.dW = dW; P.r0 = r0; P.x = x; P.y = y; P.ks = ks; P.B = B;
  if (clk) {
    P.clock = clk->clock; P.s_tab = clk->s_table; P.s_next = clk->s_next; P.traj = clk->traj;
    P.keep_step = clk->keep_out ? clk->keep_step : nullptr; P.keep_out = clk->keep_out; P.inc_t0 = clk->include_t0;
  }
  const int grid = (int)std::min<long long>(B, (long long)ctx->num_sms * 8);
  stage_update_kernel<<<grid, 256, 0, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int row_norm(msgm_ctx* ctx, const float* x, float* r, int d, int64_t B, cudaStream_t stream) {
  row_norm_kernel<<<(int)std::min<long long>(B, (long long)ctx->num_sms * 8), 256, 0, stream>>>(x, r, d, B);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int philox_normal(msgm_ctx* ctx, float* out, int d, int64_t B, float scale, uint64_t seed, uint64_t poff, uint32_t step,
                  cudaStream_t stream, const msgm_step_clock* clk, const float* noise, int nstage) {
  const long long total = B * ((d + 3) / 4);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 16);
  philox_normal_kernel<<<grid, 256, 0, stream>>>(out, d, B, scale, seed, poff, step, clk ? clk->clock : nullptr, noise,
                                                 clk ? clk->s_table : nullptr, clk ? clk->s_next : nullptr, nstage);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int clock_advance(msgm_ctx* ctx, int32_t* clock, cudaStream_t stream) {
  clock_advance_kernel<<<1, 1, 0, stream>>>(clock);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
