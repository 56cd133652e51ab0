// Forward noising y_t | y_0 of the multiplicative SDE by simulation: SDE.sample_scheme (SDEs.py:78-122) =
// rk4_stratonovich_sampler (sde_scheme.py:174-269) on the forward_SDE adapter (SDEs.py:30-47), all num_steps_forward
// steps of every row in ONE launch.  This is the training-time prologue of the score-matching step.
//
// The forward Stratonovich drift of the MSGM SDE is zero (f_strato = 0), so a stage increment is pure diffusion:
//     K_i = sqrt(beta(t_s)) * sum_{j,k} G_ijk y_j dW_k.
// dW is shared by the four stages of a step (sde_scheme.py:227), so the k-contraction is hoisted out of the stage
// loop:  M_ij = sum_k G_ijk dW_k  once per step (d^3 flop per particle), then K = sqrt(beta) M y per stage (d^2).
// Work decomposition: DP = pow2(d) lanes per particle, lane i owns component i (row i of M, y_i); a stage gathers y
// with DP warp shuffles.  A batch of 256 rows is 256*DP threads spread over many SMs instead of 4 CTAs of the fused
// sampler kernel, and nothing but G lives in shared memory.
//
// Row semantics follow sampler_fp32.cu (the parity kernel): row k runs n_k = trunc(N t_k / T) steps of the common grid;
// a row with n_k == 0 takes one step of size t_k; noise is injected (parity tests, CUDA-graphed training) or Philox.
#include <algorithm>
#include <cmath>

#include "msgm_common.cuh"

namespace msgm {

struct NoiseParams {
  int d, N;
  float bmin, bdel, Tsde;
  float delta, delta_half, sqrt_delta;
  const float* G;             // dense: (d,d,d)
  const float* ts;            // (N+1) fp32 time grid of the reference (sde_scheme.py:201)
  const float* t_noise;       // (B) noise time of each row
  const float* noise;         // (N,B,d) or NULL
  const float* noise_single;  // (B,d) or NULL
  unsigned long long seed, poff;
  float* x;                   // (B,d) in/out
  long long B;
  // fused training prologue (msgm_ssm_prepare): draw t and the probe v in the same launch, read x from x_in
  const float* x_in;          // (B,d) clean samples, or NULL = x holds them
  float* t_out;               // (B) drawn noise times, or NULL = t_noise given
  float* v_out;               // (B,d) Hutchinson probe, or NULL
  int vtype;                  // MSGM_V_*
  float t_eps;
  const unsigned long long* seed_off;  // device counter added to the seed (CUDA-graph replays), or NULL
};

constexpr uint32_t STREAM_T = 0xFFFF0010u, STREAM_V = 0xFFFF0011u, STREAM_SGM = 0xFFFF0012u;

template <int DP>
struct GLayout {
  static constexpr int SLAB = DP * DP + 4;  // floats per row-slab i, padded so that the DP lanes hit distinct banks
  static constexpr size_t BYTES = sizeof(float) * DP * SLAB;
};

template <int DP, int KIND>
__global__ void __launch_bounds__(128) noise_forward_kernel(const __grid_constant__ NoiseParams P) {
  extern __shared__ __align__(16) float sG[];
  constexpr int SLAB = GLayout<DP>::SLAB;
  constexpr int PPB = 128 / DP;  // particles per block pass
  const int tid = threadIdx.x;
  const int i = tid % DP;        // component owned by this lane
  const int d = P.d;
  if (KIND == MSGM_SDE_MSGM_DENSE) {
    for (int e = tid; e < DP * DP * DP; e += 128) {
      const int a = e / (DP * DP), j = (e / DP) % DP, k = e % DP;
      sG[a * SLAB + j * DP + k] = (a < d && j < d && k < d) ? __ldg(P.G + ((long long)a * d + j) * d + k) : 0.0f;
    }
    __syncthreads();
  }
  const unsigned full = 0xffffffffu;
  const unsigned long long seed = P.seed + (P.seed_off ? *P.seed_off : 0ull);
  const long long ngroups = (P.B + PPB - 1) / PPB;
  for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
    const long long gp = grp * PPB + tid / DP;
    const bool live = gp < P.B;
    const bool mine = live && i < d;
    float x = mine ? (P.x_in ? __ldg(P.x_in + gp * d + i) : P.x[gp * d + i]) : 0.0f;
    float y = x, ks = 0.0f;

    // noise time: given, or t ~ U(0,T) floored at t_epsilon (PluginReverseSDE.sample_t, SDEs.py:684-693)
    float tk = 0.0f;
    if (P.t_out) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)(P.poff + gp), (uint32_t)((P.poff + gp) >> 32), STREAM_T, 0u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      tk = u01(r.x) * P.Tsde;
      tk = tk <= P.t_eps ? P.t_eps : tk;
      if (live && i == 0) P.t_out[gp] = tk;
    } else if (live) {
      tk = __ldg(P.t_noise + gp);
    }
    // Hutchinson probe (sample_v, SDEs.py:514-536)
    if (P.v_out) {
      const unsigned long long pid = P.poff + (unsigned long long)gp;
      float vi;
      if (P.vtype == MSGM_V_RADEMACHER) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), STREAM_V, (uint32_t)(i >> 2)),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const int q = i & 3;
        const uint32_t w = q == 0 ? r.x : (q == 1 ? r.y : (q == 2 ? r.z : r.w));
        vi = (w >> 31) ? 1.0f : -1.0f;
      } else {
        const float4 z = philox_normal4(seed, pid, STREAM_V, (uint32_t)(i >> 2));
        const int q = i & 3;
        vi = q == 0 ? z.x : (q == 1 ? z.y : (q == 2 ? z.z : z.w));
        if (P.vtype == MSGM_V_SPHERE) {  // X / |X| (randu_on_sphere, SDEs.py:520-526)
          float n2 = i < d ? vi * vi : 0.0f;
#pragma unroll
          for (int o = DP / 2; o > 0; o >>= 1) n2 += __shfl_xor_sync(full, n2, o, DP);
          vi = vi / sqrtf(n2);
        }
      }
      if (mine) P.v_out[gp * d + i] = vi;
    }
    if (KIND == MSGM_SDE_SGM) {
      // closed-form VP marginal (SDE.sample_Song_et_al, SDEs.py:134-146): y = mean_weight(t) x + sqrt(var(t)) eps
      const float4 z = philox_normal4(seed, P.poff + (unsigned long long)gp, STREAM_SGM, (uint32_t)(i >> 2));
      const int q = i & 3;
      const float eps = q == 0 ? z.x : (q == 1 ? z.y : (q == 2 ? z.z : z.w));
      const float e1 = expf(-0.25f * tk * tk * P.bdel - 0.5f * tk * P.bmin);
      const float var = 1.0f - expf(-0.5f * tk * tk * P.bdel - tk * P.bmin);
      if (mine) P.x[gp * d + i] = fmaf(eps, sqrtf(var), e1 * x);
      continue;
    }

    // row schedule (SDEs.py:86-118)
    float delta = P.delta, delta_half = P.delta_half, sqrt_delta = P.sqrt_delta;
    const int nk = tk >= P.Tsde ? P.N : (int)truncf(__fdiv_rn(__fmul_rn((float)P.N, tk), P.Tsde));
    const bool single = nk == 0;
    const int my_steps = single ? 1 : nk;
    if (single) {
      delta = tk;
      delta_half = (float)((double)tk * 0.5);
      sqrt_delta = (float)sqrt((double)tk);
    }
    // the warp walks as many steps as its longest row needs
    int steps_here = live ? my_steps : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps_here = max(steps_here, __shfl_xor_sync(full, steps_here, o));

    for (int step = 0; step < steps_here; ++step) {
      const bool active = step < my_steps;
      const float tcur = single ? 0.0f : (P.ts ? __ldg(P.ts + step) : __fmul_rn((float)step, delta));
      // ---- this lane's normal, then the whole increment vector by shuffles -------------------------------------
      float xi = 0.0f;
      if (i < d) {
        if (single && P.noise_single) {
          xi = (live && step == 0) ? __ldg(P.noise_single + gp * d + i) : 0.0f;
        } else if (P.noise && !single) {
          xi = live ? __ldg(P.noise + ((long long)step * P.B + gp) * d + i) : 0.0f;
        } else {
          const float4 z = philox_normal4(seed, P.poff + (unsigned long long)gp, single ? 0xFFFF0002u : (uint32_t)step,
                                          (uint32_t)(i >> 2));
          const int q = i & 3;
          xi = q == 0 ? z.x : (q == 1 ? z.y : (q == 2 ? z.z : z.w));
        }
      }
      const float dw = sqrt_delta * xi;

      [[maybe_unused]] float M[DP];
      [[maybe_unused]] float dw_prev = 0.0f;
      if (KIND == MSGM_SDE_MSGM_DENSE) {
        float dwv[DP];
#pragma unroll
        for (int k = 0; k < DP; ++k) dwv[k] = __shfl_sync(full, dw, k, DP);
        const float* Gi = sG + i * SLAB;
#pragma unroll
        for (int j = 0; j < DP; ++j) {
          float m = 0.0f;
          if constexpr (DP >= 4) {
#pragma unroll
            for (int k4 = 0; k4 < DP; k4 += 4) {
              const float4 g4 = *reinterpret_cast<const float4*>(Gi + j * DP + k4);
              m = fmaf(g4.x, dwv[k4], m);
              m = fmaf(g4.y, dwv[k4 + 1], m);
              m = fmaf(g4.z, dwv[k4 + 2], m);
              m = fmaf(g4.w, dwv[k4 + 3], m);
            }
          } else {
#pragma unroll
            for (int k = 0; k < DP; ++k) m = fmaf(Gi[j * DP + k], dwv[k], m);
          }
          M[j] = m;
        }
      } else {
        const int cp = (i == 0) ? d - 1 : i - 1;
        dw_prev = __shfl_sync(full, dw, cp & (DP - 1), DP);
      }

#pragma unroll
      for (int st = 0; st < 4; ++st) {
        float tst = tcur;
        if (st > 0) tst = st < 3 ? __fadd_rn(tcur, delta_half) : __fadd_rn(tcur, delta);
        const float sb = sqrtf(beta_of(P.bmin, P.bdel, tst));
        float K;
        if (KIND == MSGM_SDE_MSGM_DENSE) {
          float acc = 0.0f;
#pragma unroll
          for (int j = 0; j < DP; ++j) acc = fmaf(M[j], __shfl_sync(full, y, j, DP), acc);
          K = sb * acc;
        } else {
          // cyclic stencil of SDEs.py:369-399 with the runtime dimension d
          const int cn = (i + 1 == d) ? 0 : i + 1, cp = (i == 0) ? d - 1 : i - 1;
          const float yn = __shfl_sync(full, y, cn & (DP - 1), DP), yp = __shfl_sync(full, y, cp & (DP - 1), DP);
          K = (SQRT_HALF * (sb * yn)) * dw + (-SQRT_HALF * (sb * yp)) * dw_prev;
        }
        if (!active || i >= d) K = 0.0f;
        if (st == 0) { ks = K; y = x + K / 2.0f; }
        else if (st == 1) { ks = ks + 2.0f * K; y = x + K / 2.0f; }
        else if (st == 2) { ks = ks + 2.0f * K; y = x + K; }
        else { x = x + (ks + K) / 6.0f; }
      }
      y = x;
    }
    if (mine) P.x[gp * d + i] = x;
  }
}

// ---- large states (U-Net configs: d = 1000 / 1024, sparse cyclic tensor or SGM), d <= 4096 ----------------------------------
// One CTA per row, thread = groups of 4 consecutive components (one Philox4x32 call = their four normals), state x / y /
// running RK sum / dW in registers for all steps.  The cyclic 3-point stencil needs y_{c+1}, y_{c-1} and dW_{c-1}: inside a
// group they are registers, across groups they come from a double-buffered copy of y in shared memory (one barrier per
// stage).  This replaces N_fwd x (4 stage launches + 1 noise launch) = 640 launches per training iteration at N_fwd = 128.
constexpr int NB_THREADS = 256, NB_MAXG = 4;  // 256 threads x 4 groups x 4 components = 4096

__device__ __forceinline__ float nb_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  v = l < NB_THREADS / 32 ? red[l] : 0.0f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int KIND>
__global__ void __launch_bounds__(NB_THREADS) noise_forward_big_kernel(const __grid_constant__ NoiseParams P) {
  extern __shared__ __align__(16) float sbig[];  // [2][d] stage inputs y, [d] dW
  __shared__ float red[NB_THREADS / 32];
  const int d = P.d, tid = threadIdx.x;
  float* sy = sbig;
  float* sdw = sbig + 2 * d;
  const unsigned long long seed = P.seed + (P.seed_off ? *P.seed_off : 0ull);
  for (long long gp = blockIdx.x; gp < P.B; gp += gridDim.x) {
    const unsigned long long pid = P.poff + (unsigned long long)gp;
    float x[NB_MAXG][4], y[NB_MAXG][4], ks[NB_MAXG][4], dw[NB_MAXG][4];
#pragma unroll
    for (int e = 0; e < NB_MAXG; ++e)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * (tid + e * NB_THREADS) + k;
        x[e][k] = c < d ? (P.x_in ? __ldg(P.x_in + gp * d + c) : P.x[gp * d + c]) : 0.0f;
        y[e][k] = x[e][k];
        ks[e][k] = 0.0f;
        dw[e][k] = 0.0f;
      }
    // noise time: given, or t ~ U(0,T) floored at t_epsilon (PluginReverseSDE.sample_t, SDEs.py:684-693)
    float tk;
    if (P.t_out) {
      const uint4 r = philox4x32_10(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), STREAM_T, 0u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      tk = u01(r.x) * P.Tsde;
      tk = tk <= P.t_eps ? P.t_eps : tk;
      if (tid == 0) P.t_out[gp] = tk;
    } else {
      tk = __ldg(P.t_noise + gp);
    }
    // Hutchinson probe (sample_v, SDEs.py:514-536)
    if (P.v_out) {
      float vv[NB_MAXG][4], n2 = 0.0f;
#pragma unroll
      for (int e = 0; e < NB_MAXG; ++e) {
        const int g = tid + e * NB_THREADS;
        if (4 * g >= d) continue;
        if (P.vtype == MSGM_V_RADEMACHER) {
          const uint4 r = philox4x32_10(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), STREAM_V, (uint32_t)g),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
          vv[e][0] = (r.x >> 31) ? 1.0f : -1.0f; vv[e][1] = (r.y >> 31) ? 1.0f : -1.0f;
          vv[e][2] = (r.z >> 31) ? 1.0f : -1.0f; vv[e][3] = (r.w >> 31) ? 1.0f : -1.0f;
        } else {
          const float4 z = philox_normal4(seed, pid, STREAM_V, (uint32_t)g);
          vv[e][0] = z.x; vv[e][1] = z.y; vv[e][2] = z.z; vv[e][3] = z.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < d) n2 = fmaf(vv[e][k], vv[e][k], n2);
      }
      float sc = 1.0f;
      if (P.vtype == MSGM_V_SPHERE) sc = rsqrtf(nb_block_sum(n2, red));  // X / |X| (randu_on_sphere, SDEs.py:520-526)
#pragma unroll
      for (int e = 0; e < NB_MAXG; ++e)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 4 * (tid + e * NB_THREADS) + k;
          if (c < d) P.v_out[gp * d + c] = vv[e][k] * sc;
        }
    }
    if (KIND == MSGM_SDE_SGM) {
      // closed-form VP marginal (SDE.sample_Song_et_al, SDEs.py:134-146): y = mean_weight(t) x + sqrt(var(t)) eps
      const float e1 = expf(-0.25f * tk * tk * P.bdel - 0.5f * tk * P.bmin);
      const float sd = sqrtf(1.0f - expf(-0.5f * tk * tk * P.bdel - tk * P.bmin));
#pragma unroll
      for (int e = 0; e < NB_MAXG; ++e) {
        const int g = tid + e * NB_THREADS;
        if (4 * g >= d) continue;
        const float4 z = philox_normal4(seed, pid, STREAM_SGM, (uint32_t)g);
        const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (4 * g + k < d) P.x[gp * d + 4 * g + k] = fmaf(zz[k], sd, e1 * x[e][k]);
      }
      continue;
    }

    // row schedule (SDEs.py:86-118): n_k steps of the common grid, or one step of size t_k
    float delta = P.delta, delta_half = P.delta_half, sqrt_delta = P.sqrt_delta;
    const int nk = tk >= P.Tsde ? P.N : (int)truncf(__fdiv_rn(__fmul_rn((float)P.N, tk), P.Tsde));
    const bool single = nk == 0;
    const int my_steps = single ? 1 : nk;
    if (single) {
      delta = tk;
      delta_half = (float)((double)tk * 0.5);
      sqrt_delta = (float)sqrt((double)tk);
    }
    int buf = 0;
    __syncthreads();  // the previous row's last reads of sy / sdw are done
    for (int step = 0; step < my_steps; ++step) {
      const float tcur = single ? 0.0f : (P.ts ? __ldg(P.ts + step) : __fmul_rn((float)step, delta));
#pragma unroll
      for (int e = 0; e < NB_MAXG; ++e) {
        const int g = tid + e * NB_THREADS;
        if (4 * g >= d) continue;
        float z4[4];
        if (single && P.noise_single) {
#pragma unroll
          for (int k = 0; k < 4; ++k) z4[k] = 4 * g + k < d ? __ldg(P.noise_single + gp * d + 4 * g + k) : 0.0f;
        } else if (P.noise && !single) {
#pragma unroll
          for (int k = 0; k < 4; ++k) z4[k] = 4 * g + k < d ? __ldg(P.noise + ((long long)step * P.B + gp) * d + 4 * g + k) : 0.0f;
        } else {
          const float4 z = philox_normal4(seed, pid, single ? 0xFFFF0002u : (uint32_t)step, (uint32_t)g);
          z4[0] = z.x; z4[1] = z.y; z4[2] = z.z; z4[3] = z.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          dw[e][k] = sqrt_delta * z4[k];
          if (4 * g + k < d) {
            sdw[4 * g + k] = dw[e][k];
            sy[buf * d + 4 * g + k] = y[e][k];
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        float tst = tcur;
        if (st > 0) tst = st < 3 ? __fadd_rn(tcur, delta_half) : __fadd_rn(tcur, delta);
        const float sb = sqrtf(beta_of(P.bmin, P.bdel, tst));
        const float* yin = sy + buf * d;
        float* yout = sy + (buf ^ 1) * d;
#pragma unroll
        for (int e = 0; e < NB_MAXG; ++e) {
          const int g = tid + e * NB_THREADS;
          if (4 * g >= d) continue;
          float ynew[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = 4 * g + k;
            ynew[k] = 0.0f;
            if (c >= d) continue;
            // cyclic stencil of SDEs.py:369-399: K_c = sqrt(beta / 2) (y_{c+1} dW_c - y_{c-1} dW_{c-1})
            const float yn = (k < 3 && c + 1 < d) ? y[e][k + 1] : yin[c + 1 == d ? 0 : c + 1];
            const float yp = k > 0 ? y[e][k - 1] : yin[c == 0 ? d - 1 : c - 1];
            const float wp = k > 0 ? dw[e][k - 1] : sdw[c == 0 ? d - 1 : c - 1];
            const float K = (SQRT_HALF * (sb * yn)) * dw[e][k] + (-SQRT_HALF * (sb * yp)) * wp;
            if (st == 0) { ks[e][k] = K; ynew[k] = x[e][k] + K / 2.0f; }
            else if (st == 1) { ks[e][k] = ks[e][k] + 2.0f * K; ynew[k] = x[e][k] + K / 2.0f; }
            else if (st == 2) { ks[e][k] = ks[e][k] + 2.0f * K; ynew[k] = x[e][k] + K; }
            else { x[e][k] = x[e][k] + (ks[e][k] + K) / 6.0f; ynew[k] = x[e][k]; }
            yout[c] = ynew[k];
          }
          // the registers take the next stage input only after the whole group was evaluated with the old one
#pragma unroll
          for (int k = 0; k < 4; ++k) y[e][k] = ynew[k];
        }
        buf ^= 1;
        __syncthreads();
      }
    }
#pragma unroll
    for (int e = 0; e < NB_MAXG; ++e)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * (tid + e * NB_THREADS) + k;
        if (c < d) P.x[gp * d + c] = x[e][k];
      }
  }
}

template <int KIND>
static int launch_noise_big(msgm_ctx* ctx, const NoiseParams& P, cudaStream_t stream) {
  const size_t smem = sizeof(float) * 3 * (size_t)P.d;
  auto kern = noise_forward_big_kernel<KIND>;
  if (smem > 48 * 1024) MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(int)std::min<long long>(P.B, (long long)ctx->num_sms * 8), NB_THREADS, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int DP, int KIND>
static int launch_noise(msgm_ctx* ctx, const NoiseParams& P, cudaStream_t stream) {
  auto kern = noise_forward_kernel<DP, KIND>;
  const size_t smem = KIND == MSGM_SDE_MSGM_DENSE ? GLayout<DP>::BYTES : 0;
  if (smem > 48 * 1024) MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ngroups = (P.B + 128 / DP - 1) / (128 / DP);
  // G is re-staged per CTA: cap the grid where that staging would dominate (d > 8), otherwise one pass per CTA
  const long long cap = (long long)ctx->num_sms * (DP >= 32 ? 1 : (DP >= 16 ? 8 : 64));
  kern<<<(int)std::min<long long>(ngroups, cap), 128, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int DP>
static int launch_noise_kind(msgm_ctx* ctx, int kind, const NoiseParams& P, cudaStream_t stream) {
  if (kind == MSGM_SDE_MSGM_DENSE) return launch_noise<DP, MSGM_SDE_MSGM_DENSE>(ctx, P, stream);
  if (kind == MSGM_SDE_SGM) return launch_noise<DP, MSGM_SDE_SGM>(ctx, P, stream);
  return launch_noise<DP, MSGM_SDE_MSGM_SPARSE>(ctx, P, stream);
}

static int run_noise(msgm_ctx* ctx, const msgm_sde_desc* sde, NoiseParams& P, int num_steps, cudaStream_t stream) {
  const int d = sde->dim;
  P.d = d;
  P.N = num_steps;
  P.bmin = sde->beta_min;
  P.bdel = sde->beta_delta;
  P.Tsde = sde->T;
  // delta = T/N is a Python double in the reference (sde_scheme.py:200), rounded to fp32 where it meets tensors
  const double delta = (double)sde->T / (double)num_steps;
  P.delta = (float)delta;
  P.delta_half = (float)(delta / 2.0);
  P.sqrt_delta = (float)std::sqrt(delta);
  P.G = sde->G;
  if (d > 32) {  // U-Net sized states: the one-CTA-per-row kernel (sparse multiplicative SDE or SGM)
    if (sde->kind == MSGM_SDE_MSGM_SPARSE) return launch_noise_big<MSGM_SDE_MSGM_SPARSE>(ctx, P, stream);
    if (sde->kind == MSGM_SDE_SGM) return launch_noise_big<MSGM_SDE_SGM>(ctx, P, stream);
    set_error("forward noising of a dense tensor is built for d <= 32 (O(d^3) per step)");
    return MSGM_ERR_UNSUPPORTED;
  }
  const int DP = d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32;
  switch (DP) {
    case 2: return launch_noise_kind<2>(ctx, sde->kind, P, stream);
    case 4: return launch_noise_kind<4>(ctx, sde->kind, P, stream);
    case 8: return launch_noise_kind<8>(ctx, sde->kind, P, stream);
    case 16: return launch_noise_kind<16>(ctx, sde->kind, P, stream);
    default: return launch_noise_kind<32>(ctx, sde->kind, P, stream);
  }
}

int noise_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* t, float* y_inout, int num_steps, const float* ts,
                  const float* noise, const float* noise_single, uint64_t seed, uint64_t poff, int64_t B,
                  cudaStream_t stream) {
  NoiseParams P{};
  P.ts = ts;
  P.t_noise = t;
  P.noise = noise;
  P.noise_single = noise_single;
  P.seed = seed;
  P.poff = poff;
  P.x = y_inout;
  P.B = B;
  return run_noise(ctx, sde, P, num_steps, stream);
}

// Training prologue in one launch: t ~ U(0,T) floored at t_epsilon, the Hutchinson probe v, and y_t | x.
int ssm_prepare(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* x, float* t_out, float* v_out, float* y_out,
                int num_steps, const float* ts, float t_eps, int vtype, uint64_t seed, const uint64_t* seed_off,
                uint64_t poff, int64_t B, cudaStream_t stream) {
  NoiseParams P{};
  P.ts = ts;
  P.seed = seed;
  P.seed_off = reinterpret_cast<const unsigned long long*>(seed_off);
  P.poff = poff;
  P.x_in = x;
  P.x = y_out;
  P.t_out = t_out;
  P.v_out = v_out;
  P.vtype = vtype;
  P.t_eps = t_eps;
  P.B = B;
  return run_noise(ctx, sde, P, num_steps, stream);
}

}  // namespace msgm
