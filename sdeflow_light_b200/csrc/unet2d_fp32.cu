// 2-D U-Net score-net layers (NNUnet.py:80-245, model/unet.py:40-517, model/nn_utils.py:39-148), fp32 CUDA cores, NCHW.
//
// The reference runs, per ResBlock, GroupNorm -> SiLU -> conv3x3 -> (+ Linear(SiLU(emb))) -> GroupNorm -> SiLU -> conv3x3
// -> (+ skip) as ~12 separate library kernels that each stream the activation through HBM.  Here:
//   gn_stats_kernel     per-(sample, group) mean / rstd of the (possibly concatenated) conv input;
//   conv2d_kernel       conv 3x3 (stride 1|2) or 1x1 over the channel concat [x1, x2] read in place, with the
//                       normalise (+SiLU) applied while staging the input tile (the normalised tensor is never
//                       written), optional nearest x2 upsampling folded into the input indexing (Upsample), and
//                       bias + per-(sample,channel) embedding term + residual tensor fused in the epilogue;
//   emb_proj_kernel     Linear(SiLU(emb)) of ResBlock.emb_layers;  sincos_embed_mlp_kernel  timestep_embedding + time_embed
//                       / scale_embed MLPs;  attention_kernel  single-head softmax(q k^T / sqrt(C)) v (T <= 1024);
//   vort_pre_kernel     NormalizeLogRadius * sqrt(d) / 5 with C- or F-order reshape, vort_post_kernel the inverse (x5).
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

__device__ __forceinline__ float siluf(float x) { return x / (1.0f + expf(-x)); }

// ---- GroupNorm statistics ---------------------------------------------------------------------------------------
// stats[(b*G + g)*2 + {0,1}] = mean, rstd over channels [g*cpg, (g+1)*cpg) x HW of the concat [x1 (C1), x2 (C2)]
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                       int C2, int HW, int G, float eps, float* __restrict__ stats) {
  __shared__ float rs[8], rq[8];
  const int b = blockIdx.x / G, g = blockIdx.x % G, cpg = (C1 + C2) / G, tid = threadIdx.x;
  const int n = cpg * HW;
  float s = 0.0f, q = 0.0f;
  for (int e = tid; e < n; e += 256) {
    const int c = g * cpg + e / HW, p = e % HW;
    const float v = c < C1 ? x1[((size_t)b * C1 + c) * HW + p] : x2[((size_t)b * C2 + (c - C1)) * HW + p];
    s += v;
    q = fmaf(v, v, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  if ((tid & 31) == 0) { rs[tid >> 5] = s; rq[tid >> 5] = q; }
  __syncthreads();
  if (tid == 0) {
    float ts = 0.0f, tq = 0.0f;
    for (int w = 0; w < 8; ++w) { ts += rs[w]; tq += rq[w]; }
    const float mean = ts / n, var = fmaxf(tq / n - mean * mean, 0.0f);
    stats[blockIdx.x * 2] = mean;
    stats[blockIdx.x * 2 + 1] = rsqrtf(var + eps);
  }
}

// ---- convolution -----------------------------------------------------------------------------------------------------
constexpr int C2_CO = 32, C2_P = 128, C2_CI = 8, C2_STAGE = 17 * 34;  // tile: 32 out-channels x 128 positions

struct Conv2dParams {
  const float* x1; int C1;
  const float* x2; int C2;
  const float* W;        // (Cout, C1+C2, K, K)
  const float* bias;     // (Cout) or NULL
  const float* ebias;    // (B, Cout) or NULL
  const float* res;      // (B, Cout, Ho, Wo) or NULL
  const float* stats;    // (B, G, 2) or NULL
  const float* gamma;    // (C1+C2)
  const float* beta;
  float* out;
  int G, prologue;       // 0 none, 1 GroupNorm, 2 GroupNorm + SiLU
  int Cout, K, stride, up, Hs, Ws, Ho, Wo;  // Hs, Ws: stored input size; the conv sees (Hs*up, Ws*up)
};

__global__ void __launch_bounds__(256) conv2d_kernel(const __grid_constant__ Conv2dParams P) {
  __shared__ float sx[C2_CI][C2_STAGE];
  __shared__ float sw[C2_CI][9][C2_CO + 1];
  const int tid = threadIdx.x, tl = tid & 31, tc = tid >> 5;
  const int p0 = blockIdx.x * C2_P, co0 = blockIdx.y * C2_CO, b = blockIdx.z;
  const int Hi = P.Hs * P.up, Wi = P.Ws * P.up, pad = P.K == 3 ? 1 : 0, KK = P.K * P.K;
  const int HWo = P.Ho * P.Wo, HWs = P.Hs * P.Ws;
  // input rows covered by this tile of output positions
  const int oy0 = p0 / P.Wo, oy1 = min(P.Ho - 1, (p0 + C2_P - 1) / P.Wo);
  const int iy0 = oy0 * P.stride - pad, rows = (oy1 - oy0) * P.stride + P.K, wcols = Wi + 2 * pad;
  int oy[4], ox[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = min(p0 + tl + 32 * j, HWo - 1);
    oy[j] = p / P.Wo - oy0;
    ox[j] = p % P.Wo;
  }
  float acc[4][4] = {};
  const int Cin = P.C1 + P.C2, cpg = P.G > 0 ? Cin / P.G : 1;
  for (int c0 = 0; c0 < Cin; c0 += C2_CI) {
    __syncthreads();
    for (int e = tid; e < C2_CI * rows * wcols; e += 256) {
      const int ci = e / (rows * wcols), r = (e / wcols) % rows, cx = e % wcols, c = c0 + ci;
      const int iy = iy0 + r, ix = cx - pad;
      float v = 0.0f;
      if (c < Cin && iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) {
        const int sy = iy / P.up, sxx = ix / P.up;
        v = c < P.C1 ? P.x1[((size_t)b * P.C1 + c) * HWs + sy * P.Ws + sxx]
                     : P.x2[((size_t)b * P.C2 + (c - P.C1)) * HWs + sy * P.Ws + sxx];
        if (P.prologue) {
          const float mean = P.stats[((size_t)b * P.G + c / cpg) * 2], rstd = P.stats[((size_t)b * P.G + c / cpg) * 2 + 1];
          v = fmaf((v - mean) * rstd, P.gamma[c], P.beta[c]);
          if (P.prologue == 2) v = siluf(v);
        }
      }
      sx[ci][r * wcols + cx] = v;
    }
    for (int e = tid; e < C2_CI * KK * C2_CO; e += 256) {
      const int co = e % C2_CO, k = (e / C2_CO) % KK, ci = e / (C2_CO * KK), c = c0 + ci;
      sw[ci][k][co] = (c < Cin && co0 + co < P.Cout) ? P.W[((size_t)(co0 + co) * Cin + c) * KK + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < C2_CI; ++ci) {
      for (int ky = 0; ky < P.K; ++ky)
        for (int kx = 0; kx < P.K; ++kx) {
          float wv[4], xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[i] = sw[ci][ky * P.K + kx][tc * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[j] = sx[ci][(oy[j] * P.stride + ky) * wcols + ox[j] * P.stride + kx];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tc * 4 + i;
    if (co >= P.Cout) continue;
    float add = P.bias ? P.bias[co] : 0.0f;
    if (P.ebias) add += P.ebias[(size_t)b * P.Cout + co];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = p0 + tl + 32 * j;
      if (p >= HWo) continue;
      const size_t o = ((size_t)b * P.Cout + co) * HWo + p;
      float v = acc[i][j] + add;
      if (P.res) v += P.res[o];
      P.out[o] = v;
    }
  }
}

// out[b,co] = bias[co] + sum_i W[co,i] silu(emb[b,i])   (ResBlock.emb_layers, model/unet.py:146-152)
__global__ void __launch_bounds__(128) emb_proj_kernel(const float* __restrict__ emb, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ out, int E,
                                                       int Cout, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * Cout) return;
  const int co = idx % Cout, b = idx / Cout;
  float s = bias[co];
  for (int i = 0; i < E; ++i) s = fmaf(W[(size_t)co * E + i], siluf(emb[(size_t)b * E + i]), s);
  out[idx] = s;
}

// out (B,E) (+)= W2 silu(W1 sincos(t) + b1) + b2 ; sincos(t) = [cos(t w_k), sin(t w_k)], w_k = 1e4^(-k/half), dim = 2 half
// (timestep_embedding model/nn_utils.py:130-148 + time_embed / scale_embed model/unet.py:338-342, NNUnet.py:88-106)
__global__ void __launch_bounds__(256) sincos_embed_mlp_kernel(const float* __restrict__ t, const float* __restrict__ W1,
                                                               const float* __restrict__ b1, const float* __restrict__ W2,
                                                               const float* __restrict__ b2, float* __restrict__ out,
                                                               int dim, int E, int accumulate) {
  __shared__ float se[64], sh[256];
  const int b = blockIdx.x, tid = threadIdx.x, half = dim / 2;
  if (tid < half) {
    const float freq = expf(-logf(10000.0f) * (float)tid / (float)half);
    const float ang = t[b] * freq;
    se[tid] = cosf(ang);
    se[half + tid] = sinf(ang);
  }
  __syncthreads();
  if (tid < E) {
    float s = b1[tid];
    for (int i = 0; i < dim; ++i) s = fmaf(W1[(size_t)tid * dim + i], se[i], s);
    sh[tid] = siluf(s);
  }
  __syncthreads();
  if (tid < E) {
    float s = b2[tid];
    for (int i = 0; i < E; ++i) s = fmaf(W2[(size_t)tid * E + i], sh[i], s);
    out[(size_t)b * E + tid] = accumulate ? out[(size_t)b * E + tid] + s : s;
  }
}

// ---- single-head attention over T = H*W tokens (QKVAttention, model/unet.py:236-250) ---------------------------------
// qkv (B, 3C, T) -> out (B, C, T).  One CTA per (sample, 16 queries).
constexpr int AT_Q = 16;
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int C, int T) {
  extern __shared__ float sm[];
  float* sq = sm;                 // [C][AT_Q]
  float* sc = sm + C * AT_Q;      // [AT_Q][T]
  const int b = blockIdx.y, t0 = blockIdx.x * AT_Q, tid = threadIdx.x;
  const float* q = qkv + (size_t)b * 3 * C * T;
  const float* k = q + (size_t)C * T;
  const float* v = k + (size_t)C * T;
  const float scale2 = 1.0f / sqrtf((float)C);  // (1/sqrt(sqrt(C)))^2: both q and k are scaled in the reference
  for (int e = tid; e < C * AT_Q; e += 256) {
    const int c = e / AT_Q, tq = e % AT_Q;
    sq[e] = t0 + tq < T ? q[(size_t)c * T + t0 + tq] : 0.0f;
  }
  __syncthreads();
  const int tq = tid >> 4, sl = tid & 15;
  for (int s = sl; s < T; s += 16) {
    float d = 0.0f;
    for (int c = 0; c < C; ++c) d = fmaf(sq[c * AT_Q + tq], k[(size_t)c * T + s], d);
    sc[tq * T + s] = d * scale2;
  }
  __syncthreads();
  // softmax over s for each query row: 16 threads per row
  float mx = -INFINITY;
  for (int s = sl; s < T; s += 16) mx = fmaxf(mx, sc[tq * T + s]);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.0f;
  for (int s = sl; s < T; s += 16) {
    const float e = expf(sc[tq * T + s] - mx);
    sc[tq * T + s] = e;
    sum += e;
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  __syncthreads();
  for (int c = sl; c < C; c += 16) {
    float a = 0.0f;
    for (int s = 0; s < T; ++s) a = fmaf(sc[tq * T + s], v[(size_t)c * T + s], a);
    if (t0 + tq < T) out[((size_t)b * C + c) * T + t0 + tq] = a * inv;
  }
}

// ---- VorticityUNet wrapper (NNUnet.py:26-77,195-245) -------------------------------------------------------------------
// img[b,0,h,w] = x[b, idx(h,w)] * pre_scale[b] / 5 ; F order: idx = w*H + h ; lognorm[b] = log(|x_b| + eps)
__global__ void __launch_bounds__(256) vort_pre_kernel(const float* __restrict__ x, float* __restrict__ img,
                                                       float* __restrict__ lognorm, int H, int W, int forder, int pre) {
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x, d = H * W;
  if (pre) {
    float sq = 0.0f;
    for (int e = tid; e < d; e += 256) sq = fmaf(x[(size_t)b * d + e], x[(size_t)b * d + e], sq);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((tid & 31) == 0) red[tid >> 5] = sq;
    __syncthreads();
    float tot = 0.0f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float rn = sqrtf(tot) + 1e-6f;
    if (tid == 0) lognorm[b] = logf(rn);
    for (int e = tid; e < d; e += 256) {
      const int h = e / W, w = e % W;
      img[(size_t)b * d + e] = x[(size_t)b * d + (forder ? w * H + h : e)] / rn * sqrtf((float)d) / 5.0f;
    }
    return;
  }
  for (int e = tid; e < d; e += 256) {
    const int h = e / W, w = e % W;
    img[(size_t)b * d + e] = x[(size_t)b * d + (forder ? w * H + h : e)] / 5.0f;
  }
}

__global__ void __launch_bounds__(256) vort_post_kernel(const float* __restrict__ img, float* __restrict__ y, int H, int W,
                                                        int forder, long long total) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i >= total) return;
  const int d = H * W, e = (int)(i % d);
  const long long b = i / d;
  const int h = e / W, w = e % W;
  y[b * d + (forder ? w * H + h : e)] = 5.0f * img[i];
}

// ---- host wrappers -----------------------------------------------------------------------------------------------------
int gn_stats(msgm_ctx* ctx, const float* x1, int C1, const float* x2, int C2, int HW, int G, int B, float* stats,
             cudaStream_t stream) {
  gn_stats_kernel<<<B * G, 256, 0, stream>>>(x1, C1, x2, x2 ? C2 : 0, HW, G, 1e-5f, stats);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int conv2d(msgm_ctx* ctx, const msgm_conv2d_desc* D, cudaStream_t stream) {
  Conv2dParams P{};
  P.x1 = D->x1; P.C1 = D->C1; P.x2 = D->x2; P.C2 = D->x2 ? D->C2 : 0;
  P.W = D->W; P.bias = D->bias; P.ebias = D->ebias; P.res = D->res; P.stats = D->stats; P.gamma = D->gamma; P.beta = D->beta;
  P.out = D->out; P.G = D->G; P.prologue = D->prologue; P.Cout = D->Cout; P.K = D->K; P.stride = D->stride; P.up = D->up;
  P.Hs = D->Hs; P.Ws = D->Ws;
  const int pad = D->K == 3 ? 1 : 0;
  P.Ho = (D->Hs * D->up + 2 * pad - D->K) / D->stride + 1;
  P.Wo = (D->Ws * D->up + 2 * pad - D->K) / D->stride + 1;
  dim3 grid((P.Ho * P.Wo + C2_P - 1) / C2_P, (D->Cout + C2_CO - 1) / C2_CO, D->B);
  conv2d_kernel<<<grid, 256, 0, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int emb_proj(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int E, int Cout, int B,
             cudaStream_t stream) {
  emb_proj_kernel<<<(B * Cout + 127) / 128, 128, 0, stream>>>(emb, W, bias, out, E, Cout, B);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int sincos_embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2,
                     float* out, int B, int dim, int E, int accumulate, cudaStream_t stream) {
  sincos_embed_mlp_kernel<<<B, 256, 0, stream>>>(t, W1, b1, W2, b2, out, dim, E, accumulate);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int attention(msgm_ctx* ctx, const float* qkv, float* out, int B, int C, int T, cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)C * AT_Q + (size_t)AT_Q * T);
  if (smem > 48 * 1024)
    MSGM_CUDA_TRY(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_kernel<<<dim3((T + AT_Q - 1) / AT_Q, B), 256, smem, stream>>>(qkv, out, C, T);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int vort_pre(msgm_ctx* ctx, const float* x, float* img, float* lognorm, int B, int H, int W, int forder, int pre,
             cudaStream_t stream) {
  vort_pre_kernel<<<B, 256, 0, stream>>>(x, img, lognorm, H, W, forder, pre);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int vort_post(msgm_ctx* ctx, const float* img, float* y, int B, int H, int W, int forder, cudaStream_t stream) {
  const long long total = (long long)B * H * W;
  vort_post_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(img, y, H, W, forder, total);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
