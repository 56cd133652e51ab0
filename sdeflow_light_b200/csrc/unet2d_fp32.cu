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

// Thread = a strip of 4 consecutive output columns of one output row x NCO output channels: the three horizontal taps of
// a 3x3 filter share the 6 (stride 1) / 9 (stride 2) staged inputs of the strip, weights come as broadcast float4s.
// CTA = 32 strips (128 output positions, row-major) x 8 channel groups (COT = 32 or 64 output channels); CI input channels
// are staged per shared-memory round (8 for 3x3, 32 for 1x1 where a channel's tile is only 128 floats).
template <int K, int STRIDE, int COT, int CI>
__global__ void __launch_bounds__(256) conv2d_kernel(const __grid_constant__ Conv2dParams P) {
  constexpr int KK = K * K, PAD = K == 3 ? 1 : 0, NIN = 3 * STRIDE + K, NCO = COT / 8;
  constexpr int XS = K == 3 ? C2_STAGE : 160;  // floats staged per input channel
  __shared__ __align__(16) float sx[CI * XS];
  __shared__ __align__(16) float sw[CI][KK][COT + 4];  // +4: the transposing store is 4-way instead of 32-way conflicted
  const int tid = threadIdx.x, strip = tid & 31, cg = tid >> 5;
  const int p0 = blockIdx.x * C2_P, co0 = blockIdx.y * COT, b = blockIdx.z;
  const int Hi = P.Hs * P.up, Wi = P.Ws * P.up;
  const int HWo = P.Ho * P.Wo, HWs = P.Hs * P.Ws;
  const int oy0 = p0 / P.Wo, oy1 = min(P.Ho - 1, (p0 + C2_P - 1) / P.Wo);
  const int iy0 = oy0 * STRIDE - PAD, rows = (oy1 - oy0) * STRIDE + K, wcols = Wi + 2 * PAD;
  const int pstrip = p0 + strip * 4;                       // first output position of this strip (Wo % 4 == 0)
  const bool live = pstrip < HWo;
  const int oy = live ? pstrip / P.Wo - oy0 : 0, ox = live ? pstrip % P.Wo : 0;
  const int xbase = (oy * STRIDE) * wcols + ox * STRIDE;
  float acc[NCO][4] = {};
  const int Cin = P.C1 + P.C2, cpg = P.G > 0 ? Cin / P.G : 1;
  for (int c0 = 0; c0 < Cin; c0 += CI) {
    __syncthreads();
    // ---- stage CI input channels: rows x wcols patch, normalise(+SiLU) fused, zero padding after the activation
    for (int cr = tid / 32; cr < CI * rows; cr += 8) {   // one warp per (channel, row)
      const int ci = cr / rows, r = cr % rows, c = c0 + ci, iy = iy0 + r;
      const bool rowok = c < Cin && iy >= 0 && iy < Hi;
      float mean = 0.f, rstd = 1.f, ga = 1.f, be = 0.f;
      const float* src = nullptr;
      if (rowok) {
        src = (c < P.C1 ? P.x1 + ((size_t)b * P.C1 + c) * HWs : P.x2 + ((size_t)b * P.C2 + (c - P.C1)) * HWs) + (iy / P.up) * P.Ws;
        if (P.prologue) {
          mean = P.stats[((size_t)b * P.G + c / cpg) * 2];
          rstd = P.stats[((size_t)b * P.G + c / cpg) * 2 + 1];
          ga = P.gamma[c];
          be = P.beta[c];
        }
      }
      for (int cx = tid & 31; cx < wcols; cx += 32) {
        const int ix = cx - PAD;
        float v = 0.0f;
        if (rowok && ix >= 0 && ix < Wi) {
          v = src[ix / P.up];
          if (P.prologue) {
            v = fmaf((v - mean) * rstd, ga, be);
            if (P.prologue == 2) v = siluf(v);
          }
        }
        sx[ci * XS + r * wcols + cx] = v;
      }
    }
    for (int e = tid; e < CI * KK * COT; e += 256) {
      const int k = e % KK, ci = (e / KK) % CI, co = e / (KK * CI), c = c0 + ci;  // k fastest: coalesced over W
      sw[ci][k][co] = (c < Cin && co0 + co < P.Cout) ? P.W[((size_t)(co0 + co) * Cin + c) * KK + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        float xin[NIN];
#pragma unroll
        for (int q = 0; q < NIN; ++q) xin[q] = sx[ci * XS + xbase + ky * wcols + q];
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
#pragma unroll
          for (int h4 = 0; h4 < NCO / 4; ++h4) {
            const float4 w4 = *reinterpret_cast<const float4*>(&sw[ci][ky * K + kx][cg * NCO + h4 * 4]);
            const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[h4 * 4 + i][j] = fmaf(wv[i], xin[j * STRIDE + kx], acc[h4 * 4 + i][j]);
          }
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int i = 0; i < NCO; ++i) {
    const int co = co0 + cg * NCO + i;
    if (co >= P.Cout) continue;
    float add = P.bias ? P.bias[co] : 0.0f;
    if (P.ebias) add += P.ebias[(size_t)b * P.Cout + co];
    const size_t o = ((size_t)b * P.Cout + co) * HWo + pstrip;
    float4 v = make_float4(acc[i][0] + add, acc[i][1] + add, acc[i][2] + add, acc[i][3] + add);
    if (P.res) {
      const float4 r4 = *reinterpret_cast<const float4*>(P.res + o);
      v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
    }
    *reinterpret_cast<float4*>(P.out + o) = v;
  }
}

// ---- degenerate shapes of the same conv: one input channel (the U-Net's first conv) / one output channel (its last) ------
// Cin = 1, 3x3, stride 1: thread = one output position, all Cout channels (weights + bias in shared memory); stores are
// coalesced per channel.  HBM-bound on the (B, Cout, H, W) output.
__global__ void __launch_bounds__(256) conv2d_cin1_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out, int Cout,
                                                          int H, int Wd, long long npos) {
  extern __shared__ float swb[];  // [Cout][9] + [Cout]
  for (int e = threadIdx.x; e < Cout * 9; e += 256) swb[e] = W[e];
  for (int e = threadIdx.x; e < Cout; e += 256) swb[Cout * 9 + e] = bias ? bias[e] : 0.0f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= npos) return;
  const int HW = H * Wd, b = (int)(p / HW), rem = (int)(p % HW), y = rem / Wd, xx = rem % Wd;
  float in[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int iy = y + ky - 1, ix = xx + kx - 1;
      in[ky * 3 + kx] = (iy >= 0 && iy < H && ix >= 0 && ix < Wd) ? __ldg(x + (size_t)b * HW + iy * Wd + ix) : 0.0f;
    }
  float* o = out + (size_t)b * Cout * HW + rem;
  for (int co = 0; co < Cout; ++co) {
    float a = swb[Cout * 9 + co];
#pragma unroll
    for (int k = 0; k < 9; ++k) a = fmaf(swb[co * 9 + k], in[k], a);
    o[(size_t)co * HW] = a;
  }
}

// Cout = 1, 3x3, stride 1, GroupNorm(+SiLU) prologue: CTA = one sample x a band of rows; the normalised (+SiLU) bands of
// CH input channels with their zero rings are staged in shared memory per barrier (every global load of the stage in
// flight at once: one channel per barrier made this kernel a chain of 32 load latencies, 72 us for 17 MB), then every
// thread accumulates its 9 taps per channel.  HBM-bound on the (B, Cin, H, W) input.
__global__ void __launch_bounds__(256) conv2d_cout1_kernel(const __grid_constant__ Conv2dParams P, int TR, int CH) {
  extern __shared__ float sband[];  // [CH][(TR + 2) * (Wi + 2)]
  const int Hi = P.Hs, Wi = P.Ws, wc = Wi + 2, bandsz = (TR + 2) * wc;
  const int b = blockIdx.y, y0 = blockIdx.x * TR, tid = threadIdx.x;
  const int ry = tid / Wi, xx = tid % Wi;
  const bool active = ry < TR && y0 + ry < Hi;
  const int Cin = P.C1, cpg = P.G > 0 ? Cin / P.G : 1, HW = Hi * Wi;
  float acc = P.bias ? P.bias[0] : 0.0f;
  for (int c0 = 0; c0 < Cin; c0 += CH) {
    const int nch = min(CH, Cin - c0);
    if (c0 > 0) __syncthreads();  // the previous stage has been consumed
    for (int e = tid; e < nch * bandsz; e += 256) {
      const int cl = e / bandsz, r = e - cl * bandsz, c = c0 + cl;
      const int iy = y0 + r / wc - 1, ix = r % wc - 1;
      float v = 0.0f;
      if (iy >= 0 && iy < Hi && ix >= 0 && ix < Wi) {
        v = __ldg(P.x1 + ((size_t)b * Cin + c) * HW + iy * Wi + ix);
        if (P.prologue) {
          const float mean = P.stats[((size_t)b * P.G + c / cpg) * 2], rstd = P.stats[((size_t)b * P.G + c / cpg) * 2 + 1];
          v = fmaf((v - mean) * rstd, P.gamma[c], P.beta[c]);
          if (P.prologue == 2) v = siluf(v);
        }
      }
      sband[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int cl = 0; cl < nch; ++cl) {  // channel order and tap order as before: bit-identical sums
        const float* w = P.W + (c0 + cl) * 9;
        const float* sb = sband + cl * bandsz;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) acc = fmaf(__ldg(w + ky * 3 + kx), sb[(ry + ky) * wc + xx + kx], acc);
      }
    }
  }
  if (active) {
    const size_t o = (size_t)b * HW + (y0 + ry) * Wi + xx;
    if (P.ebias) acc += P.ebias[b];
    if (P.res) acc += P.res[o];
    P.out[o] = acc;
  }
}

// out[b,co] = bias[co] + sum_i W[co,i] silu(emb[b,i])   (ResBlock.emb_layers, model/unet.py:146-152)
__global__ void __launch_bounds__(256) emb_proj_kernel(const float* __restrict__ emb, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ out, int E,
                                                       int Cout, int B) {
  extern __shared__ float se_[];  // silu(emb[b, :])
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < E; i += 256) se_[i] = siluf(emb[(size_t)b * E + i]);
  __syncthreads();
  for (int co = warp; co < Cout; co += 8) {  // one warp per output row: coalesced reads of W[co, :]
    float s = 0.0f;
    for (int i = lane; i < E; i += 32) s = fmaf(__ldg(W + (size_t)co * E + i), se_[i], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(size_t)b * Cout + co] = s + bias[co];
  }
}

// All ResBlocks' embedding projections of one forward in ONE launch: W (Ctot, E) / bias (Ctot) are the blocks' Linear
// layers stacked along the output dimension; segment i = rows [seg[i], seg[i+1]) is written as its own contiguous
// (B, len_i) matrix at out + B * seg[i], which is what the conv epilogue of that block reads as ebias.
struct EmbSegs { int n; int start[65]; };
constexpr int EP_S = 4;  // samples per CTA: every weight row is read once per EP_S samples (sums in emb_proj_kernel's order)
__global__ void __launch_bounds__(256) emb_proj_multi_kernel(const float* __restrict__ emb, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ out, int E,
                                                             int Ctot, int B, const __grid_constant__ EmbSegs S) {
  extern __shared__ float se_[];  // [EP_S][E] silu(emb[b0 + s, :])
  const int b0 = blockIdx.x * EP_S, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < EP_S * E; i += 256) se_[i] = b0 + i / E < B ? siluf(emb[(size_t)b0 * E + i]) : 0.0f;
  __syncthreads();
  for (int co = blockIdx.y * 8 + warp; co < Ctot; co += 8 * gridDim.y) {
    float s[EP_S];
#pragma unroll
    for (int q = 0; q < EP_S; ++q) s[q] = 0.0f;
    for (int i = lane; i < E; i += 32) {
      const float w = __ldg(W + (size_t)co * E + i);
#pragma unroll
      for (int q = 0; q < EP_S; ++q) s[q] = fmaf(w, se_[q * E + i], s[q]);
    }
#pragma unroll
    for (int q = 0; q < EP_S; ++q) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s[q] += __shfl_xor_sync(0xffffffffu, s[q], o);
    }
    if (lane == 0) {
      int sg = 0;
      while (sg + 1 < S.n && co >= S.start[sg + 1]) ++sg;
      const int s0 = S.start[sg], len = S.start[sg + 1] - s0;
#pragma unroll
      for (int q = 0; q < EP_S; ++q)
        if (b0 + q < B) out[(size_t)B * s0 + (size_t)(b0 + q) * len + (co - s0)] = s[q] + bias[co];
    }
  }
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
// qkv (B, 3C, T) -> out (B, C, T).  One CTA per (sample, 32 queries): S = q^T k / sqrt(C) by 64-key chunks with K staged
// in shared memory (2 queries x 4 keys per thread), row softmax in shared memory, then out = V P^T with V staged per
// 64-key chunk (thread = one query x C/8 channels).  S is kept transposed [key][query] (+1 pad) so that the P.V phase
// reads it conflict-free.
constexpr int AT_Q = 32, AT_K = 64;
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int C, int T) {
  extern __shared__ __align__(16) float sm[];
  float* sq = sm;                          // [C][AT_Q]
  float* skv = sq + C * AT_Q;              // [C][AT_K]   K chunk, later V chunk
  float* ss = skv + C * AT_K;              // [T][AT_Q + 1]
  const int b = blockIdx.y, t0 = blockIdx.x * AT_Q, tid = threadIdx.x;
  const float* q = qkv + (size_t)b * 3 * C * T;
  const float* k = q + (size_t)C * T;
  const float* v = k + (size_t)C * T;
  const float scale2 = 1.0f / sqrtf((float)C);  // (1/sqrt(sqrt(C)))^2: both q and k are scaled in the reference
  for (int e = tid; e < C * AT_Q; e += 256) {
    const int c = e / AT_Q, tq = e % AT_Q;
    sq[e] = t0 + tq < T ? q[(size_t)c * T + t0 + tq] : 0.0f;
  }
  const int tq2 = (tid >> 4) * 2, tk4 = (tid & 15) * 4;
  for (int s0 = 0; s0 < T; s0 += AT_K) {
    __syncthreads();
    for (int e = tid; e < C * AT_K; e += 256) {
      const int c = e / AT_K, sk = e % AT_K;
      skv[e] = s0 + sk < T ? k[(size_t)c * T + s0 + sk] : 0.0f;
    }
    __syncthreads();
    float d[2][4] = {};
    for (int c = 0; c < C; ++c) {
      const float2 q2 = *reinterpret_cast<const float2*>(sq + c * AT_Q + tq2);
      const float4 k4 = *reinterpret_cast<const float4*>(skv + c * AT_K + tk4);
      d[0][0] = fmaf(q2.x, k4.x, d[0][0]); d[0][1] = fmaf(q2.x, k4.y, d[0][1]);
      d[0][2] = fmaf(q2.x, k4.z, d[0][2]); d[0][3] = fmaf(q2.x, k4.w, d[0][3]);
      d[1][0] = fmaf(q2.y, k4.x, d[1][0]); d[1][1] = fmaf(q2.y, k4.y, d[1][1]);
      d[1][2] = fmaf(q2.y, k4.z, d[1][2]); d[1][3] = fmaf(q2.y, k4.w, d[1][3]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (s0 + tk4 + j < T) ss[(s0 + tk4 + j) * (AT_Q + 1) + tq2 + i] = d[i][j] * scale2;
  }
  __syncthreads();
  {  // softmax over keys for each query row: 8 threads per row
    const int row = tid >> 3, sub = tid & 7;
    float mx = -INFINITY;
    for (int s = sub; s < T; s += 8) mx = fmaxf(mx, ss[s * (AT_Q + 1) + row]);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.0f;
    for (int s = sub; s < T; s += 8) {
      const float e = expf(ss[s * (AT_Q + 1) + row] - mx);
      ss[s * (AT_Q + 1) + row] = e;
      sum += e;
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    __syncthreads();
    for (int s = sub; s < T; s += 8) ss[s * (AT_Q + 1) + row] *= inv;
  }
  // out[c][t] = sum_s P[t][s] V[c][s]: thread = query (lane) x channel group (warp): C/8 channels each, C <= 128
  const int tq = tid & 31, cgp = tid >> 5, cpw = C / 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.0f;
  for (int s0 = 0; s0 < T; s0 += AT_K) {
    __syncthreads();
    for (int e = tid; e < C * AT_K; e += 256) {
      const int c = e / AT_K, sk = e % AT_K;
      skv[e] = s0 + sk < T ? v[(size_t)c * T + s0 + sk] : 0.0f;
    }
    __syncthreads();
    const int smax = min(AT_K, T - s0);
    for (int sk = 0; sk < smax; ++sk) {
      const float p = ss[(s0 + sk) * (AT_Q + 1) + tq];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i < cpw) acc[i] = fmaf(p, skv[(cgp * cpw + i) * AT_K + sk], acc[i]);
    }
  }
  if (t0 + tq < T) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < cpw) out[((size_t)b * C + cgp * cpw + i) * T + t0 + tq] = acc[i];
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
  if (P.Wo % 4) {
    set_error("msgm_conv2d: output width must be a multiple of 4");
    return MSGM_ERR_UNSUPPORTED;
  }
  if (D->K == 3 && D->stride == 1 && D->up == 1 && P.C2 == 0 && P.C1 == 1 && !D->prologue && !D->ebias && !D->res &&
      D->Cout <= 512) {
    const long long npos = (long long)D->B * P.Ho * P.Wo;
    conv2d_cin1_kernel<<<(unsigned)((npos + 255) / 256), 256, sizeof(float) * 10 * D->Cout, stream>>>(
        P.x1, P.W, P.bias, P.out, D->Cout, P.Hs, P.Ws, npos);
    ctx->launches += 1;
    MSGM_CUDA_TRY(cudaGetLastError());
    return MSGM_OK;
  }
  if (D->K == 3 && D->stride == 1 && D->up == 1 && P.C2 == 0 && D->Cout == 1 && P.Ws <= 256) {
    const int TR = std::max(1, std::min(256 / P.Ws, 8));
    const size_t band = sizeof(float) * (TR + 2) * (P.Ws + 2);
    const int CH = (int)std::max<size_t>(1, std::min<size_t>(32, (size_t)(44 * 1024) / band));  // channels staged per barrier
    conv2d_cout1_kernel<<<dim3((P.Hs + TR - 1) / TR, D->B), 256, CH * band, stream>>>(P, TR, CH);
    ctx->launches += 1;
    MSGM_CUDA_TRY(cudaGetLastError());
    return MSGM_OK;
  }
  const bool wide = D->Cout >= 64;  // 64 output channels per CTA (8 per thread) when there are that many
  dim3 grid((P.Ho * P.Wo + C2_P - 1) / C2_P, (D->Cout + (wide ? 64 : 32) - 1) / (wide ? 64 : 32), D->B);
  if (D->K == 3 && D->stride == 1) {
    if (wide) conv2d_kernel<3, 1, 64, 8><<<grid, 256, 0, stream>>>(P);
    else conv2d_kernel<3, 1, 32, 8><<<grid, 256, 0, stream>>>(P);
  } else if (D->K == 3) {
    if (wide) conv2d_kernel<3, 2, 64, 8><<<grid, 256, 0, stream>>>(P);
    else conv2d_kernel<3, 2, 32, 8><<<grid, 256, 0, stream>>>(P);
  } else {
    if (D->stride != 1) {
      set_error("msgm_conv2d: 1x1 convolutions are built for stride 1");
      return MSGM_ERR_UNSUPPORTED;
    }
    if (wide) conv2d_kernel<1, 1, 64, 32><<<grid, 256, 0, stream>>>(P);
    else conv2d_kernel<1, 1, 32, 32><<<grid, 256, 0, stream>>>(P);
  }
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int emb_proj(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int E, int Cout, int B,
             cudaStream_t stream) {
  emb_proj_kernel<<<B, 256, sizeof(float) * E, stream>>>(emb, W, bias, out, E, Cout, B);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int emb_proj_multi(msgm_ctx* ctx, const float* emb, const float* W, const float* bias, float* out, int E, int Ctot, int B,
                   int nseg, const int* seg_start, cudaStream_t stream) {
  EmbSegs S{};
  S.n = nseg;
  for (int i = 0; i <= nseg; ++i) S.start[i] = seg_start[i];
  emb_proj_multi_kernel<<<dim3((B + EP_S - 1) / EP_S, 16), 256, sizeof(float) * EP_S * E, stream>>>(emb, W, bias, out, E, Ctot, B, S);
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
  const size_t smem = sizeof(float) * ((size_t)C * AT_Q + (size_t)C * AT_K + (size_t)T * (AT_Q + 1));
  if (C % 8 || C > 128 || smem > 200 * 1024) {
    set_error("msgm_attention: built for C in {8..128} multiple of 8 and T up to ~1000");
    return MSGM_ERR_UNSUPPORTED;
  }
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
