// 1-D convolution kernels for the UNet1D score net (NNUnet1D.py:13-179), fp32 CUDA cores (the reference is fp32).
//
// The reference concatenates the 128-channel time(+scale) embedding, constant along the signal, in front of every conv
// block (NNUnet1D.py:81,90,102,156,162,175): for the first conv of a block that is 128 of its 129..384 input channels.
// A channel that is constant along l contributes  sum_k W[co,ci,k] e[ci]  to every interior output position and loses
// one tap at each zero-padded border, so those channels are folded into a per-(sample, out-channel, tap) table
//     E[b,co,k] = sum_{ci in emb} W[co, C_real + ci, k] * emb[b,ci]
// (emb_fold_kernel) and never materialised: 30-99 % fewer MACs in those layers and no (B,128,L) broadcast / concat.
// The remaining real channels may come from two tensors (decoder: upsampled features + skip), again without a concat.
//
// conv1d_kernel       Conv1d  k in {1,3,4}, stride in {1,2}, zero padding, optional second input, optional folded
//                     embedding, bias, optional exact (erf) GELU epilogue.  NCL layout, fp32.
// convt1d_kernel      ConvTranspose1d k=4, stride=2, padding=1 (NNUnet1D.py:98) with right zero-pad to a given length.
// embed_mlp_kernel    Linear(1,E) -> GELU -> Linear(E,E) of the time / log-radius embeddings (NNUnet1D.py:52-68).
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

constexpr int CT_CO = 32;    // output channels per CTA
constexpr int CT_L = 128;    // output positions per CTA
constexpr int CT_CI = 8;     // input channels per shared-memory stage

struct Conv1dParams {
  const float* x1; int C1;      // (B,C1,Lin)
  const float* x2; int C2;      // (B,C2,Lin) or NULL
  const float* W;               // (Cout, C1+C2+Cemb, K)
  const float* bias;            // (Cout)
  const float* E;               // (B,Cout,K) folded embedding or NULL
  float* out;                   // (B,Cout,Lout)
  int Cw;                       // C1 + C2 + Cemb: channel stride of W
  int Cout, K, stride, pad, Lin, Lout, gelu;
};

__global__ void __launch_bounds__(256) conv1d_kernel(const __grid_constant__ Conv1dParams P) {
  // grid: x = position tile, y = out-channel tile, z = sample
  __shared__ float sx[CT_CI][CT_L * 2 + 8];       // input stage: positions l0*stride - pad .. (+ CT_L*stride + K)
  __shared__ float sw[CT_CI][4][CT_CO + 1];        // weight stage: [ci][k][co]
  const int tid = threadIdx.x;
  const int tl = tid & 31, tc = tid >> 5;          // thread: positions tl + 32 j (j<4), out channels tc*4 + i (i<4)
  const int l0 = blockIdx.x * CT_L, co0 = blockIdx.y * CT_CO, b = blockIdx.z;
  const int span = (CT_L - 1) * P.stride + P.K;    // input positions needed by this tile
  const int in0 = l0 * P.stride - P.pad;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  const int Creal = P.C1 + P.C2;
  for (int c0 = 0; c0 < Creal; c0 += CT_CI) {
    __syncthreads();
    for (int e = tid; e < CT_CI * span; e += 256) {
      const int ci = e / span, p = e % span, c = c0 + ci, li = in0 + p;
      float v = 0.0f;
      if (c < Creal && li >= 0 && li < P.Lin)
        v = c < P.C1 ? P.x1[((size_t)b * P.C1 + c) * P.Lin + li] : P.x2[((size_t)b * P.C2 + (c - P.C1)) * P.Lin + li];
      sx[ci][p] = v;
    }
    for (int e = tid; e < CT_CI * P.K * CT_CO; e += 256) {
      const int co = e % CT_CO, k = (e / CT_CO) % P.K, ci = e / (CT_CO * P.K), c = c0 + ci;
      sw[ci][k][co] = (c < Creal && co0 + co < P.Cout) ? P.W[((size_t)(co0 + co) * P.Cw + c) * P.K + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CT_CI; ++ci) {
      for (int k = 0; k < P.K; ++k) {
        float wv[4], xv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = sw[ci][k][tc * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = sx[ci][(tl + 32 * j) * P.stride + k];
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
    const float bv = P.bias ? P.bias[co] : 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + tl + 32 * j;
      if (l >= P.Lout) continue;
      float v = acc[i][j] + bv;
      if (P.E) {  // folded constant channels: a tap contributes unless it falls into the zero padding
        for (int k = 0; k < P.K; ++k) {
          const int li = l * P.stride - P.pad + k;
          if (li >= 0 && li < P.Lin) v += P.E[((size_t)b * P.Cout + co) * P.K + k];
        }
      }
      P.out[((size_t)b * P.Cout + co) * P.Lout + l] = P.gelu ? gelu_erf(v) : v;
    }
  }
}

// ---- degenerate shapes: ONE real input channel (the U-Net's first conv) / ONE output channel (its final 1x1 conv) -------
// Both are HBM-bound on the (B, 32, L) tensor they write / read; thread = one position, coalesced along the signal.
// First conv: k3, stride 1, padding 1, C1 = 1 (+ folded embedding table E, bias, GELU), Cout <= 128.
__global__ void __launch_bounds__(256) conv1d_cin1_kernel(const __grid_constant__ Conv1dParams P, long long npos) {
  extern __shared__ float swb1[];  // [Cout][3] weights of the real channel + [Cout] bias
  for (int e = threadIdx.x; e < P.Cout * 3; e += 256) swb1[e] = P.W[(size_t)(e / 3) * P.Cw * 3 + e % 3];
  for (int e = threadIdx.x; e < P.Cout; e += 256) swb1[P.Cout * 3 + e] = P.bias ? P.bias[e] : 0.0f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= npos) return;
  const int b = (int)(p / P.Lout), l = (int)(p % P.Lout);
  const float* x = P.x1 + (size_t)b * P.Lin;
  const bool hasl = l > 0, hasr = l + 1 < P.Lin;
  const float xm = hasl ? __ldg(x + l - 1) : 0.0f, x0 = __ldg(x + l), xp = hasr ? __ldg(x + l + 1) : 0.0f;
  float* o = P.out + (size_t)b * P.Cout * P.Lout + l;
  for (int co = 0; co < P.Cout; ++co) {
    float a = swb1[P.Cout * 3 + co];
    a = fmaf(swb1[co * 3], xm, a);
    a = fmaf(swb1[co * 3 + 1], x0, a);
    a = fmaf(swb1[co * 3 + 2], xp, a);
    if (P.E) {  // constant embedding channels: a tap contributes where it reads inside the signal
      const float* e = P.E + ((size_t)b * P.Cout + co) * 3;
      a += __ldg(e + 1);
      if (hasl) a += __ldg(e);
      if (hasr) a += __ldg(e + 2);
    }
    o[(size_t)co * P.Lout] = P.gelu ? gelu_erf(a) : a;
  }
}

// Final conv: k1, stride 1, Cout = 1 over [x1, x2]
__global__ void __launch_bounds__(256) conv1d_cout1_k1_kernel(const __grid_constant__ Conv1dParams P, long long npos) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= npos) return;
  const int b = (int)(p / P.Lout), l = (int)(p % P.Lout);
  float a = P.bias ? P.bias[0] : 0.0f;
  const float* x = P.x1 + (size_t)b * P.C1 * P.Lin + l;
  for (int c = 0; c < P.C1; ++c) a = fmaf(__ldg(P.W + c), __ldg(x + (size_t)c * P.Lin), a);
  if (P.C2) {
    const float* y = P.x2 + (size_t)b * P.C2 * P.Lin + l;
    for (int c = 0; c < P.C2; ++c) a = fmaf(__ldg(P.W + P.C1 + c), __ldg(y + (size_t)c * P.Lin), a);
  }
  if (P.E) a += __ldg(P.E + b);
  P.out[p] = P.gelu ? gelu_erf(a) : a;
}

// E[b,co,k] = sum_ci W[co, Coff + ci, k] * emb[b,ci]
// CTA = 8 samples (their embeddings in shared memory), warp = one output channel at a time: the lanes walk the contiguous
// (ci, k) run W[co, Coff.., :] once (coalesced) and accumulate the 8 x K partial sums, then one shuffle reduction each.
constexpr int EF_S = 8;
__global__ void __launch_bounds__(256) emb_fold_kernel(const float* __restrict__ W, const float* __restrict__ emb,
                                                       float* __restrict__ E, int Cw, int Coff, int Cemb, int Cout, int K,
                                                       int B) {
  extern __shared__ float semb[];  // [EF_S][Cemb]
  const int b0 = blockIdx.x * EF_S, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < EF_S * Cemb; e += 256) semb[e] = b0 + e / Cemb < B ? emb[(size_t)b0 * Cemb + e] : 0.0f;
  __syncthreads();
  const int run = Cemb * K;
  for (int co = blockIdx.y * 8 + warp; co < Cout; co += 8 * gridDim.y) {
    const float* w = W + ((size_t)co * Cw + Coff) * K;
    for (int k = 0; k < K; ++k) {
      float acc[EF_S] = {};
      for (int ci = lane; ci < Cemb; ci += 32) {
        const float wv = __ldg(w + ci * K + k);
#pragma unroll
        for (int sidx = 0; sidx < EF_S; ++sidx) acc[sidx] = fmaf(wv, semb[sidx * Cemb + ci], acc[sidx]);
      }
#pragma unroll
      for (int sidx = 0; sidx < EF_S; ++sidx) {
        float v = acc[sidx];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && b0 + sidx < B) E[((size_t)(b0 + sidx) * Cout + co) * K + k] = v;
      }
    }
  }
  (void)run;
}

// Every embedding table of one forward in ONE launch (grid.z = layer): the tables depend on the embedding vector only, so
// the seven launches of a forward (one per ConvBlock1D) collapse into one that runs ahead of the first conv.
__global__ void __launch_bounds__(256) emb_fold_multi_kernel(const __grid_constant__ msgm_emb_fold_multi_desc D) {
  extern __shared__ float semb[];  // [EF_S][Cemb]
  const int z = blockIdx.z, Cemb = D.Cemb, B = D.B;
  const float* __restrict__ W = D.W[z];
  float* __restrict__ E = D.E[z];
  const int Cw = D.Cw[z], Coff = D.Coff[z], Cout = D.Cout[z], K = D.K[z];
  const int b0 = blockIdx.x * EF_S, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < EF_S * Cemb; e += 256) semb[e] = b0 + e / Cemb < B ? D.emb[(size_t)b0 * Cemb + e] : 0.0f;
  __syncthreads();
  for (int co = blockIdx.y * 8 + warp; co < Cout; co += 8 * gridDim.y) {
    const float* w = W + ((size_t)co * Cw + Coff) * K;
    for (int k = 0; k < K; ++k) {
      float acc[EF_S] = {};
      for (int ci = lane; ci < Cemb; ci += 32) {
        const float wv = __ldg(w + ci * K + k);
#pragma unroll
        for (int sidx = 0; sidx < EF_S; ++sidx) acc[sidx] = fmaf(wv, semb[sidx * Cemb + ci], acc[sidx]);
      }
#pragma unroll
      for (int sidx = 0; sidx < EF_S; ++sidx) {
        float v = acc[sidx];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && b0 + sidx < B) E[((size_t)(b0 + sidx) * Cout + co) * K + k] = v;
      }
    }
  }
}

// ConvTranspose1d(k=4, s=2, p=1): out[b,co,l] = bias[co] + sum_ci sum_{k: (l+1-k) even, 0 <= (l+1-k)/2 < Lin} W[ci,co,k] x[b,ci,(l+1-k)/2]
// Positions l >= 2*Lin (right zero padding up to Lout, NNUnet1D.py:168-169) are written as 0.
__global__ void __launch_bounds__(256) convt1d_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                      const float* __restrict__ bias, float* __restrict__ out, int Cin,
                                                      int Cout, int Lin, int Lout) {
  __shared__ float sx[CT_CI][CT_L / 2 + 4];
  __shared__ float sw[CT_CI][4][CT_CO + 1];
  const int tid = threadIdx.x, tl = tid & 31, tc = tid >> 5;
  const int l0 = blockIdx.x * CT_L, co0 = blockIdx.y * CT_CO, b = blockIdx.z;
  const int m0 = l0 / 2 - 1;  // first input position needed: (l0 + 1 - 3) / 2
  float acc[4][4] = {};
  for (int c0 = 0; c0 < Cin; c0 += CT_CI) {
    __syncthreads();
    for (int e = tid; e < CT_CI * (CT_L / 2 + 2); e += 256) {
      const int ci = e / (CT_L / 2 + 2), p = e % (CT_L / 2 + 2), c = c0 + ci, m = m0 + p;
      sx[ci][p] = (c < Cin && m >= 0 && m < Lin) ? x[((size_t)b * Cin + c) * Lin + m] : 0.0f;
    }
    for (int e = tid; e < CT_CI * 4 * CT_CO; e += 256) {
      const int co = e % CT_CO, k = (e / CT_CO) % 4, ci = e / (CT_CO * 4), c = c0 + ci;
      sw[ci][k][co] = (c < Cin && co0 + co < Cout) ? W[((size_t)c * Cout + co0 + co) * 4 + k] : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CT_CI; ++ci) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = l0 + tl + 32 * j;
        // even l: taps k=1 (m = l/2) and k=3 (m = l/2 - 1); odd l: k=0 (m = (l+1)/2) and k=2 (m = (l-1)/2)
        const int ka = (l & 1) ? 0 : 1, kb = ka + 2;
        const int ma = (l + 1 - ka) / 2 - m0, mb = (l + 1 - kb) / 2 - m0;
        const float xa = sx[ci][ma], xb = sx[ci][mb];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(sw[ci][ka][tc * 4 + i], xa, fmaf(sw[ci][kb][tc * 4 + i], xb, acc[i][j]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + tc * 4 + i;
    if (co >= Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int l = l0 + tl + 32 * j;
      if (l < Lout) out[((size_t)b * Cout + co) * Lout + l] = l < 2 * Lin ? acc[i][j] + bias[co] : 0.0f;
    }
  }
}

// out[b,:] (+)= W2 gelu(W1 t[b] + b1) + b2,  W1 (E,1), W2 (E,E);  one CTA per sample, E <= 256
__global__ void __launch_bounds__(256) embed_mlp_kernel(const float* __restrict__ t, const float* __restrict__ W1,
                                                        const float* __restrict__ b1, const float* __restrict__ W2,
                                                        const float* __restrict__ b2, float* __restrict__ out, int E,
                                                        int accumulate) {
  __shared__ float h[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid < E) h[tid] = gelu_erf(fmaf(W1[tid], t[b], b1[tid]));
  __syncthreads();
  if (tid < E) {
    float s = b2[tid];
    for (int j = 0; j < E; ++j) s = fmaf(W2[(size_t)tid * E + j], h[j], s);
    out[(size_t)b * E + tid] = accumulate ? out[(size_t)b * E + tid] + s : s;
  }
}

// Both embedding MLPs of a forward in one launch: out[b,:] = MLP_a(t[b]) (+ MLP_b(u[b]) when u != NULL), each
// Linear(1,E) -> GELU -> Linear(E,E) (NNUnet1D.py:52-68,132-141).  CTA = 8 samples; W2 is read ONCE per CTA, whole, into
// shared memory with every load in flight at once (embed_mlp_kernel reads it once per sample, a row per thread); for
// E <= 128 the 256 threads split the samples (thread = output feature x sample half).  The sums run in the same order as
// embed_mlp_kernel's, so the result is bit-identical to two of its launches.  E <= 224.
constexpr int EM_S = 8;
__global__ void __launch_bounds__(256) embed_mlp2_kernel(const float* __restrict__ t, const float* __restrict__ W1a,
                                                         const float* __restrict__ b1a, const float* __restrict__ W2a,
                                                         const float* __restrict__ b2a, const float* __restrict__ u,
                                                         const float* __restrict__ W1b, const float* __restrict__ b1b,
                                                         const float* __restrict__ W2b, const float* __restrict__ b2b,
                                                         float* __restrict__ out, int B, int E) {
  extern __shared__ float em_smem[];
  float* h = em_smem;                  // [EM_S][E]
  float* sw = em_smem + EM_S * E;      // [E][E + 1]
  const int b0 = blockIdx.x * EM_S, tid = threadIdx.x;
  const int groups = E <= 128 ? 2 : 1, ns = EM_S / groups;  // sample groups handled by different threads
  const int o = groups == 2 ? (tid & 127) : tid, s0 = groups == 2 ? (tid >> 7) * ns : 0;
  const bool active = o < E;
  float tot[EM_S];
#pragma unroll
  for (int si = 0; si < EM_S; ++si) tot[si] = 0.0f;
  for (int m = 0; m < (u ? 2 : 1); ++m) {
    const float* val = m ? u : t;
    const float* W1 = m ? W1b : W1a;
    const float* b1 = m ? b1b : b1a;
    const float* W2 = m ? W2b : W2a;
    const float* b2 = m ? b2b : b2a;
    __syncthreads();
    for (int e = tid; e < E * E; e += 256) sw[(e / E) * (E + 1) + e % E] = __ldg(W2 + e);
    for (int e = tid; e < EM_S * E; e += 256) {
      const int si = e / E, i = e % E;
      h[e] = b0 + si < B ? gelu_erf(fmaf(W1[i], val[b0 + si], b1[i])) : 0.0f;
    }
    __syncthreads();
    if (active) {
      float acc[EM_S];
#pragma unroll
      for (int si = 0; si < EM_S; ++si) acc[si] = b2[o];
      for (int j = 0; j < E; ++j) {
        const float w = sw[o * (E + 1) + j];
#pragma unroll
        for (int si = 0; si < EM_S; ++si)
          if (si < ns) acc[si] = fmaf(w, h[(s0 + si) * E + j], acc[si]);
      }
#pragma unroll
      for (int si = 0; si < EM_S; ++si) tot[si] = m ? tot[si] + acc[si] : acc[si];
    }
  }
  if (active) {
#pragma unroll
    for (int si = 0; si < EM_S; ++si)
      if (si < ns && b0 + s0 + si < B) out[(size_t)(b0 + s0 + si) * E + o] = tot[si];
  }
}

// x -> x / (|x| + eps) * sqrt(L), lognorm = log(|x| + eps)   (NN.py:64-70 + NNUnet1D.py:139)
__global__ void __launch_bounds__(256) normalize_log_radius_kernel(const float* __restrict__ x, float* __restrict__ xn,
                                                                   float* __restrict__ lognorm, int L) {
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  float sq = 0.0f;
  for (int c = tid; c < L; c += 256) sq = fmaf(x[(size_t)b * L + c], x[(size_t)b * L + c], sq);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  float tot = 0.0f;
  for (int w = 0; w < 8; ++w) tot += red[w];
  const float rn = sqrtf(tot) + 1e-6f, sc = sqrtf((float)L);
  for (int c = tid; c < L; c += 256) xn[(size_t)b * L + c] = x[(size_t)b * L + c] / rn * sc;
  if (tid == 0) lognorm[b] = logf(rn);
}

// ---- host wrappers -----------------------------------------------------------------------------------------------
int conv1d(msgm_ctx* ctx, const msgm_conv1d_desc* D, cudaStream_t stream) {
  Conv1dParams P{};
  P.x1 = D->x1; P.C1 = D->C1; P.x2 = D->x2; P.C2 = D->x2 ? D->C2 : 0;
  P.W = D->W; P.bias = D->bias; P.E = D->E; P.out = D->out;
  P.Cw = D->C1 + P.C2 + D->Cemb;
  P.Cout = D->Cout; P.K = D->K; P.stride = D->stride; P.pad = D->pad; P.Lin = D->Lin; P.Lout = D->Lout; P.gelu = D->gelu;
  const long long npos = (long long)D->B * D->Lout;
  if (D->K == 3 && D->stride == 1 && D->pad == 1 && P.C1 == 1 && P.C2 == 0 && D->Cout <= 128) {
    conv1d_cin1_kernel<<<(unsigned)((npos + 255) / 256), 256, sizeof(float) * 4 * D->Cout, stream>>>(P, npos);
    ctx->launches += 1;
    MSGM_CUDA_TRY(cudaGetLastError());
    return MSGM_OK;
  }
  if (D->K == 1 && D->stride == 1 && D->pad == 0 && D->Cout == 1 && D->Cemb == 0) {
    conv1d_cout1_k1_kernel<<<(unsigned)((npos + 255) / 256), 256, 0, stream>>>(P, npos);
    ctx->launches += 1;
    MSGM_CUDA_TRY(cudaGetLastError());
    return MSGM_OK;
  }
  dim3 grid((D->Lout + CT_L - 1) / CT_L, (D->Cout + CT_CO - 1) / CT_CO, D->B);
  conv1d_kernel<<<grid, 256, 0, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int emb_fold(msgm_ctx* ctx, const float* W, const float* emb, float* E, int Cw, int Coff, int Cemb, int Cout, int K, int B,
             cudaStream_t stream) {
  const int gy = std::max(1, std::min((Cout + 7) / 8, 4));
  emb_fold_kernel<<<dim3((B + EF_S - 1) / EF_S, gy), 256, sizeof(float) * EF_S * Cemb, stream>>>(W, emb, E, Cw, Coff, Cemb, Cout,
                                                                                             K, B);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int emb_fold_multi(msgm_ctx* ctx, const msgm_emb_fold_multi_desc* D, cudaStream_t stream) {
  int cmax = 1;
  for (int i = 0; i < D->n; ++i) cmax = std::max(cmax, D->Cout[i]);
  const int gy = std::max(1, std::min((cmax + 7) / 8, 4));
  emb_fold_multi_kernel<<<dim3((D->B + EF_S - 1) / EF_S, gy, D->n), 256, sizeof(float) * EF_S * D->Cemb, stream>>>(*D);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int embed_mlp2(msgm_ctx* ctx, const float* t, const float* W1a, const float* b1a, const float* W2a, const float* b2a,
               const float* u, const float* W1b, const float* b1b, const float* W2b, const float* b2b, float* out, int B, int E,
               cudaStream_t stream) {
  const size_t smem = sizeof(float) * ((size_t)EM_S * E + (size_t)E * (E + 1));
  static bool attr_set = false;  // up to 8 KB + 257 KB of floats at E = 256: raise the dynamic shared-memory limit once
  if (!attr_set) {
    MSGM_CUDA_TRY(cudaFuncSetAttribute(embed_mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  if (smem > 227 * 1024) {
    set_error("msgm_embed_mlp2: embedding width too large for one shared-memory weight tile (E <= 224)");
    return MSGM_ERR_UNSUPPORTED;
  }
  embed_mlp2_kernel<<<(B + EM_S - 1) / EM_S, 256, smem, stream>>>(t, W1a, b1a, W2a, b2a, u, W1b, b1b, W2b, b2b, out, B, E);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int convt1d(msgm_ctx* ctx, const float* x, const float* W, const float* bias, float* out, int B, int Cin, int Cout, int Lin,
            int Lout, cudaStream_t stream) {
  dim3 grid((Lout + CT_L - 1) / CT_L, (Cout + CT_CO - 1) / CT_CO, B);
  convt1d_kernel<<<grid, 256, 0, stream>>>(x, W, bias, out, Cin, Cout, Lin, Lout);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int embed_mlp(msgm_ctx* ctx, const float* t, const float* W1, const float* b1, const float* W2, const float* b2, float* out,
              int B, int E, int accumulate, cudaStream_t stream) {
  embed_mlp_kernel<<<B, 256, 0, stream>>>(t, W1, b1, W2, b2, out, E, accumulate);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int normalize_log_radius(msgm_ctx* ctx, const float* x, float* xn, float* lognorm, int B, int L, cudaStream_t stream) {
  normalize_log_radius_kernel<<<B, 256, 0, stream>>>(x, xn, lognorm, L);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
