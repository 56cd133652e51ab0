// Building blocks of the hand-written score-matching TRAINING path of the U-Net score nets (NNUnet1D.py:110-179,
// NNUnet.py / model/unet.py:101-250 under PluginReverseSDE.ssm_loss, SDEs.py:616-646).
//
// The loss needs the net output a(y) and its directional derivative adot = (da/dy) v.  Every activation of the training path
// is therefore a PAIR stacked along the batch axis: samples [0,B) hold the primal, samples [B,2B) the tangent (forward
// mode).  Linear layers (convolutions, Linear) act on both halves alike, so the inference kernels (tensor-core convs) run
// them unchanged on the 2B-sample tensor; this file holds what inference does not have:
//   pair_act_fwd / pair_act_bwd   (h; hdot) = (phi(z); phi'(z) zdot) and its backward
//                                 (zbar; zdotbar) = (hbar phi' + hdotbar phi'' zdot; hdotbar phi'), phi = exact GELU or SiLU
//   rows_bias_add                 bias on the primal half only
//   conv_wgrad                    gW[co][ci][ky][kx] = sum_{n,oy,ox} cot[n][co][oy][ox] in[n][ci][oy s + ky - p][ox s + kx - p]
//                                 over all 2B samples (1-D: H = 1), register-tiled fp32 with atomics across row slices
//   channel_sums                  bias gradient: sum of cot over the primal half and all positions
//   tap_sums_1d                   cotangent of the folded embedding table of the 1-D U-Net (NNUnet1D.py:156,162,175): per tap
//                                 the sum of cot over the output positions whose tap lands inside the signal
//   gemm_f32                      small dense products of the embedding MLPs and table folds (C = op(A) op(B) [+ C])
//   premodule_pair                NormalizeLogRadius (NN.py:56-70) and its tangent, x sqrt(d) as the U-Nets apply it
//   sparse_ssm_loss / _cot        loss_b = q . adot + |a|^2/2 (+ beta |v|^2/2, SGM) with q from the cyclic sparse tensor
//                                 (SDEs.py:369-399) and the output cotangent pair (gout a; gout q)
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

// ---- activations -----------------------------------------------------------------------------------------------------------
struct ActD {
  float h, d1, d2;
};
__device__ __forceinline__ ActD act_derivs(float z, int act) {
  ActD r;
  if (act == 0) {  // exact GELU: z Phi(z)
    const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * z * z);
    r.h = z * cdf;
    r.d1 = cdf + z * pdf;
    r.d2 = pdf * (2.0f - z * z);
  } else {  // SiLU
    const float sg = 1.0f / (1.0f + expf(-z));
    r.h = z * sg;
    r.d1 = sg * (1.0f + z * (1.0f - sg));
    r.d2 = sg * (1.0f - sg) * (2.0f + z * (1.0f - 2.0f * sg));
  }
  return r;
}

// z: (2B, n) flattened per sample; first half primal, second half tangent
__global__ void __launch_bounds__(256) pair_act_fwd_kernel(const float* __restrict__ z, float* __restrict__ h, long long half,
                                                           int act) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half; i += stride) {
    const ActD a = act_derivs(z[i], act);
    h[i] = a.h;
    h[half + i] = a.d1 * z[half + i];
  }
}

__global__ void __launch_bounds__(256) pair_act_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gh,
                                                           float* __restrict__ gz, long long half, int act) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half; i += stride) {
    const ActD a = act_derivs(z[i], act);
    const float hb = gh[i], hdb = gh[half + i];
    gz[i] = fmaf(hb, a.d1, hdb * a.d2 * z[half + i]);
    gz[half + i] = hdb * a.d1;
  }
}

// Range scaling around the tensor-core data-gradient convs.  The tcgen05 convs split every fp32 operand into fp16 hi + lo; the
// cotangents of the deep layers are ~1e-7 and smaller, i.e. inside / below the fp16 subnormal range, where the split loses its
// bits.  amax_kernel finds max|x| (non-negative float bits are ordered as integers), pow2_scale_kernel multiplies by the power
// of two that brings max|x| to 2^target (or by its inverse), so the conv sees O(1e4) values and nothing is rounded by the
// scaling itself.
__global__ void __launch_bounds__(256) amax_kernel(const float* __restrict__ x, long long n, unsigned int* __restrict__ out) {
  float m = 0.0f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0f && isfinite(m)) atomicMax(out, __float_as_uint(m));
}

// max|.| of up to three tensors in one launch: tensor i goes to word out[slot[i]] (words zeroed by the caller's memset)
struct AmaxMulti {
  const float* x[3];
  long long n[3];
  int slot[3];
};
__global__ void __launch_bounds__(256) amax_multi_kernel(const __grid_constant__ AmaxMulti A, unsigned int* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll
  for (int t = 0; t < 3; ++t) {
    if (!A.x[t]) continue;
    float m = 0.0f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < A.n[t]; i += stride) m = fmaxf(m, fabsf(A.x[t][i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0f && isfinite(m)) atomicMax(out + A.slot[t], __float_as_uint(m));
  }
}

__device__ __forceinline__ float pow2_factor(const unsigned int* amax_bits, int target_exp, int inverse) {
  const float m = __uint_as_float(*amax_bits);
  if (!(m > 0.0f)) return 1.0f;
  int e;
  frexpf(m, &e);  // m = f 2^e, f in [0.5, 1)
  const int k = max(-120, min(120, target_exp - e));
  return ldexpf(1.0f, inverse ? -k : k);
}

__global__ void __launch_bounds__(256) pow2_scale_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                         const unsigned int* __restrict__ amax_bits, int target_exp, int inverse) {
  const float f = pow2_factor(amax_bits, target_exp, inverse);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) y[i] = x[i] * f;
}

// x: (N, C, P): x[n][c][p] += bias[c] for n < nrows
__global__ void __launch_bounds__(256) rows_bias_add_kernel(float* __restrict__ x, const float* __restrict__ bias, long long nrows,
                                                            int C, long long P) {
  const long long total = nrows * C * P, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) x[i] += bias[(i / P) % C];
}

// out[c] = sum_{n < nrows, p} x[n][c][p]
__global__ void __launch_bounds__(256) channel_sums_kernel(const float* __restrict__ x, float* __restrict__ out, long long nrows,
                                                           int C, long long P) {
  const int c = blockIdx.x;
  float s = 0.0f;
  for (long long n = blockIdx.y; n < nrows; n += gridDim.y)
    for (long long p = threadIdx.x; p < P; p += blockDim.x) s += x[(n * C + c) * P + p];
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out + c, red[0]);
}

// E[n][co][k] = sum over output positions p with 0 <= p*stride + k - pad < Lin of cot[n][co][p]     (1-D)
__global__ void __launch_bounds__(128) tap_sums_1d_kernel(const float* __restrict__ cot, float* __restrict__ E, int K, int stride,
                                                          int pad, int Lin, int Lout) {
  const long long nc = blockIdx.x;  // n * Cout + co
  const float* row = cot + nc * Lout;
  float s = 0.0f;
  for (int p = threadIdx.x; p < Lout; p += blockDim.x) s += row[p];
  __shared__ float red[128];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x < K) {  // total minus the positions whose tap k falls outside [0, Lin)
    const int k = threadIdx.x;
    float t = red[0];
    for (int p = 0; p < Lout && p * stride + k - pad < 0; ++p) t -= row[p];
    for (int p = Lout - 1; p >= 0 && p * stride + k - pad >= Lin; --p) t -= row[p];
    E[nc * K + k] = t;
  }
}

// ---- register-tiled fp32 products ------------------------------------------------------------------------------------------
// Both the conv weight gradient and the small dense products below are C (TM x TN) += A^T B over a K range staged through
// shared memory as sA[k][TM], sB[k][TN].  A thread owns 8 x 8 outputs: rows {4 rg .. 4 rg + 3} and {TM/2 + 4 rg ..}, columns
// likewise, so that one k step is four 16-byte shared-memory loads (conflict-free: the lanes of a quarter warp read
// consecutive 16-byte words or the same word) for 64 FMAs.  Row strides TM + 4 / TN + 4 keep the loads aligned and put
// neighbouring k rows four banks apart.
template <int TM, int TN>
struct Tile {
  static constexpr int NT = (TM / 8) * (TN / 8);  // thread tiles covering the C tile
  static constexpr int SA = TM + 4, SB = TN + 4;
};

// k steps [k_first, k_end) in steps of k_step of the staged tiles
template <int TM, int TN>
__device__ __forceinline__ void tile_fma(const float* __restrict__ sA, const float* __restrict__ sB, int tt, int k_first, int k_end,
                                         int k_step, float (&acc)[8][8]) {
  using T = Tile<TM, TN>;
  const int rg = tt % (TM / 8), cg = tt / (TM / 8);
  for (int k = k_first; k < k_end; k += k_step) {
    const float4 a0 = *reinterpret_cast<const float4*>(sA + k * T::SA + 4 * rg);
    const float4 a1 = *reinterpret_cast<const float4*>(sA + k * T::SA + TM / 2 + 4 * rg);
    const float4 b0 = *reinterpret_cast<const float4*>(sB + k * T::SB + 4 * cg);
    const float4 b1 = *reinterpret_cast<const float4*>(sB + k * T::SB + TN / 2 + 4 * cg);
    const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}
template <int T_>
__device__ __forceinline__ int tile_index(int g, int i) { return (i < 4 ? 0 : T_ / 2 - 4) + 4 * g + i; }  // row / column of output i

// ---- weight gradient of a convolution ------------------------------------------------------------------------------------
struct WgradConvParams {
  const float* cot;   // (N, Cout, Ho, Wo)
  const float* in1;   // (N, C1, Hi, Wi)
  const float* in2;   // (N, C2, Hi, Wi) or NULL: channel concat [in1, in2]
  float* gW;          // (Cout, Cw, KH, KW): channels [coff, coff + C1 + C2) of the weight's input axis are written
  int N, Cout, C1, C2, Cw, coff, KH, KW, stride, pad_h, pad_w, up;  // up = 2: the conv reads the nearest-upsampled input
  int Hi, Wi, Ho, Wo;
  long long pos_per_slice;  // output positions (n, oy, ox) per z-slice
};

// grid: x = TM x TN tile of (co, ci), y = tap, z = slice of the positions.  256 threads = Tile::NT thread tiles x KG groups that
// take every KG-th position of a staged chunk; partial sums meet in gW through atomics (as the slices do).
// Staging: a warp instruction covers 8 consecutive positions of 4 channels (32-byte global segments; banks 4 pp + ch: no
// shared-memory conflict), and a thread keeps its position for the whole chunk, so (n, oy, ox) is decomposed once per chunk.
// The global loads of chunk i + 1 are issued into registers before the FMAs of chunk i (the small tiles are latency-bound).
template <int TM, int TN>
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const __grid_constant__ WgradConvParams P) {
  using T = Tile<TM, TN>;
  // positions per staged chunk: more for the small tiles, whose chunks would otherwise be two k steps between two barriers
  constexpr int KG = 256 / T::NT, KT = TM + TN <= 64 ? 128 : (TM + TN <= 128 ? 64 : 32), NJ = KT / 32;
  const int Cin = P.C1 + P.C2;
  const int tiles_ci = (Cin + TN - 1) / TN;
  const int tm = (blockIdx.x / tiles_ci) * TM, tn = (blockIdx.x % tiles_ci) * TN;
  const int ky = blockIdx.y / P.KW, kx = blockIdx.y % P.KW;
  __shared__ __align__(16) float tile[KT * (T::SA + T::SB)];  // >= 16 KB for every instantiation: reused by the K-group reduction
  float* sC = tile;
  float* sI = tile + KT * T::SA;
  const int tid = threadIdx.x, tt = tid % T::NT, kg = tid / T::NT;
  const int pp = (tid & 7) + 8 * ((tid >> 5) & 3), ch0 = ((tid >> 3) & 3) + 4 * (tid >> 7);  // staging role: position, first channel
  float acc[8][8] = {};
  const long long npos = (long long)P.N * P.Ho * P.Wo;
  const long long p_begin = (long long)blockIdx.z * P.pos_per_slice, p_end = min(npos, p_begin + P.pos_per_slice);
  const int Hu = P.Hi * P.up, Wu = P.Wi * P.up;  // extent the conv sees
  const long long HWo = (long long)P.Ho * P.Wo, HWi = (long long)P.Hi * P.Wi;
  float rc[NJ][TM / 8], ri[NJ][TN / 8];
  auto fetch = [&](long long p0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const long long p = p0 + pp + 32 * j;
      const bool live = p < p_end;
      long long n = 0;
      int ox = 0, oy = 0;
      if (live) {
        n = p / HWo;
        const int r = (int)(p - n * HWo);
        oy = r / P.Wo;
        ox = r - oy * P.Wo;
      }
      const int iy = oy * P.stride + ky - P.pad_h, ix = ox * P.stride + kx - P.pad_w;
      const bool inside = live && iy >= 0 && iy < Hu && ix >= 0 && ix < Wu;
      const float* cp = P.cot + (n * P.Cout) * HWo + (long long)oy * P.Wo + ox;
      const long long ioff = (long long)(iy / P.up) * P.Wi + ix / P.up;
#pragma unroll
      for (int q = 0; q < TM / 8; ++q) {  // cotangent tile: TM channels x KT positions
        const int ch = ch0 + 8 * q;
        rc[j][q] = (live && tm + ch < P.Cout) ? __ldg(cp + (long long)(tm + ch) * HWo) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < TN / 8; ++q) {  // input tile at this tap: TN channels x KT positions
        const int ci = tn + ch0 + 8 * q;
        float x = 0.0f;
        if (inside && ci < Cin)
          x = ci < P.C1 ? __ldg(P.in1 + (n * P.C1 + ci) * HWi + ioff) : __ldg(P.in2 + (n * P.C2 + ci - P.C1) * HWi + ioff);
        ri[j][q] = x;
      }
    }
  };
  if (p_begin < p_end) fetch(p_begin);
  for (long long p0 = p_begin; p0 < p_end; p0 += KT) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
#pragma unroll
      for (int q = 0; q < TM / 8; ++q) sC[(pp + 32 * j) * T::SA + ch0 + 8 * q] = rc[j][q];
#pragma unroll
      for (int q = 0; q < TN / 8; ++q) sI[(pp + 32 * j) * T::SB + ch0 + 8 * q] = ri[j][q];
    }
    __syncthreads();
    if (p0 + KT < p_end) fetch(p0 + KT);
    tile_fma<TM, TN>(sC, sI, tt, kg, KT, KG, acc);
    __syncthreads();
  }
  if constexpr (KG > 1) {  // sum the K groups' partial tiles in shared memory first: KG times fewer atomics on the same addresses
    static_assert(KT * (T::SA + T::SB) >= 32 * 128, "reduction buffer");
#pragma unroll
    for (int half = KG / 2; half >= 1; half >>= 1) {
#pragma unroll
      for (int ih = 0; ih < 2; ++ih) {  // rows 0..3 and 4..7 of the thread tiles in turn (16 KB buffer)
        if (kg >= half && kg < 2 * half) {
          const int slot = (kg - half) * T::NT + tt;  // < 128
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) tile[(i * 8 + j) * 128 + slot] = acc[4 * ih + i][j];
        }
        __syncthreads();
        if (kg < half) {
          const int slot = kg * T::NT + tt;
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[4 * ih + i][j] += tile[(i * 8 + j) * 128 + slot];
        }
        __syncthreads();
      }
    }
    if (kg != 0) return;
  }
  const int rg = tt % (TM / 8), cg = tt / (TM / 8);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = tm + tile_index<TM>(rg, i), ci = tn + tile_index<TN>(cg, j);
      if (co < P.Cout && ci < Cin)
        atomicAdd(P.gW + (((size_t)co * P.Cw + P.coff + ci) * P.KH + ky) * P.KW + kx, acc[i][j]);
    }
}

// ---- small dense products: C (M x N, ldc) = [C +] alpha sum_s op(A_s) op(B_s), op = identity or transpose -----------------------
// A_s is (M x K) [lda] or, transposed, stored (K x M); B_s is (K x N) [ldb] or, transposed, stored (N x K).  One launch carries a
// GROUP of up to MSGM_GEMM_MAX_PROBLEMS problems of one shape, each the sum of up to two products (the attention of the training
// path is 21 small batched products per block; its forward-mode pairs are sums such as Sdot = qdot^T k + q^T kdot), each
// batched over `batch` matrices: blockIdx.z = problem * batch + matrix.  256 threads per TM x TN tile = Tile::NT thread tiles x
// KG groups; group g takes every KG-th k of a staged chunk of 16 and the groups' partial tiles are summed through shared memory
// (tree, log2 KG rounds) -- these are many small matrices, so the threads have to come from the K axis.
struct GemmGroupParams {
  msgm_gemm_problem prob[MSGM_GEMM_MAX_PROBLEMS];
  int nprob, M, N, K, batch;
};

template <int TM, int TN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const __grid_constant__ GemmGroupParams G) {
  using T = Tile<TM, TN>;
  constexpr int KT = 16, NT = T::NT, KG = 256 / NT;
  constexpr int EA = KT * TM / 256, EB = KT * TN / 256;  // staged elements per thread
  __shared__ __align__(16) float sA[KT * T::SA];
  __shared__ __align__(16) float sB[KT * T::SB];
  __shared__ float red[KG > 1 ? 64 * 128 : 1];
  const msgm_gemm_problem& Q = G.prob[blockIdx.z / G.batch];
  const long long bz = blockIdx.z % G.batch;
  const int M = G.M, N = G.N, K = G.K;
  const int tm = blockIdx.y * TM, tn = blockIdx.x * TN, tid = threadIdx.x, tt = tid % NT, kg = tid / NT;
  float acc[8][8] = {};
  for (int sgm = 0; sgm < Q.nseg; ++sgm) {
    const float* __restrict__ A = Q.A[sgm] + bz * Q.stride_a[sgm];
    const float* __restrict__ B = Q.B[sgm] + bz * Q.stride_b[sgm];
    const int ta = Q.trans_a[sgm], tb = Q.trans_b[sgm], lda = Q.lda[sgm], ldb = Q.ldb[sgm];
    float ra[EA], rb[EB];
    auto fetch = [&](int k0) {  // lanes run along the operand's contiguous axis
#pragma unroll
      for (int q = 0; q < EA; ++q) {
        const int e = tid + 256 * q;
        const int kk = ta ? e / TM : e % KT, mm = ta ? e % TM : e / KT;
        const int k = k0 + kk, m = tm + mm;
        ra[q] = (k < K && m < M) ? __ldg(ta ? A + (size_t)k * lda + m : A + (size_t)m * lda + k) : 0.0f;
      }
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const int e = tid + 256 * q;
        const int kk = tb ? e % KT : e / TN, nn = tb ? e / KT : e % TN;
        const int k = k0 + kk, n = tn + nn;
        rb[q] = (k < K && n < N) ? __ldg(tb ? B + (size_t)n * ldb + k : B + (size_t)k * ldb + n) : 0.0f;
      }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += KT) {
#pragma unroll
      for (int q = 0; q < EA; ++q) {
        const int e = tid + 256 * q;
        sA[(ta ? e / TM : e % KT) * T::SA + (ta ? e % TM : e / KT)] = ra[q];
      }
#pragma unroll
      for (int q = 0; q < EB; ++q) {
        const int e = tid + 256 * q;
        sB[(tb ? e % KT : e / TN) * T::SB + (tb ? e / KT : e % TN)] = rb[q];
      }
      __syncthreads();
      if (k0 + KT < K) fetch(k0 + KT);
      tile_fma<TM, TN>(sA, sB, tt, kg, KT, KG, acc);
      __syncthreads();
    }
  }
  if constexpr (KG > 1) {  // tree sum over the K groups: the upper half writes, the lower half adds
#pragma unroll
    for (int half = KG / 2; half >= 1; half >>= 1) {
      if (kg >= half && kg < 2 * half) {
        const int slot = (kg - half) * NT + tt;  // < 128
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) red[(i * 8 + j) * 128 + slot] = acc[i][j];
      }
      __syncthreads();
      if (kg < half) {
        const int slot = kg * NT + tt;
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += red[(i * 8 + j) * 128 + slot];
      }
      __syncthreads();
    }
    if (kg != 0) return;
  }
  float* __restrict__ Cm = Q.C + bz * Q.stride_c;
  const int rg = tt % (TM / 8), cg = tt / (TM / 8);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int m = tm + tile_index<TM>(rg, i), n = tn + tile_index<TN>(cg, j);
      if (m < M && n < N) {
        float* c = Cm + (size_t)m * Q.ldc + n;
        *c = Q.accumulate ? fmaf(Q.alpha, acc[i][j], *c) : Q.alpha * acc[i][j];
      }
    }
}

// ---- premodule: NormalizeLogRadius (NN.py:56-70) and its tangent, times `scale` (the U-Nets multiply by sqrt(d)) ------------------
// x, v: (B, d) -> xn: (2B, d), logn: (2B): one warp per sample
__global__ void __launch_bounds__(256) premodule_pair_kernel(const float* __restrict__ x, const float* __restrict__ v,
                                                             float* __restrict__ xn, float* __restrict__ logn, long long B, int d,
                                                             float scale) {
  const int lane = threadIdx.x & 31;
  const long long w0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long b = w0; b < B; b += nw) {
    float r2 = 0.0f, xv = 0.0f;
    for (int c = lane; c < d; c += 32) {
      const float xc = x[b * d + c];
      r2 = fmaf(xc, xc, r2);
      xv = fmaf(xc, v[b * d + c], xv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      xv += __shfl_xor_sync(0xffffffffu, xv, o);
    }
    const float r = sqrtf(r2), rn = r + 1e-6f, rdot = xv / r;
    for (int c = lane; c < d; c += 32) {
      const float xc = x[b * d + c], vc = v[b * d + c];
      xn[b * d + c] = scale * (xc / rn);
      xn[(B + b) * d + c] = scale * (vc / rn - xc * rdot / (rn * rn));
    }
    if (lane == 0) {
      logn[b] = logf(rn);
      logn[B + b] = rdot / rn;
    }
  }
}

// ---- loss and output cotangents for the sparse multiplicative SDE / the additive SDE (state width d up to 4096) ------------------
// a: (2B, d) net output pair.  q_k = c sqrt(beta) (v_k y_{k+1} - v_{k+1} y_k) (cyclic) or sqrt(beta) v_k (SGM).
__device__ __forceinline__ float ssm_q(int kind, float sb, const float* y, const float* v, int k, int d) {
  if (kind == MSGM_SDE_SGM) return sb * v[k];
  const int kn = (k + 1 == d) ? 0 : k + 1;
  return SQRT_HALF * sb * (v[k] * y[kn] - v[kn] * y[k]);
}

__global__ void __launch_bounds__(256) sparse_ssm_loss_kernel(const float* __restrict__ a, const float* __restrict__ y,
                                                              const float* __restrict__ v, const float* __restrict__ t,
                                                              float* __restrict__ loss, long long B, int d, int kind, float bmin,
                                                              float bdel) {
  const long long b = blockIdx.x;
  const float bt = beta_of(bmin, bdel, t[b]), sb = sqrtf(bt);
  float s = 0.0f;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float ak = a[b * d + k], adk = a[(B + b) * d + k];
    s = fmaf(ssm_q(kind, sb, y + b * d, v + b * d, k, d), adk, s);
    s = fmaf(0.5f * ak, ak, s);
    if (kind == MSGM_SDE_SGM) s = fmaf(0.5f * bt, v[b * d + k] * v[b * d + k], s);
  }
  __shared__ float red[256];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[b] = red[0];
}

__global__ void __launch_bounds__(256) sparse_ssm_cot_kernel(const float* __restrict__ a, const float* __restrict__ y,
                                                             const float* __restrict__ v, const float* __restrict__ t,
                                                             const float* __restrict__ gout, float* __restrict__ cot, long long B,
                                                             int d, int kind, float bmin, float bdel) {
  const long long b = blockIdx.x;
  const float sb = sqrtf(beta_of(bmin, bdel, t[b])), g = gout[b];
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    cot[b * d + k] = g * a[b * d + k];
    cot[(B + b) * d + k] = g * ssm_q(kind, sb, y + b * d, v + b * d, k, d);
  }
}


// ---- GroupNorm on a pair (model/nn_utils.py:39-46,107-114) ----------------------------------------------------------------------
// Primal: xh = (x - mu) r, y = gamma xh + beta.  Tangent: ydot = gamma r (u - c xh), u = xdot - mean(xdot), c = mean(xh xdot).
// x: (2B, C, HW).  One CTA per (sample, group).  stats: (B, G, 4) = mu, r, mean(xdot), c.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.0f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(256) gn_pair_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* __restrict__ y,
                                                          float* __restrict__ stats, int B, int C, int G, int HW, float eps) {
  __shared__ float red[8];
  const int b = blockIdx.x / G, g = blockIdx.x % G, cg = C / G, n = cg * HW;
  const float* xp = x + ((size_t)b * C + g * cg) * HW;
  const float* xt = x + ((size_t)(B + b) * C + g * cg) * HW;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { s1 += xp[i]; s2 += xt[i]; }
  const float mu = block_sum(s1, red) / n, m1 = block_sum(s2, red) / n;
  float v = 0.f, cc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = xp[i] - mu;
    v = fmaf(d, d, v);
    cc = fmaf(d, xt[i], cc);
  }
  const float var = block_sum(v, red) / n, r = rsqrtf(var + eps);
  const float c = block_sum(cc, red) / n * r;  // mean(xh xdot)
  if (threadIdx.x == 0) {
    float* st = stats + ((size_t)b * G + g) * 4;
    st[0] = mu; st[1] = r; st[2] = m1; st[3] = c;
  }
  float* yp = y + ((size_t)b * C + g * cg) * HW;
  float* yt = y + ((size_t)(B + b) * C + g * cg) * HW;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int ch = g * cg + i / HW;
    const float xh = (xp[i] - mu) * r, ga = gamma[ch];
    yp[i] = fmaf(ga, xh, beta[ch]);
    yt[i] = ga * r * (xt[i] - m1 - c * xh);
  }
}

// gy: (2B, C, HW) cotangents of (y; ydot) -> gx: cotangents of (x; xdot); ggamma / gbeta accumulated with atomics
__global__ void __launch_bounds__(256) gn_pair_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ stats, const float* __restrict__ gy,
                                                          float* __restrict__ gx, float* __restrict__ ggamma,
                                                          float* __restrict__ gbeta, int B, int C, int G, int HW) {
  __shared__ float red[8];
  const int b = blockIdx.x / G, g = blockIdx.x % G, cg = C / G, n = cg * HW;
  const size_t op = ((size_t)b * C + g * cg) * HW, ot = ((size_t)(B + b) * C + g * cg) * HW;
  const float* st = stats + ((size_t)b * G + g) * 4;
  const float mu = st[0], r = st[1], m1 = st[2], c = st[3];
  float sa = 0.f, sax = 0.f, sb = 0.f, sbx = 0.f, sbu = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ga = gamma[g * cg + i / HW];
    const float xh = (x[op + i] - mu) * r, u = x[ot + i] - m1;
    const float a = ga * gy[op + i], bb = ga * gy[ot + i];
    sa += a; sax = fmaf(a, xh, sax); sb += bb; sbx = fmaf(bb, xh, sbx); sbu = fmaf(bb, u, sbu);
  }
  const float ma = block_sum(sa, red) / n, max_ = block_sum(sax, red) / n, mb = block_sum(sb, red) / n;
  const float qm = block_sum(sbx, red) / n, p = block_sum(sbu, red) / n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float ga = gamma[g * cg + i / HW];
    const float xh = (x[op + i] - mu) * r, u = x[ot + i] - m1;
    const float a = ga * gy[op + i], bb = ga * gy[ot + i];
    gx[ot + i] = r * (bb - mb - xh * qm);
    gx[op + i] = r * (a - ma - xh * max_) - r * r * (xh * (p - 3.0f * c * qm) + qm * u + c * (bb - mb));
  }
  // per-channel parameter gradients: gbeta_c = sum ybar, ggamma_c = sum (ybar xh + ydotbar xhdot), xhdot = r (u - c xh)
  for (int ch = 0; ch < cg; ++ch) {
    float sg = 0.f, sbeta = 0.f;
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
      const int e = ch * HW + i;
      const float xh = (x[op + e] - mu) * r, u = x[ot + e] - m1;
      sg = fmaf(gy[op + e], xh, sg);
      sg = fmaf(gy[ot + e], r * (u - c * xh), sg);
      sbeta += gy[op + e];
    }
    const float tg = block_sum(sg, red), tb = block_sum(sbeta, red);
    if (threadIdx.x == 0) {
      atomicAdd(ggamma + g * cg + ch, tg);
      atomicAdd(gbeta + g * cg + ch, tb);
    }
  }
}

// ---- softmax on a pair of logit matrices (QKVAttention, model/unet.py:236-250) ------------------------------------------------------
// rows: (nrows, T).  forward: P = softmax(S), Pdot = P (Sdot - sum_l P_l Sdot_l).  One warp per row.
__global__ void __launch_bounds__(256) softmax_pair_fwd_kernel(const float* __restrict__ S, const float* __restrict__ Sd,
                                                               float* __restrict__ Pm, float* __restrict__ Pd, long long nrows,
                                                               int T) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const float* s = S + row * T;
  const float* sd = Sd + row * T;
  float m = -INFINITY;
  for (int j = lane; j < T; j += 32) m = fmaxf(m, s[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float z = 0.f, d = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(s[j] - m);
    z += e;
    d = fmaf(e, sd[j], d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { z += __shfl_xor_sync(0xffffffffu, z, o); d += __shfl_xor_sync(0xffffffffu, d, o); }
  const float iz = 1.0f / z;
  d *= iz;
  for (int j = lane; j < T; j += 32) {
    const float pj = expf(s[j] - m) * iz;
    Pm[row * T + j] = pj;
    Pd[row * T + j] = pj * (sd[j] - d);
  }
}

// backward: given P, Sdot and the direct cotangents A = dL/dP, Pdb = dL/dPdot:
//   d = sum P Sdot, e = sum Pdb P, Pbar = A + Pdb (Sdot - d) - e Sdot, Sdotbar = P (Pdb - e), Sbar = P (Pbar - sum P Pbar)
__global__ void __launch_bounds__(256) softmax_pair_bwd_kernel(const float* __restrict__ Pm, const float* __restrict__ Sd,
                                                               const float* __restrict__ A, const float* __restrict__ Pdb,
                                                               float* __restrict__ Sb, float* __restrict__ Sdb, long long nrows,
                                                               int T) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const float* p = Pm + row * T;
  const float* sd = Sd + row * T;
  const float* a = A + row * T;
  const float* pdb = Pdb + row * T;
  float d = 0.f, e = 0.f;
  for (int j = lane; j < T; j += 32) { d = fmaf(p[j], sd[j], d); e = fmaf(pdb[j], p[j], e); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { d += __shfl_xor_sync(0xffffffffu, d, o); e += __shfl_xor_sync(0xffffffffu, e, o); }
  float w = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float pb = a[j] + pdb[j] * (sd[j] - d) - e * sd[j];
    w = fmaf(p[j], pb, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  for (int j = lane; j < T; j += 32) {
    const float pb = a[j] + pdb[j] * (sd[j] - d) - e * sd[j];
    Sb[row * T + j] = p[j] * (pb - w);
    Sdb[row * T + j] = p[j] * (pdb[j] - e);
  }
}

// ---- sinusoidal embedding of a pair of scalars (model/nn_utils.py:130-148): [cos(x w_k), sin(x w_k)] and its tangent --------------
__global__ void sincos_pair_kernel(const float* __restrict__ val, float* __restrict__ emb, int B, int dim) {
  const int half = dim / 2;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < B * half; e += gridDim.x * blockDim.x) {
    const int b = e / half, k = e % half;
    const float w = expf(-logf(10000.0f) * (float)k / (float)half);
    const float ang = val[b] * w, xd = val[B + b] * w;
    float sn, cs;
    sincosf(ang, &sn, &cs);
    emb[(size_t)b * dim + k] = cs;
    emb[(size_t)b * dim + half + k] = sn;
    emb[(size_t)(B + b) * dim + k] = -sn * xd;
    emb[(size_t)(B + b) * dim + half + k] = cs * xd;
  }
}

// ---- x2 resampling adjoints ------------------------------------------------------------------------------------------------------
// mode 0: out (NC, 2H, 2W): x at the even positions, zeros elsewhere (adjoint of taking every second position: stride-2 conv)
// mode 1: out (NC, H, W): sums of the 2x2 blocks of x (NC, 2H, 2W) (adjoint of nearest-neighbour upsampling)
__global__ void __launch_bounds__(256) resample2_kernel(const float* __restrict__ x, float* __restrict__ out, long long NC, int H,
                                                        int W, int mode) {
  const long long total = mode == 0 ? NC * 4 * H * W : NC * H * W, stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    if (mode == 0) {
      const int ox = (int)(i % (2 * W)), oy = (int)((i / (2 * W)) % (2 * H));
      const long long nc = i / (4LL * H * W);
      out[i] = ((ox | oy) & 1) ? 0.0f : x[(nc * H + (oy >> 1)) * W + (ox >> 1)];
    } else {
      const int ox = (int)(i % W), oy = (int)((i / W) % H);
      const long long nc = i / ((long long)H * W);
      const float* s = x + (nc * 2 * H + 2 * oy) * 2 * W + 2 * ox;
      out[i] = s[0] + s[1] + s[2 * W] + s[2 * W + 1];
    }
  }
}

// tile choice: the largest of 128 x 128, 64 x 64, 32 x 32 that still gives every SM two CTAs (or 32 x 32)
static void launch_gemm_group(msgm_ctx* ctx, const GemmGroupParams& G, cudaStream_t st) {
  const int M = G.M, N = G.N;
  const long long nz = (long long)G.nprob * G.batch;
  auto ctas = [&](int t) { return (long long)((M + t - 1) / t) * ((N + t - 1) / t) * nz; };
  const long long want = 2LL * ctx->num_sms;
  if (M > 64 && N > 64 && ctas(128) >= want) {
    gemm_f32_kernel<128, 128><<<dim3((N + 127) / 128, (M + 127) / 128, (unsigned)nz), 256, 0, st>>>(G);
  } else if (M > 32 && N > 32 && ctas(64) >= want) {
    gemm_f32_kernel<64, 64><<<dim3((N + 63) / 64, (M + 63) / 64, (unsigned)nz), 256, 0, st>>>(G);
  } else {
    gemm_f32_kernel<32, 32><<<dim3((N + 31) / 32, (M + 31) / 32, (unsigned)nz), 256, 0, st>>>(G);
  }
}

static void launch_gemm_f32(msgm_ctx* ctx, const float* A, const float* B, float* Cm, int M, int N, int K, int lda, int ldb, int ldc,
                            long long sa, long long sb, long long sc, int batch, int ta, int tb, float alpha, int accumulate,
                            cudaStream_t st) {
  GemmGroupParams G{};
  G.nprob = 1; G.M = M; G.N = N; G.K = K; G.batch = batch;
  msgm_gemm_problem& Q = G.prob[0];
  Q.A[0] = A; Q.B[0] = B; Q.C = Cm; Q.nseg = 1;
  Q.lda[0] = lda; Q.ldb[0] = ldb; Q.ldc = ldc; Q.trans_a[0] = ta; Q.trans_b[0] = tb;
  Q.stride_a[0] = sa; Q.stride_b[0] = sb; Q.stride_c = sc; Q.alpha = alpha; Q.accumulate = accumulate;
  launch_gemm_group(ctx, G, st);
}

}  // namespace msgm

using namespace msgm;

static int ut_invalid(const char* m) {
  set_error(m);
  return MSGM_ERR_INVALID;
}
static int ut_grid(const msgm_ctx* ctx, long long n) { return (int)std::min<long long>((n + 255) / 256, (long long)ctx->num_sms * 8); }

extern "C" {

int msgm_pair_act(msgm_ctx* ctx, const float* z, const float* grad_h_or_null, float* out, int64_t half_elems, int32_t act,
                  void* stream) {
  if (!ctx || !z || !out || half_elems < 0 || (act != 0 && act != 1)) return ut_invalid("msgm_pair_act: bad argument");
  if (half_elems == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  if (grad_h_or_null)
    pair_act_bwd_kernel<<<ut_grid(ctx, half_elems), 256, 0, (cudaStream_t)stream>>>(z, grad_h_or_null, out, half_elems, act);
  else
    pair_act_fwd_kernel<<<ut_grid(ctx, half_elems), 256, 0, (cudaStream_t)stream>>>(z, out, half_elems, act);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_amax(msgm_ctx* ctx, const float* x, int64_t n, float* amax_out, void* stream) {
  if (!ctx || !x || !amax_out || n < 1) return ut_invalid("msgm_amax: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  MSGM_CUDA_TRY(cudaMemsetAsync(amax_out, 0, sizeof(float), (cudaStream_t)stream));
  amax_kernel<<<ut_grid(ctx, n), 256, 0, (cudaStream_t)stream>>>(x, n, reinterpret_cast<unsigned int*>(amax_out));
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_amax2(msgm_ctx* ctx, const float* cot, int64_t n_cot, const float* x1, int64_t n1, const float* x2_or_null, int64_t n2,
               float* amax_out2, void* stream) {
  if (!ctx || !cot || !x1 || !amax_out2 || n_cot < 1 || n1 < 1 || (x2_or_null && n2 < 1)) return ut_invalid("msgm_amax2: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  MSGM_CUDA_TRY(cudaMemsetAsync(amax_out2, 0, 2 * sizeof(float), (cudaStream_t)stream));
  AmaxMulti A{};
  A.x[0] = cot; A.n[0] = n_cot; A.slot[0] = 0;
  A.x[1] = x1; A.n[1] = n1; A.slot[1] = 1;
  A.x[2] = x2_or_null; A.n[2] = x2_or_null ? n2 : 0; A.slot[2] = 1;
  amax_multi_kernel<<<ut_grid(ctx, std::max<long long>(n_cot, n1)), 256, 0, (cudaStream_t)stream>>>(
      A, reinterpret_cast<unsigned int*>(amax_out2));
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_pow2_scale(msgm_ctx* ctx, const float* x, float* y, int64_t n, const float* amax_dev, int32_t target_exp,
                    int32_t inverse, void* stream) {
  if (!ctx || !x || !y || !amax_dev || n < 1) return ut_invalid("msgm_pow2_scale: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  pow2_scale_kernel<<<ut_grid(ctx, n), 256, 0, (cudaStream_t)stream>>>(x, y, n, reinterpret_cast<const unsigned int*>(amax_dev),
                                                                      target_exp, inverse);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_rows_bias_add(msgm_ctx* ctx, float* x, const float* bias, int64_t nrows, int32_t C, int64_t P, void* stream) {
  if (!ctx || !x || !bias || nrows < 0 || C < 1 || P < 1) return ut_invalid("msgm_rows_bias_add: bad argument");
  if (nrows == 0) return MSGM_OK;
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  rows_bias_add_kernel<<<ut_grid(ctx, nrows * C * P), 256, 0, (cudaStream_t)stream>>>(x, bias, nrows, C, P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_channel_sums(msgm_ctx* ctx, const float* x, float* out_accumulate, int64_t nrows, int32_t C, int64_t P, void* stream) {
  if (!ctx || !x || !out_accumulate || nrows < 1 || C < 1 || P < 1) return ut_invalid("msgm_channel_sums: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  const int gy = (int)std::min<long long>(nrows, std::max(1, ctx->num_sms * 4 / C));
  channel_sums_kernel<<<dim3(C, gy), 256, 0, (cudaStream_t)stream>>>(x, out_accumulate, nrows, C, P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_tap_sums_1d(msgm_ctx* ctx, const float* cot, float* E, int64_t N, int32_t Cout, int32_t K, int32_t stride, int32_t pad,
                     int32_t Lin, int32_t Lout, void* stream) {
  if (!ctx || !cot || !E || N < 1 || Cout < 1 || K < 1 || K > 128 || Lout < 1) return ut_invalid("msgm_tap_sums_1d: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  tap_sums_1d_kernel<<<(unsigned)(N * Cout), 128, 0, (cudaStream_t)stream>>>(cot, E, K, stride, pad, Lin, Lout);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_conv_wgrad(msgm_ctx* ctx, const float* cot, const float* in1, const float* in2, float* gW_accumulate, int32_t N,
                    int32_t Cout, int32_t C1, int32_t C2, int32_t Cw, int32_t coff, int32_t KH, int32_t KW, int32_t stride,
                    int32_t pad, int32_t up, int32_t Hi, int32_t Wi, int32_t Ho, int32_t Wo, void* stream) {
  if (!ctx || !cot || !in1 || !gW_accumulate || N < 1 || Cout < 1 || C1 < 1 || C2 < 0 || (C2 > 0 && !in2) || coff < 0 ||
      coff + C1 + C2 > Cw || KH < 1 || KW < 1 || KH * KW > 64 || stride < 1 || (up != 1 && up != 2))
    return ut_invalid("msgm_conv_wgrad: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  // 1-D convolutions come in as Hi = Ho = KH = 1: no padding along that axis
  WgradConvParams P{cot, in1, in2, gW_accumulate, N, Cout, C1, C2, Cw, coff, KH, KW, stride, KH == 1 ? 0 : pad, pad, up, Hi, Wi, Ho, Wo, 0};
  const long long npos = (long long)N * Ho * Wo;
  const int Cin = C1 + C2;
  const int TMs = Cout > 64 ? 128 : (Cout > 32 ? 64 : 32), TNs = Cin > 64 ? 128 : (Cin > 32 ? 64 : 32);
  const int tiles = ((Cout + TMs - 1) / TMs) * ((Cin + TNs - 1) / TNs);
  // position slices: about three CTAs per SM in total, at least 256 positions each
  long long slices = std::max<long long>(1, std::min<long long>((npos + 255) / 256,
                                                                (long long)ctx->num_sms * 3 / std::max(1, tiles * KH * KW) + 1));
  slices = std::min<long long>(slices, 65535);
  P.pos_per_slice = ((npos + slices - 1) / slices + 127) / 128 * 128;
  const int nslice = (int)((npos + P.pos_per_slice - 1) / P.pos_per_slice);
  const dim3 grid(tiles, KH * KW, nslice);
  cudaStream_t st = (cudaStream_t)stream;
#define MSGM_WGRAD(TM_, TN_) if (TMs == TM_ && TNs == TN_) conv_wgrad_kernel<TM_, TN_><<<grid, 256, 0, st>>>(P)
  MSGM_WGRAD(128, 128); MSGM_WGRAD(128, 64); MSGM_WGRAD(128, 32);
  MSGM_WGRAD(64, 128); MSGM_WGRAD(64, 64); MSGM_WGRAD(64, 32);
  MSGM_WGRAD(32, 128); MSGM_WGRAD(32, 64); MSGM_WGRAD(32, 32);
#undef MSGM_WGRAD
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_gemm_f32(msgm_ctx* ctx, const float* A, const float* B, float* Cm, int32_t M, int32_t N, int32_t K, int32_t lda,
                  int32_t ldb, int32_t ldc, int32_t trans_a, int32_t trans_b, int32_t accumulate, void* stream) {
  if (!ctx || !A || !B || !Cm || M < 1 || N < 1 || K < 1) return ut_invalid("msgm_gemm_f32: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  launch_gemm_f32(ctx, A, B, Cm, M, N, K, lda, ldb, ldc, 0, 0, 0, 1, trans_a, trans_b, 1.0f, accumulate, (cudaStream_t)stream);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_premodule_pair(msgm_ctx* ctx, const float* x, const float* v, float* xn_pair, float* logn_pair, int64_t B, int32_t d,
                        float scale, void* stream) {
  if (!ctx || !x || !v || !xn_pair || !logn_pair || B < 1 || d < 1) return ut_invalid("msgm_premodule_pair: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  premodule_pair_kernel<<<(int)std::min<long long>((B + 7) / 8, (long long)ctx->num_sms * 8), 256, 0, (cudaStream_t)stream>>>(
      x, v, xn_pair, logn_pair, B, d, scale);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_sparse_ssm_loss(msgm_ctx* ctx, const msgm_sde_desc* sde, const float* a_pair, const float* y, const float* v,
                         const float* t, const float* gout_or_null, float* out, int64_t B, void* stream) {
  if (!ctx || !sde || !a_pair || !y || !v || !t || !out || B < 1 || sde->dim < 1)
    return ut_invalid("msgm_sparse_ssm_loss: bad argument");
  if (sde->kind != MSGM_SDE_SGM && sde->kind != MSGM_SDE_MSGM_SPARSE) {
    set_error("msgm_sparse_ssm_loss: built for the additive SDE and the sparse multiplicative SDE");
    return MSGM_ERR_UNSUPPORTED;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  if (gout_or_null)
    sparse_ssm_cot_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(a_pair, y, v, t, gout_or_null, out, B, sde->dim, sde->kind,
                                                                         sde->beta_min, sde->beta_delta);
  else
    sparse_ssm_loss_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(a_pair, y, v, t, out, B, sde->dim, sde->kind, sde->beta_min,
                                                                          sde->beta_delta);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_bgemm_f32(msgm_ctx* ctx, const float* A, const float* B, float* Cm, int32_t M, int32_t N, int32_t K, int32_t lda,
                   int32_t ldb, int32_t ldc, int64_t stride_a, int64_t stride_b, int64_t stride_c, int32_t batch, int32_t trans_a,
                   int32_t trans_b, float alpha, int32_t accumulate, void* stream) {
  if (!ctx || !A || !B || !Cm || M < 1 || N < 1 || K < 1 || batch < 1 || batch > 65535) return ut_invalid("msgm_bgemm_f32: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  launch_gemm_f32(ctx, A, B, Cm, M, N, K, lda, ldb, ldc, stride_a, stride_b, stride_c, batch, trans_a, trans_b, alpha, accumulate,
                  (cudaStream_t)stream);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_gemm_group_f32(msgm_ctx* ctx, const msgm_gemm_problem* problems, int32_t n_problems, int32_t M, int32_t N, int32_t K,
                        int32_t batch, void* stream) {
  if (!ctx || !problems || n_problems < 1 || n_problems > MSGM_GEMM_MAX_PROBLEMS || M < 1 || N < 1 || K < 1 || batch < 1 ||
      (long long)n_problems * batch > 65535)
    return ut_invalid("msgm_gemm_group_f32: bad argument");
  GemmGroupParams G{};
  G.nprob = n_problems; G.M = M; G.N = N; G.K = K; G.batch = batch;
  for (int i = 0; i < n_problems; ++i) {
    const msgm_gemm_problem& Q = problems[i];
    if (Q.nseg < 1 || Q.nseg > 2 || !Q.C || !Q.A[0] || !Q.B[0] || (Q.nseg == 2 && (!Q.A[1] || !Q.B[1])))
      return ut_invalid("msgm_gemm_group_f32: bad problem descriptor");
    G.prob[i] = Q;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  launch_gemm_group(ctx, G, (cudaStream_t)stream);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_gn_pair(msgm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* stats, const float* grad_y_or_null,
                 float* out, float* ggamma, float* gbeta, int32_t B, int32_t C, int32_t G, int32_t HW, void* stream) {
  if (!ctx || !x || !gamma || !stats || !out || B < 1 || C < 1 || G < 1 || C % G || HW < 1) return ut_invalid("msgm_gn_pair: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  if (grad_y_or_null) {
    if (!ggamma || !gbeta) return ut_invalid("msgm_gn_pair: backward needs ggamma / gbeta");
    gn_pair_bwd_kernel<<<B * G, 256, 0, (cudaStream_t)stream>>>(x, gamma, stats, grad_y_or_null, out, ggamma, gbeta, B, C, G, HW);
  } else {
    if (!beta) return ut_invalid("msgm_gn_pair: forward needs beta");
    gn_pair_fwd_kernel<<<B * G, 256, 0, (cudaStream_t)stream>>>(x, gamma, beta, out, stats, B, C, G, HW, 1e-5f);
  }
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_softmax_pair(msgm_ctx* ctx, const float* S_or_P, const float* Sdot, const float* A_or_null, const float* Pdotbar_or_null,
                      float* out1, float* out2, int64_t nrows, int32_t T, void* stream) {
  if (!ctx || !S_or_P || !Sdot || !out1 || !out2 || nrows < 1 || T < 1) return ut_invalid("msgm_softmax_pair: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  const unsigned blocks = (unsigned)((nrows * 32 + 255) / 256);
  if (A_or_null) {
    if (!Pdotbar_or_null) return ut_invalid("msgm_softmax_pair: backward needs both cotangents");
    softmax_pair_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(S_or_P, Sdot, A_or_null, Pdotbar_or_null, out1, out2, nrows, T);
  } else {
    softmax_pair_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(S_or_P, Sdot, out1, out2, nrows, T);
  }
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_sincos_pair(msgm_ctx* ctx, const float* val_pair, float* emb_pair, int32_t B, int32_t dim, void* stream) {
  if (!ctx || !val_pair || !emb_pair || B < 1 || dim < 2 || dim % 2) return ut_invalid("msgm_sincos_pair: bad argument (even dim)");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  sincos_pair_kernel<<<(B * dim / 2 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(val_pair, emb_pair, B, dim);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int msgm_resample2(msgm_ctx* ctx, const float* x, float* out, int64_t NC, int32_t H, int32_t W, int32_t mode, void* stream) {
  if (!ctx || !x || !out || NC < 1 || H < 1 || W < 1 || (mode != 0 && mode != 1)) return ut_invalid("msgm_resample2: bad argument");
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  const long long total = mode == 0 ? NC * 4 * H * W : NC * H * W;
  resample2_kernel<<<ut_grid(ctx, total), 256, 0, (cudaStream_t)stream>>>(x, out, NC, H, W, mode);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // extern "C"
