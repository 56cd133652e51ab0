// Prior (latent) sampling and the MMD metric -- the two steps either side of the sampler.
//
// latent_sample_kernel  MSGMsde.latent_sample (SDEs.py:438-493, 520-526): r = quantile(r_T, U) [exp(.) - 1e-6 for the
//                       log map], s = z/|z| with z ~ N(0,I), x0 = r s; SGM: x0 = z (SDEs.py:201-203).  One warp per
//                       particle row, Philox noise keyed by the global particle index (or injected U / Z for parity).
// mmd_sums_kernel       compute_mmd (quantitative_comparison.py:22-47): sums of exp(-|a-b|^2 / d^2) over all pairs of
//                       (x,x), (y,y), (x,y) as a tiled pairwise reduction; the reference materialises an (N,M,d)
//                       broadcast (800 MB at N = M = 1e4, d = 2).
#include <algorithm>

#include <cmath>

#include "msgm_common.cuh"

namespace msgm {

constexpr uint32_t STREAM_Z = 0xFFFF0000u, STREAM_U = 0xFFFF0001u;  // Philox "step" ids outside any sampler step

__global__ void __launch_bounds__(256) latent_sample_kernel(const float* __restrict__ rT_sorted, int n_r, int log_map,
                                                            int msgm, const float* __restrict__ U_in,
                                                            const float* __restrict__ Z_in, float* __restrict__ out, int d,
                                                            long long B, unsigned long long seed, unsigned long long poff) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nblk = (d + 3) / 4;
  for (long long row = warp0; row < B; row += nwarps) {
    const unsigned long long pid = poff + (unsigned long long)row;
    float sq = 0.0f;
    for (int blk = lane; blk < nblk; blk += 32) {
      float z[4];
      if (Z_in) {
#pragma unroll
        for (int c = 0; c < 4; ++c) z[c] = (blk * 4 + c < d) ? Z_in[row * d + blk * 4 + c] : 0.0f;
      } else {
        const float4 n4 = philox_normal4(seed, pid, STREAM_Z, (uint32_t)blk);
        z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (blk * 4 + c < d) sq = fmaf(z[c], z[c], sq);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    float scale = 1.0f;
    if (msgm) {
      float u;
      if (U_in) u = U_in[row];
      else u = u01(philox4x32_10(make_uint4((uint32_t)pid, (uint32_t)(pid >> 32), STREAM_U, 0u),
                                 make_uint2((uint32_t)seed, (uint32_t)(seed >> 32))).x);
      // torch.quantile(r_T, U), interpolation='linear': rank = u (n-1), lerp between the neighbouring order statistics
      const float rank = u * (float)(n_r - 1);
      const int lo = (int)floorf(rank), hi = min(lo + 1, n_r - 1);
      const float w = rank - (float)lo, a = rT_sorted[lo], b = rT_sorted[hi];
      float r = (w < 0.5f) ? a + w * (b - a) : b - (b - a) * (1.0f - w);
      if (log_map) r = expf(r) - 1e-6f;
      scale = r / sqrtf(sq);
    }
    for (int blk = lane; blk < nblk; blk += 32) {
      float z[4];
      if (Z_in) {
#pragma unroll
        for (int c = 0; c < 4; ++c) z[c] = (blk * 4 + c < d) ? Z_in[row * d + blk * 4 + c] : 0.0f;
      } else {
        const float4 n4 = philox_normal4(seed, pid, STREAM_Z, (uint32_t)blk);
        z[0] = n4.x; z[1] = n4.y; z[2] = n4.z; z[3] = n4.w;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (blk * 4 + c < d) out[row * d + blk * 4 + c] = scale * z[c];
    }
  }
}

// sums[0] += sum_{i,j} k(x_i,x_j), sums[1] += sum k(y_i,y_j), sums[2] += sum k(x_i,y_j);  k = exp(-|a-b|^2 / d^2)
__global__ void __launch_bounds__(256) mmd_sums_kernel(const float* __restrict__ x, long long N, const float* __restrict__ y,
                                                       long long M, int d, double* __restrict__ sums) {
  constexpr int T = 64, DC = 32;
  __shared__ float sa[T][DC + 1], sb[T][DC + 1];
  __shared__ double red[8];
  const int which = blockIdx.z;  // 0: xx, 1: yy, 2: xy
  const float* A = which == 1 ? y : x;
  const float* Bm = which == 0 ? x : y;
  const long long na = which == 1 ? M : N, nb = which == 0 ? N : M;
  const long long i0 = (long long)blockIdx.x * T, j0 = (long long)blockIdx.y * T;
  if (i0 >= na || j0 >= nb) return;
  const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;  // thread: rows ti*4..+3, cols tj*4..+3
  float dist[4][4] = {};
  for (int c0 = 0; c0 < d; c0 += DC) {
    for (int e = tid; e < T * DC; e += 256) {
      const int r = e / DC, c = e % DC;
      sa[r][c] = (i0 + r < na && c0 + c < d) ? A[(i0 + r) * d + c0 + c] : 0.0f;
      sb[r][c] = (j0 + r < nb && c0 + c < d) ? Bm[(j0 + r) * d + c0 + c] : 0.0f;
    }
    __syncthreads();
    const int cmax = min(DC, d - c0);
    for (int c = 0; c < cmax; ++c) {
      float av[4], bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { av[q] = sa[ti * 4 + q][c]; bv[q] = sb[tj * 4 + q][c]; }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) { const float df = av[p] - bv[q]; dist[p][q] = fmaf(df, df, dist[p][q]); }
    }
    __syncthreads();
  }
  const float inv = 1.0f / ((float)d * (float)d);  // mean over d, then / d  (quantitative_comparison.py:33)
  double part = 0.0;
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (i0 + ti * 4 + p < na && j0 + tj * 4 + q < nb) part += (double)expf(-dist[p][q] * inv);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int w = 0; w < 8; ++w) tot += red[w];
    atomicAdd(sums + which, tot);
  }
}

int latent_sample(msgm_ctx* ctx, const float* rT_sorted, int n_r, int log_map, int msgm, const float* U, const float* Z,
                  float* out, int d, int64_t B, uint64_t seed, uint64_t poff, cudaStream_t stream) {
  const long long warps = std::min<long long>(B, (long long)ctx->num_sms * 64);
  const int grid = (int)((warps * 32 + 255) / 256);
  latent_sample_kernel<<<grid, 256, 0, stream>>>(rT_sorted, n_r, log_map, msgm, U, Z, out, d, B, seed, poff);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

// ---- 1-D Gaussian kernel density, log pdf at the query points (sklearn KernelDensity.score_samples, exact sum) -------
// out_q = logsumexp_i( -((q - s_i)/h)^2 / 2 ) - log(n h sqrt(2 pi)): one CTA per query, per-thread online logsumexp
// over a strided slice of the samples, then a block-level combine.
__global__ void __launch_bounds__(256) kde_logpdf_kernel(const float* __restrict__ samples, int n, float inv_h,
                                                         float log_norm, const float* __restrict__ queries,
                                                         float* __restrict__ out, int m) {
  __shared__ float sM[8], sS[8];
  for (int q = blockIdx.x; q < m; q += gridDim.x) {
    const float x = queries[q];
    float mx = -INFINITY, sum = 0.0f;
    for (int i = threadIdx.x; i < n; i += 256) {
      const float u = (x - __ldg(samples + i)) * inv_h;
      const float e = -0.5f * u * u;
      if (e > mx) { sum = sum * expf(mx - e) + 1.0f; mx = e; }
      else sum += expf(e - mx);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, mx, o), s2 = __shfl_xor_sync(0xffffffffu, sum, o);
      const float mm = fmaxf(mx, m2);
      sum = (mx == -INFINITY ? 0.0f : sum * expf(mx - mm)) + (m2 == -INFINITY ? 0.0f : s2 * expf(m2 - mm));
      mx = mm;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { sM[threadIdx.x >> 5] = mx; sS[threadIdx.x >> 5] = sum; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float mm = sM[0], ss = sS[0];
      for (int w = 1; w < 8; ++w) {
        const float m2 = sM[w], s2 = sS[w], mn = fmaxf(mm, m2);
        ss = (mm == -INFINITY ? 0.0f : ss * expf(mm - mn)) + (m2 == -INFINITY ? 0.0f : s2 * expf(m2 - mn));
        mm = mn;
      }
      out[q] = mm + logf(ss) - log_norm;
    }
  }
}

int kde_logpdf(msgm_ctx* ctx, const float* samples, int n, float bandwidth, const float* queries, float* out, int m,
               cudaStream_t stream) {
  const float log_norm = (float)(std::log((double)n) + std::log((double)bandwidth) + 0.5 * std::log(2.0 * 3.14159265358979323846));
  kde_logpdf_kernel<<<std::min(m, ctx->num_sms * 8), 256, 0, stream>>>(samples, n, 1.0f / bandwidth, log_norm, queries, out, m);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int mmd_sums(msgm_ctx* ctx, const float* x, int64_t N, const float* y, int64_t M, int d, double* sums, cudaStream_t stream) {
  MSGM_CUDA_TRY(cudaMemsetAsync(sums, 0, 3 * sizeof(double), stream));
  const long long mx = std::max(N, M);
  dim3 grid((unsigned)((mx + 63) / 64), (unsigned)((mx + 63) / 64), 3);
  mmd_sums_kernel<<<grid, 256, 0, stream>>>(x, N, y, M, d, sums);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
