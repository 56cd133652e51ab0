// Sample-quality metrics of the reference's evaluation step: the numbers behind its survival plot
// and its covariance / energy report, computed on the GPU where the generated particles already live.
//
// row_norm_stats_kernel   |x_b| (optionally after a per-dimension scale, own_plotting.py:646-653,729-736) plus the
//                         smallest positive and the largest norm, which span the shared radius grid (own_plotting.py:616-632).
// survival_hist_kernel /  S(R) = P(|x| > R) on a sorted radius grid: counts[g] = #{b : |x_b| > R_g}, what the reference
// survival_suffix_kernel  gets from sort + searchsorted(side='right') (own_plotting.py:635-640).  One pass over the norms:
//                         binary search of each norm in the grid (shared memory), per-CTA histogram, global atomics, then a
//                         suffix sum.  Integer work, HBM-bound (4 B per particle).
// moments_kernel          column sums and the Gram matrix sum_b x_b x_b^T in double: mean, torch.cov, per-dimension
//                         variance and the energy E|x|^2 of own_plotting.py:339-394 follow on 2 + d + d^2 numbers.
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

// order-preserving map float -> uint32 for atomicMin / atomicMax on non-negative floats (bit pattern is monotone)
__global__ void __launch_bounds__(256) row_norm_stats_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                             float* __restrict__ norms, unsigned int* __restrict__ minmax,
                                                             int d, long long n) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned int lo = 0xFFFFFFFFu, hi = 0u;
  for (long long row = warp0; row < n; row += nwarps) {
    double sq = 0.0;  // double accumulation: the rounded result is the correctly rounded fp32 norm for any d
    for (int c = lane; c < d; c += 32) {
      const float v = x[row * d + c] * (scale ? scale[c] : 1.0f);
      sq += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float r = (float)sqrt(sq);
    if (lane == 0) {
      norms[row] = r;
      const unsigned int bits = __float_as_uint(r);
      if (r > 0.0f) lo = min(lo, bits);
      hi = max(hi, bits);
    }
  }
  if (lane == 0) {
    if (lo != 0xFFFFFFFFu) atomicMin(minmax, lo);
    atomicMax(minmax + 1, hi);
  }
}

constexpr int SURV_MAX_GRID = 4096;

__global__ void __launch_bounds__(256) survival_hist_kernel(const float* __restrict__ norms, long long n,
                                                            const double* __restrict__ grid, int ng,
                                                            unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned char surv_smem[];
  double* sg = reinterpret_cast<double*>(surv_smem);
  unsigned int* sh = reinterpret_cast<unsigned int*>(sg + ng);
  for (int i = threadIdx.x; i < ng; i += blockDim.x) sg[i] = grid[i];
  for (int i = threadIdx.x; i <= ng; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double r = (double)norms[i];  // numpy compares the float32 norm with the float64 grid in double
    int lo = 0, hi = ng;                // b = #{g : R_g < r}
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sg[mid] < r) lo = mid + 1; else hi = mid;
    }
    atomicAdd(sh + lo, 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i <= ng; i += blockDim.x)
    if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// counts[g] = sum_{b > g} hist[b]: one CTA, ng <= 4096
__global__ void survival_suffix_kernel(const unsigned long long* __restrict__ hist, int ng, long long* __restrict__ counts) {
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int g = ng - 1; g >= 0; --g) {
      acc += hist[g + 1];
      counts[g] = (long long)acc;
    }
  }
}

// Gram tile (32 x 32) over a slice of the rows; fp32 products accumulated in double per 32-row chunk
__global__ void __launch_bounds__(256) moments_kernel(const float* __restrict__ x, long long n, int d,
                                                      double* __restrict__ colsum, double* __restrict__ gram) {
  __shared__ float sa[32][33], sb[32][33];
  const int ti = blockIdx.x * 32, tj = blockIdx.y * 32;
  if (tj < ti) return;  // symmetric: upper tiles only, mirrored by the host
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads, 4 outputs each
  const long long rows_per = (n + gridDim.z - 1) / gridDim.z;
  const long long r0 = blockIdx.z * rows_per, r1 = min(n, r0 + rows_per);
  double acc[4] = {0.0, 0.0, 0.0, 0.0}, csum = 0.0;
  for (long long r = r0; r < r1; r += 32) {
    for (int e = threadIdx.x; e < 32 * 32; e += 256) {
      const int rr = e >> 5, cc = e & 31;
      const long long row = r + rr;
      sa[rr][cc] = (row < r1 && ti + cc < d) ? x[row * d + ti + cc] : 0.0f;
      sb[rr][cc] = (row < r1 && tj + cc < d) ? x[row * d + tj + cc] : 0.0f;
    }
    __syncthreads();
    float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float b = sb[k][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) p[q] = fmaf(sa[k][ty + 8 * q], b, p[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] += (double)p[q];
    if (blockIdx.y == blockIdx.x && ty == 0) {
      float s = 0.f;
      for (int k = 0; k < 32; ++k) s += sa[k][tx];
      csum += (double)s;
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int i = ti + ty + 8 * q, j = tj + tx;
    if (i < d && j < d) atomicAdd(gram + (long long)i * d + j, acc[q]);
  }
  if (blockIdx.y == blockIdx.x && ty == 0 && ti + tx < d) atomicAdd(colsum + ti + tx, csum);
}

int row_norm_stats(msgm_ctx* ctx, const float* x, const float* scale, float* norms, float* minpos_max, int d, int64_t n,
                   cudaStream_t stream) {
  unsigned int init[2] = {0xFFFFFFFFu, 0u};
  MSGM_CUDA_TRY(cudaMemcpyAsync(minpos_max, init, sizeof init, cudaMemcpyHostToDevice, stream));
  const int grid = (int)std::min<long long>((n + 7) / 8, (long long)ctx->num_sms * 8);
  row_norm_stats_kernel<<<grid, 256, 0, stream>>>(x, scale, norms, reinterpret_cast<unsigned int*>(minpos_max), d, n);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int survival_counts(msgm_ctx* ctx, const float* norms, int64_t n, const double* grid, int ng, int64_t* counts,
                    void* hist_scratch, cudaStream_t stream) {
  if (ng > SURV_MAX_GRID) {
    set_error("msgm_survival_counts: at most 4096 grid points");
    return MSGM_ERR_UNSUPPORTED;
  }
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(hist_scratch);
  MSGM_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * (ng + 1), stream));
  const int blocks = (int)std::min<long long>((n + 255) / 256, (long long)ctx->num_sms * 4);
  const size_t smem = sizeof(double) * ng + sizeof(unsigned int) * (ng + 1);
  MSGM_CUDA_TRY(cudaFuncSetAttribute(survival_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  survival_hist_kernel<<<blocks, 256, smem, stream>>>(norms, n, grid, ng, hist);
  survival_suffix_kernel<<<1, 32, 0, stream>>>(hist, ng, reinterpret_cast<long long*>(counts));
  ctx->launches += 2;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int moments(msgm_ctx* ctx, const float* x, int64_t n, int d, double* colsum, double* gram, cudaStream_t stream) {
  MSGM_CUDA_TRY(cudaMemsetAsync(colsum, 0, sizeof(double) * d, stream));
  MSGM_CUDA_TRY(cudaMemsetAsync(gram, 0, sizeof(double) * (size_t)d * d, stream));
  const int t = (d + 31) / 32;
  const int tiles = t * (t + 1) / 2;
  int slices = (int)std::min<long long>((n + 1023) / 1024, std::max(1, ctx->num_sms * 4 / tiles));
  slices = std::max(1, std::min(slices, 65535));
  moments_kernel<<<dim3(t, t, slices), 256, 0, stream>>>(x, n, d, colsum, gram);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
