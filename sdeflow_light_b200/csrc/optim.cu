// Adam update of the score net's parameters (torch.optim.Adam as the reference driver uses it, MSGM_higherDim.py:792:
// lr only, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) as ONE launch over the trainer's flat gradient buffer.
//
// The gradients of all parameter tensors live in one flat fp32 buffer (what the fused SSM backward writes and what the
// NCCL all-reduce sums); the first and second moments are flat buffers of the same length.  The parameters themselves stay
// the torch tensors of the nn.Module (separate allocations): a small device table maps flat ranges to their addresses.
// The learning rate and the update counter are device scalars so that the launch can be replayed inside a CUDA graph; the
// last CTA to finish bumps the counter.  `grad_scale` folds the 1/world of a summed all-reduce into the update.
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

struct AdamSeg {
  float* param;
  long long begin;  // first flat index of this tensor
};

constexpr int ADAM_MAX_SEGS = 1024;

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamSeg* __restrict__ segs, int nseg, long long total,
                                                        const float* __restrict__ grad, float* __restrict__ m,
                                                        float* __restrict__ v, const float* __restrict__ lr_dev,
                                                        long long* __restrict__ step_dev, unsigned int* __restrict__ done,
                                                        float beta1, float beta2, float eps, float grad_scale) {
  extern __shared__ unsigned char adam_smem[];
  AdamSeg* ss = reinterpret_cast<AdamSeg*>(adam_smem);
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) ss[i] = segs[i];
  __syncthreads();
  const long long t = *step_dev + 1;  // number of this update (every CTA reads the old value; the last one out bumps it)
  const float lr = *lr_dev;
  // bias corrections in double like torch's scalar path (1 - beta^t), then fp32 arithmetic per element
  const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += stride) {
    int lo = 0, hi = nseg - 1;  // last segment with begin <= i
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (ss[mid].begin <= i) lo = mid; else hi = mid - 1;
    }
    const float g = grad[i] * grad_scale;
    const float mi = fmaf(beta1, m[i], (1.0f - beta1) * g);        // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], (1.0f - beta2) * g * g);    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    float* p = ss[lo].param + (i - ss[lo].begin);
    *p = *p - step_size * (mi / denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *step_dev = t;
      *done = 0u;
    }
  }
}

int adam_step(msgm_ctx* ctx, const void* seg_table, int nseg, int64_t total, const float* grad, float* m, float* v,
              const float* lr_dev, int64_t* step_dev, float beta1, float beta2, float eps, float grad_scale,
              cudaStream_t stream) {
  if (nseg > ADAM_MAX_SEGS) {
    set_error("msgm_adam_step: at most 1024 parameter tensors");
    return MSGM_ERR_UNSUPPORTED;
  }
  unsigned int* done = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(ctx->ws) + 16);
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 8);
  adam_step_kernel<<<blocks, 256, sizeof(AdamSeg) * nseg, stream>>>(reinterpret_cast<const AdamSeg*>(seg_table), nseg, total,
                                                                    grad, m, v, lr_dev, reinterpret_cast<long long*>(step_dev),
                                                                    done, beta1, beta2, eps, grad_scale);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
