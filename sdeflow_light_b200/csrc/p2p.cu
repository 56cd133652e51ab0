// Gradient all-reduce fused with the Adam update over NVLink peer memory (one node, one process per GPU).
//
// The score-matching step ends with "sum the flat gradient over ranks, then Adam".  At the reference's batch size the
// gradient is 136 KB and an NCCL all-reduce costs ~40 us on 8 GPUs -- a third of the iteration (profiles/bench_8gpu_cfg2_r02).
// Here every rank PUSHES its flat gradient into a receive slot of every rank (plain 16-byte stores through the NVLink
// peer mapping, ~1 MB in total per rank), raises a sequence flag at each peer, and the Adam kernel of each rank waits for the
// world's flags, sums the `world` local slots in rank order (bit-identical on every rank, so the replicas stay in lock step)
// and applies the update: two launches, no collective library call, no read over the link.
//
// Receive slots are double-buffered by the parity of a sequence counter that lives in the handle: a rank can only be one
// call ahead of its slowest peer (it waits for every peer's flag of call s before it finishes call s), so slot parity p is
// rewritten (call s+2) only after every peer has consumed call s.  Waits are bounded; a timeout raises the context's error
// word (msgm_async_error) instead of hanging.
#include <algorithm>
#include <cstring>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

constexpr int P2P_MAX_WORLD = 16;

struct AdamSeg {
  float* param;
  long long begin;
};

struct P2pTable {
  float* recv[P2P_MAX_WORLD];        // peer r's receive buffer: [2 parity][world][total] floats
  long long* flags[P2P_MAX_WORLD];   // peer r's flags:           [2 parity][world]
};

}  // namespace msgm

// (the extern "C" prototypes come from include/msgm_b200.h via msgm_common.cuh)
struct msgm_p2p {
  int world, rank;
  long long total;        // floats per gradient
  void* base;             // this rank's allocation: recv | flags | seq | done
  void* peer_base[msgm::P2P_MAX_WORLD];
  msgm::P2pTable table;   // host copy (passed by value to the kernels)
  long long* seq;         // device: number of completed calls
  unsigned int* done;     // device: CTA counters of the two kernels
};

namespace msgm {

static size_t p2p_recv_bytes(int world, long long total) { return sizeof(float) * 2 * (size_t)world * (size_t)total; }
static size_t p2p_flags_off(int world, long long total) { return (p2p_recv_bytes(world, total) + 255) & ~(size_t)255; }
static size_t p2p_alloc_bytes(int world, long long total) { return p2p_flags_off(world, total) + 8 * 2 * world + 256; }

__global__ void __launch_bounds__(256) p2p_push_kernel(const __grid_constant__ P2pTable T, int world, int rank, long long total,
                                                        const float* __restrict__ grad, const long long* __restrict__ seq,
                                                        unsigned int* __restrict__ done) {
  const long long s = *seq + 1;  // this call's sequence number
  const int par = (int)(s & 1);
  const long long n4 = total >> 2;  // total is padded to a multiple of 4 by the host
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 g = reinterpret_cast<const float4*>(grad)[i];
    for (int p = 0; p < world; ++p)
      reinterpret_cast<float4*>(T.recv[p] + ((size_t)par * world + rank) * total)[i] = g;
  }
  __threadfence_system();  // this CTA's peer stores are visible system-wide before its arrival below
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(done, 1u) == gridDim.x - 1) {  // last CTA: every slice of the gradient has been pushed
      *done = 0u;
      __threadfence_system();
      for (int p = 0; p < world; ++p) *reinterpret_cast<volatile long long*>(T.flags[p] + par * world + rank) = s;
      __threadfence_system();
    }
  }
}

__global__ void __launch_bounds__(256) p2p_adam_kernel(const __grid_constant__ P2pTable T, int world, int rank, long long total,
                                                        const AdamSeg* __restrict__ segs, int nseg, long long nparam,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ lr_dev, long long* __restrict__ step_dev,
                                                        long long* __restrict__ seq, unsigned int* __restrict__ done,
                                                        float beta1, float beta2, float eps, TcFlags err) {
  extern __shared__ unsigned char p2p_smem[];
  AdamSeg* ss = reinterpret_cast<AdamSeg*>(p2p_smem);
  __shared__ int s_ok;
  for (int i = threadIdx.x; i < nseg; i += blockDim.x) ss[i] = segs[i];
  const long long s = *seq + 1;
  const int par = (int)(s & 1);
  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  if (threadIdx.x < world) {  // wait for rank threadIdx.x's gradient of this call (bounded: ~2 s)
    volatile long long* f = reinterpret_cast<volatile long long*>(T.flags[rank] + par * world + threadIdx.x);
    const long long t0 = clock64();
    while (*f != s) {
      if (clock64() - t0 > 4000000000LL) {
        s_ok = 0;
        tc_raise(err, 3);
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
  if (s_ok) {
    const long long t = *step_dev + 1;
    const float lr = *lr_dev;
    const double bc1 = 1.0 - pow((double)beta1, (double)t), bc2 = 1.0 - pow((double)beta2, (double)t);
    const float step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2), inv_world = 1.0f / (float)world;
    const float* mine = T.recv[rank] + (size_t)par * world * total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nparam; i += stride) {
      float g = 0.0f;
      for (int r = 0; r < world; ++r) g += __ldcg(mine + (size_t)r * total + i);  // rank order: identical on every rank
      g *= inv_world;
      int lo = 0, hi = nseg - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (ss[mid].begin <= i) lo = mid; else hi = mid - 1;
      }
      const float mi = fmaf(beta1, m[i], (1.0f - beta1) * g);
      const float vi = fmaf(beta2, v[i], (1.0f - beta2) * g * g);
      m[i] = mi;
      v[i] = vi;
      float* p = ss[lo].param + (i - ss[lo].begin);
      *p = *p - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done + 1, 1u) == gridDim.x - 1) {
      done[1] = 0u;
      *seq = s;
      if (s_ok) *step_dev = *step_dev + 1;
    }
  }
}

}  // namespace msgm

using namespace msgm;

extern "C" {

int msgm_p2p_create(msgm_ctx* ctx, int64_t nfloats, int32_t world, int32_t rank, msgm_p2p** out, unsigned char* handle_out) {
  if (!ctx || !out || !handle_out || nfloats < 1 || world < 1 || world > P2P_MAX_WORLD || rank < 0 || rank >= world) {
    set_error("msgm_p2p_create: bad argument (world <= 16)");
    return MSGM_ERR_INVALID;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  msgm_p2p* h = new msgm_p2p();
  std::memset(h, 0, sizeof(*h));
  h->world = world;
  h->rank = rank;
  h->total = (nfloats + 3) & ~(int64_t)3;
  const size_t bytes = p2p_alloc_bytes(world, h->total);
  cudaError_t e = cudaMalloc(&h->base, bytes);  // cudaMalloc, not a pool: the allocation is exported through CUDA IPC
  if (e == cudaSuccess) e = cudaMemset(h->base, 0, bytes);
  cudaIpcMemHandle_t ipc;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&ipc, h->base);
  if (e != cudaSuccess) {
    if (h->base) cudaFree(h->base);
    delete h;
    return cuda_fail(e, "msgm_p2p_create");
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  std::memcpy(handle_out, &ipc, 64);
  *out = h;
  return MSGM_OK;
}

int msgm_p2p_connect(msgm_ctx* ctx, msgm_p2p* h, const unsigned char* all_handles) {
  if (!ctx || !h || !all_handles) {
    set_error("msgm_p2p_connect: NULL argument");
    return MSGM_ERR_INVALID;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  for (int r = 0; r < h->world; ++r) {
    if (r == h->rank) {
      h->peer_base[r] = h->base;
    } else {
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, all_handles + 64 * r, 64);
      MSGM_CUDA_TRY(cudaIpcOpenMemHandle(&h->peer_base[r], ipc, cudaIpcMemLazyEnablePeerAccess));
    }
    unsigned char* b = reinterpret_cast<unsigned char*>(h->peer_base[r]);
    h->table.recv[r] = reinterpret_cast<float*>(b);
    h->table.flags[r] = reinterpret_cast<long long*>(b + p2p_flags_off(h->world, h->total));
  }
  unsigned char* mine = reinterpret_cast<unsigned char*>(h->base) + p2p_flags_off(h->world, h->total) + 8 * 2 * h->world;
  h->seq = reinterpret_cast<long long*>(mine);
  h->done = reinterpret_cast<unsigned int*>(mine + 64);
  return MSGM_OK;
}

int msgm_p2p_disconnect(msgm_p2p* h) {
  if (!h) return MSGM_OK;
  for (int r = 0; r < h->world; ++r)
    if (r != h->rank && h->peer_base[r]) {
      cudaIpcCloseMemHandle(h->peer_base[r]);
      h->peer_base[r] = nullptr;
    }
  h->seq = nullptr;
  return MSGM_OK;
}

int msgm_p2p_destroy(msgm_p2p* h) {
  if (!h) return MSGM_OK;
  msgm_p2p_disconnect(h);
  if (h->base) cudaFree(h->base);
  delete h;
  return MSGM_OK;
}

int msgm_p2p_allreduce_adam(msgm_ctx* ctx, msgm_p2p* h, const void* seg_table, int32_t n_tensors, int64_t total,
                            const float* grad_flat, float* exp_avg, float* exp_avg_sq, const float* lr_dev, int64_t* step_dev,
                            float beta1, float beta2, float eps, void* stream) {
  if (!ctx || !h || !seg_table || !grad_flat || !exp_avg || !exp_avg_sq || !lr_dev || !step_dev || n_tensors < 1 ||
      n_tensors > 1024 || total < 1 || total > h->total || !h->seq) {
    set_error("msgm_p2p_allreduce_adam: bad argument (connect the handle first; total <= the size it was created for)");
    return MSGM_ERR_INVALID;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<long long>((h->total / 4 + 255) / 256, (long long)ctx->num_sms * 2);
  // grad_flat must be readable for h->total floats (the caller pads its flat buffer to a multiple of 4)
  p2p_push_kernel<<<blocks, 256, 0, st>>>(h->table, h->world, h->rank, h->total, grad_flat, h->seq, h->done);
  const int ablocks = (int)std::min<long long>((total + 255) / 256, (long long)ctx->num_sms * 2);
  p2p_adam_kernel<<<ablocks, 256, sizeof(AdamSeg) * n_tensors, st>>>(h->table, h->world, h->rank, h->total,
                                                                    reinterpret_cast<const AdamSeg*>(seg_table), n_tensors, total,
                                                                    exp_avg, exp_avg_sq, lr_dev,
                                                                    reinterpret_cast<long long*>(step_dev), h->seq, h->done, beta1,
                                                                    beta2, eps, next_tc_flags(ctx));
  ctx->launches += 2;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // extern "C"
