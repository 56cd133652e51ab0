// Shared device helpers for libmsgm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/msgm_b200.h"

namespace msgm {

constexpr int HID = MSGM_HIDDEN;
constexpr float SQRT_HALF = 0.70710678118654752440f;  // c of the cyclic sparse tensor (SDEs.py:379)

// ---- error plumbing (host) -------------------------------------------------------------------------------
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
#define MSGM_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return msgm::cuda_fail(_e, #expr); \
  } while (0)

}  // namespace msgm

struct msgm_ctx;
namespace msgm {
int dyn_smem_base(msgm_ctx* ctx, cudaStream_t stream, uint32_t* out);  // sampler_tc.cu
}

struct msgm_ctx {
  int device;
  int num_sms;
  int64_t launches;
  void* ws;          // device workspace (packed fp16 weights for the tensor-core path, ...)
  size_t ws_bytes;
  int* host_flag;      // mapped pinned host word: error code raised by a tensor-core kernel (tc_ptx.cuh, TcFlags)
  int* host_flag_dev;  // its device alias
  int launch_seq;      // ids of the tensor-core launches (1, 2, ...)
  const unsigned int* tc_in_amax;  // one-shot: range scaling of the NEXT tensor-core conv launch (msgm_tc_range_scale)
};

namespace msgm {

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (c0..c3), key = (k0,k1) --------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0;
    k.y += W1;
  }
  return c;
}

// Four independent standard normals for (seed, particle, step, block): Box-Muller on the Philox output.
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t particle, uint32_t step, uint32_t blk) {
  uint4 r = philox4x32_10(make_uint4((uint32_t)particle, (uint32_t)(particle >> 32), step, blk),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  constexpr float TWO_NEG32 = 2.3283064365386963e-10f;
  // u in (0,1]: (r + 0.5) * 2^-32 rounded can hit 1.0 but never 0
  float u0 = ((float)r.x + 0.5f) * TWO_NEG32, u1 = ((float)r.y + 0.5f) * TWO_NEG32;
  float u2 = ((float)r.z + 0.5f) * TWO_NEG32, u3 = ((float)r.w + 0.5f) * TWO_NEG32;
  float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(ra * c0, ra * s0, rb * c1, rb * s1);
}

// Uniform in [0,1) from Philox word
__device__ __forceinline__ float u01(uint32_t r) { return (float)(r >> 8) * 5.9604644775390625e-08f; }

// fp32 ops that must not be contracted into FMAs, to follow the reference's rounding of time/beta scalars
__device__ __forceinline__ float beta_of(float bmin, float bdel, float s) {
  return __fadd_rn(bmin, __fmul_rn(bdel, s));  // SDEs.py:72-73
}

}  // namespace msgm
