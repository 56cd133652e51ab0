// tcgen05 / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sampler_tc.cu, conv2d_tc.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

#include "msgm_common.cuh"

namespace msgm {

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Error reporting of the tensor-core kernels.  `abort` is a device word holding the id of the launch that gave up (so a
// stale value from an earlier launch never makes a later one bail out, and nothing has to be cleared per launch);
// `sticky` is a word in mapped pinned HOST memory that receives the error code: the host reads it at its next API call
// without synchronising (msgm_async_error) and raises.  Codes: 1 = a bounded mbarrier wait timed out, 2 = shared-memory /
// TMEM base assumption violated.
struct TcFlags {
  int* abort;
  volatile int* sticky;
  int id;
};
// Host side: the flags of the next tensor-core launch of this context.
inline TcFlags next_tc_flags(msgm_ctx* ctx) {
  ctx->launch_seq = ctx->launch_seq == 0x7fffffff ? 1 : ctx->launch_seq + 1;
  return TcFlags{reinterpret_cast<int*>(ctx->ws), ctx->host_flag_dev, ctx->launch_seq};
}
__device__ __forceinline__ void tc_raise(const TcFlags& f, int code) {
  atomicExch(f.abort, f.id);
  *f.sticky = code;
  __threadfence_system();
}
// Bounded wait: returns false (and raises the error) instead of hanging if the partner never arrives.
// Normal waits last microseconds; the limit is ~0.2 s, and a raised abort word makes every other waiter of the same launch
// bail out at once.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, const TcFlags& f) {
  if (mbar_try(bar, parity)) return true;
  const long long t0 = clock64();
  int spins = 0;
  while (clock64() - t0 < 400000000LL) {
    if (mbar_try(bar, parity)) return true;
    if ((++spins & 255) == 0 && *reinterpret_cast<volatile int*>(f.abort) == f.id) return false;
  }
  tc_raise(f, 1);
  return false;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle smem matrix descriptor (validated on B200 by tools/tc_probe.cu):
// LBO = bytes between the two 8-element k-chunks of a K=16 slice, SBO = bytes between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ constexpr uint32_t umma_idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // D=f32, A=B=f16, K-major both
}
// The MMA warp runs converged; only the instruction itself is predicated on the leader lane (`lead` != 0 in one lane).
// The issuing warp runs converged; one lane chosen by elect.sync (which ptxas turns into a uniform predicate, so no
// per-instruction "waterfall" loop is generated) issues the instruction on behalf of the CTA.  `lead` is unused padding
// kept for call-site symmetry.
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc,
                                        uint32_t /*lead*/) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint32_t /*lead*/) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}

#define TMEM_LD32(taddr, r)                                                                                         \
  asm volatile(                                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                     \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
      "%29,%30,%31}, [%32];"                                                                                        \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),      \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),     \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                   \
      : "r"(taddr)                                                                                                  \
      : "memory")
#define TMEM_LD16(taddr, r)                                                                                      \
  asm volatile(                                                                                                  \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"   \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),           \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])     \
      : "r"(taddr)                                                                                               \
      : "memory")
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
// Two fp32 values -> packed fp16 (hi, hi) and packed fp16 residuals (lo, lo).  Uses the paired conversion
// (F2FP.PACK_AB) and HADD2.F32 only: the scalar cvt.f16.f32 is an XU-pipe op (F2F) and would queue behind the tanh stream.
__device__ __forceinline__ void split2_f16(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  hi2 = pack_f16x2(a, b);
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hi2));
  lo2 = pack_f16x2(a - f.x, b - f.y);
}
__device__ __forceinline__ uint32_t pack_h2(__half lo, __half hi) {
  return (uint32_t)__half_as_ushort(lo) | ((uint32_t)__half_as_ushort(hi) << 16);
}

}  // namespace msgm
