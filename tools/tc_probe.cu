// Development probe for the tcgen05 path (not part of the library).  Checks, on a real B200:
//   1. smem-descriptor semantics for the no-swizzle K-major core-matrix layout used by sampler_tc.cu,
//   2. the TMEM layout of an f16 A operand (lane = row, 32-bit column = two consecutive k),
//   3. cycles of an 8-slice 128x128x128 MMA chain and of the TMEM ld/st epilogue traffic,
//   4. MUFU throughput of tanh.approx.f16x2 vs tanh.approx.f32.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu ; run under gpurun.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
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
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity, int* err) {
  for (int i = 0; i < (1 << 20); ++i)
    if (mbar_try(bar, parity)) return true;
  atomicExch(err, 1);
  return false;
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (Blackwell)
  return d;                // layout_type = 0 (no swizzle), base_offset = 0
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4)                     // D = f32
         | (0u << 7) | (0u << 10)      // A, B = f16
         | ((uint32_t)(N >> 3) << 17)  // N / 8
         | ((uint32_t)(M >> 4) << 24); // M / 16
}

__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
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

#define TMEM_ST16(taddr, r)                                                                                      \
  asm volatile(                                                                                                  \
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"   \
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),      \
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])             \
      : "memory")

constexpr int M = 128, N = 128, K = 128;
// core-matrix image: offset(row, k) in halves
__host__ __device__ inline int img_off(int row, int k) { return (k / 8) * (N * 8) + (row / 8) * 64 + (row % 8) * 8 + (k % 8); }

// variant bit0: 0 = A from TMEM (TS), 1 = A from smem (SS).  bit1: swap LBO/SBO fields.
__global__ void __launch_bounds__(128, 1)
probe_mma(const __half* __restrict__ Aimg, const __half* __restrict__ Arow, const __half* __restrict__ Bimg,
          float* __restrict__ out, int variant, int reps, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __half* sB = reinterpret_cast<__half*>(smem_raw);            // 32 KB
  __half* sA = sB + N * K;                                      // 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sA + M * K);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < N * K / 8; i += 128) {
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
    reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(Aimg)[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // make generic-proxy smem writes visible to the async (tensor core) proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  const uint32_t D_COL = 0, A_COL = 128;

  // ---- A into TMEM: thread = row; column j holds (A[row][2j], A[row][2j+1]) ---------------------------------
  long long t0 = clock64();
  {
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(Arow + (size_t)tid * K);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) r[j] = arow[c * 16 + j];
      TMEM_ST16(lane_base + A_COL + c * 16, r);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  long long t1 = clock64();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  const uint32_t lbo = (variant & 2) ? 128u : (uint32_t)(N * 16);  // bytes between the two k-chunks of a slice
  const uint32_t sbo = (variant & 2) ? (uint32_t)(N * 16) : 128u;  // bytes between 8-row groups
  const int Nmma = (variant & 16) ? 64 : ((variant & 32) ? 32 : N);
  const uint32_t idesc = make_idesc(M, Nmma);
  long long t2 = clock64();
  long long t_issued = t2;
  if (tid == 0) {
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
      for (int s = 0; s < K / 16; ++s) {
        const uint64_t bd = make_desc(smem_u32(sB) + s * (2 * N * 16), lbo, sbo);
        if (variant & 1) {
          const uint64_t ad = make_desc(smem_u32(sA) + s * (2 * M * 16), lbo, sbo);
          mma_ss(tbase + D_COL, ad, bd, idesc, (s > 0) ? 1u : 0u);
        } else {
          mma_ts(tbase + D_COL, tbase + A_COL + s * 8, bd, idesc, (s > 0) ? 1u : 0u);
        }
      }
    }
    t_issued = clock64();
    mma_commit(bar);
  }
  bool ok = mbar_wait_bounded(bar, 0, err);
  long long t3 = clock64();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (ok) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      TMEM_LD32(lane_base + D_COL + c * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) out[(size_t)tid * N + c * 32 + j] = __uint_as_float(r[j]);
    }
  }
  long long t4 = clock64();
  // pure TMEM epilogue traffic: 16 x (load 128 accumulator columns, store 64 packed columns), no math
  uint32_t sink = 0;
  for (int it = 0; it < 16; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      TMEM_LD32(lane_base + D_COL + c * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      uint32_t q[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) q[j] = r[2 * j] ^ r[2 * j + 1];
      sink ^= q[3];
      TMEM_ST16(lane_base + 256 + c * 16, q);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  long long t5 = clock64();
  if (sink == 0x12345678u) out[0] = 1.f;
  if (tid == 0) cycles[3] = (t5 - t4) / 16;
  if (tid == 0) {
    cycles[0] = t1 - t0;  // A store (64 cols)
    cycles[1] = t3 - t2;  // MMA chain issue -> completion observed
    cycles[4] = t_issued - t2;  // time spent issuing
    cycles[2] = t4 - t3;  // D load (128 cols) + global stores
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ---- issue patterns: how fast can 64 SS MMAs be issued? ------------------------------------------------------------------
__device__ __forceinline__ void mma_ss_pred(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc, uint32_t lead) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(lead)
      : "memory");
}
// mode 0: one lane in a divergent branch, descriptors rebuilt per MMA; 1: same, descriptors = base + s*256;
// mode 2: converged warp, instruction predicated on lane 0; 3: one lane, 4 MMAs per descriptor pair reused (no desc math)
__global__ void __launch_bounds__(128, 1) probe_issue(int mode, long long* cycles, int* err) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 16; i += 128) reinterpret_cast<uint4*>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tmem_slot;
  const uint32_t idesc = make_idesc(128, 128);
  const uint32_t sA = smem_u32(smem_raw), sB = sA + 32768;
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    t0 = clock64();
    if (mode == 2) {
      const uint32_t lead = lane == 0;
      const uint64_t a0 = make_desc(sA, 2048, 128), b0 = make_desc(sB, 2048, 128);
      for (int rep = 0; rep < 8; ++rep)
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2) mma_ss_pred(tbase, a0 + s2 * 256, b0 + s2 * 256, idesc, s2 > 0, lead);
      t1 = clock64();
      if (lane == 0) mma_commit(bar);
    } else if (lane == 0) {
      const uint64_t a0 = make_desc(sA, 2048, 128), b0 = make_desc(sB, 2048, 128);
      for (int rep = 0; rep < 8; ++rep)
#pragma unroll
        for (int s2 = 0; s2 < 8; ++s2) {
          if (mode == 0) mma_ss(tbase, make_desc(sA + s2 * 4096, 2048, 128), make_desc(sB + s2 * 4096, 2048, 128), idesc, s2 > 0);
          else if (mode == 1) mma_ss(tbase, a0 + s2 * 256, b0 + s2 * 256, idesc, s2 > 0);
          else mma_ss(tbase, a0, b0, idesc, s2 > 0);
        }
      t1 = clock64();
      mma_commit(bar);
    }
  }
  bool ok = mbar_wait_bounded(bar, 0, err);
  long long t2 = clock64();
  if (warp == 1 && lane == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t0; }
  (void)ok;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

// ---- MUFU throughput ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) probe_mufu(float* out, int iters) {
  float acc = 0.f;
  if (MODE == 0) {  // tanh.approx.f32
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  } else if (MODE == 1) {  // tanh.approx.f16x2
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0x30003000u + threadIdx.x + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(v[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += __uint_as_float(v[j]);
  } else if (MODE == 2) {  // ex2.approx.f32
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j];
  } else if (MODE == 4) {  // F2FP only: cvt.rn.f16x2.f32
    float a[8];
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[j]), "f"(a[(j + 1) & 7]));
        o ^= h;
        a[j] = __uint_as_float(__float_as_uint(a[j]) ^ (h & 1));
      }
    }
    acc = __uint_as_float(o);
  } else if (MODE == 5) {  // the sampler's epilogue mix per pair: 2 tanh.f32 + 2 ffma + 1 cvt.f16x2
    float a[8];
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 0.001f * (threadIdx.x + j);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        float t0, t1;
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a[j]));
        asm volatile("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a[j + 1]));
        const float s0 = fmaf(a[j], t0, a[j]), s1 = fmaf(a[j + 1], t1, a[j + 1]);
        uint32_t h;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(s1), "f"(s0));
        o ^= h;
        a[j] += 1e-3f;
        a[j + 1] += 1e-3f;
      }
    }
    acc = __uint_as_float(o);
  } else {  // MODE 3: full packed swish epilogue: cvt.f16x2 + tanh.f16x2 + hfma2 on 2 inputs
    float a[8], b[8];
    uint32_t o = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = 0.001f * (threadIdx.x + j); b[j] = a[j] + 0.5f; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t h, t, r;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[j]), "f"(b[j]));
        asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(h));
        asm volatile("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(h), "r"(t));
        o ^= r;
        a[j] += 1e-3f;
      }
    }
    acc = __uint_as_float(o);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
static void run_mufu(const char* name, int ops_per_iter_per_thread) {
  float* out;
  const int blocks = 148 * 8, threads = 256, iters = 4096;
  CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  probe_mufu<MODE><<<blocks, threads>>>(out, 16);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe_mufu<MODE><<<blocks, threads>>>(out, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)blocks * threads * iters * ops_per_iter_per_thread;
  int clk_khz;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf("mufu %-28s %.3f ms  %.2f Gop/s  = %.2f thread-ops/clk/SM at max clock %d MHz\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / 148.0 / (clk_khz * 1e3), clk_khz / 1000);
  cudaFree(out);
}

int main() {
  std::vector<__half> A(M * K), B(N * K), Aimg(M * K), Bimg(N * K);
  std::vector<float> Af(M * K), Bf(N * K), ref(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { Af[i] = (rand() % 2001 - 1000) / 1000.0f; A[i] = __float2half(Af[i]); Af[i] = __half2float(A[i]); }
  for (int i = 0; i < N * K; ++i) { Bf[i] = (rand() % 2001 - 1000) / 4000.0f; B[i] = __float2half(Bf[i]); Bf[i] = __half2float(B[i]); }
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) Aimg[img_off(m, k)] = A[m * K + k];
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) Bimg[img_off(n, k)] = B[n * K + k];
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)Af[m * K + k] * Bf[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __half *dA, *dAimg, *dBimg;
  float* dout;
  long long* dcyc;
  int* derr;
  CK(cudaMalloc(&dA, sizeof(__half) * M * K));
  CK(cudaMalloc(&dAimg, sizeof(__half) * M * K));
  CK(cudaMalloc(&dBimg, sizeof(__half) * N * K));
  CK(cudaMalloc(&dout, sizeof(float) * M * N));
  CK(cudaMalloc(&dcyc, sizeof(long long) * 8));
  CK(cudaMalloc(&derr, sizeof(int)));
  CK(cudaMemcpy(dA, A.data(), sizeof(__half) * M * K, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAimg, Aimg.data(), sizeof(__half) * M * K, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBimg, Bimg.data(), sizeof(__half) * N * K, cudaMemcpyHostToDevice));
  const int smem = 2 * N * K * 2 + 64;
  CK(cudaFuncSetAttribute(probe_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const char* names[4] = {"TS lbo=kchunk sbo=rowgrp", "SS lbo=kchunk sbo=rowgrp", "TS swapped", "SS swapped"};
  for (int variant = 0; variant < 2; ++variant) {
    for (int reps : {1, 8}) {
      CK(cudaMemset(dout, 0, sizeof(float) * M * N));
      CK(cudaMemset(derr, 0, sizeof(int)));
      probe_mma<<<1, 128, smem>>>(dAimg, dA, dBimg, dout, variant, reps, dcyc, derr);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("variant %d (%s): CUDA error %s\n", variant, names[variant], cudaGetErrorString(e));
        return 3;
      }
      std::vector<float> out(M * N);
      long long cyc[8];
      int err;
      CK(cudaMemcpy(out.data(), dout, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&err, derr, sizeof(int), cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
      printf("variant %d (%-26s) reps=%d: timeout=%d max|err|=%.3e  [%s]  cycles: A-st=%lld mma=%lld D-ld=%lld tmem-epi-traffic/layer=%lld\n", variant,
             names[variant], reps, err, maxerr, maxerr < 1e-3 ? "MATCH" : "mismatch", cyc[0], cyc[1], cyc[2], cyc[3]);
    }
  }
  // MMA rate vs N (results are only checked for N = 128): issue time vs completion time of 64 back-to-back MMAs
  for (int variant : {0, 1, 16, 17, 32, 33}) {
    CK(cudaMemset(derr, 0, sizeof(int)));
    probe_mma<<<1, 128, smem>>>(dAimg, dA, dBimg, dout, variant, 8, dcyc, derr);
    CK(cudaDeviceSynchronize());
    long long cyc[8];
    CK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    printf("mma-rate %s N=%d: 64 MMAs issue=%lld clk, issue->complete=%lld clk  (%.1f clk/MMA)\n", (variant & 1) ? "SS" : "TS",
           (variant & 16) ? 64 : ((variant & 32) ? 32 : 128), cyc[4], cyc[1], cyc[1] / 64.0);
  }
  {
    CK(cudaFuncSetAttribute(probe_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64));
    const char* mn[4] = {"one lane, desc rebuilt", "one lane, desc = base + s*256", "converged warp, @lane0 predicate", "one lane, same desc (no math)"};
    for (int mode = 0; mode < 4; ++mode) {
      CK(cudaMemset(derr, 0, sizeof(int)));
      probe_issue<<<1, 128, 65536 + 64>>>(mode, dcyc, derr);
      CK(cudaDeviceSynchronize());
      long long cyc[8];
      CK(cudaMemcpy(cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
      printf("issue-pattern SS N=128 (%s): 64 MMAs issued in %lld clk (%.1f/MMA), complete after %lld clk\n", mn[mode], cyc[0],
             cyc[0] / 64.0, cyc[1]);
    }
  }
  run_mufu<0>("tanh.approx.f32", 8);
  run_mufu<1>("tanh.approx.f16x2 (x2 elems)", 8);
  run_mufu<2>("ex2.approx.f32", 8);
  run_mufu<3>("cvt+tanh.f16x2+hfma2 (pairs)", 8);
  run_mufu<4>("cvt.rn.f16x2.f32 only", 8);
  run_mufu<5>("2 tanh.f32 + 2 ffma + cvt (elems)", 8);
  return 0;
}
