// Development probe (not part of the library): MN-major ("transposed") shared-memory operands of tcgen05.mma.
// The SSM tensor-core kernel wants D = P^T Q (weight gradients: sum over the batch rows) and D = Z W (data gradients)
// from tiles that were written for K-major use.  A tile stored as [c/8][r/8][r%8][c%8] (r = row, c = column, fp16) read
// K-major is the operand X[r][c] (rows = MN, columns = K); read MN-major (instruction-descriptor bits 15 / 16) it should be
// X^T: columns = MN, rows = K, with SBO = stride between 8-column groups (2048 B) and LBO = stride between 8-row groups
// (128 B).  This program checks that on hardware for both operands.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mn_probe tools/mn_probe.cu ; run under gpurun.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
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
      : "r"(taddr) : "memory")

constexpr int R = 128, Cc = 128;  // tiles are R rows x Cc columns
__host__ __device__ inline int img_off(int r, int c) { return (c / 8) * (R * 8) + (r / 8) * 64 + (r % 8) * 8 + (c % 8); }

// mode 0: D = P Q^T      (both K-major; the known-good baseline)            D[m][n] = sum_c P[m][c] Q[n][c]
// mode 1: D = P^T Q      (both MN-major; K = rows)                          D[m][n] = sum_r P[r][m] Q[r][n]
// mode 2: D = P Q        (A K-major, B MN-major; K = columns of P = rows of Q)   D[m][n] = sum_k P[m][k] Q[k][n]
// swap: exchange the LBO / SBO fields of the MN-major descriptors
__global__ void __launch_bounds__(128, 1) probe(const __half* Pimg, const __half* Qimg, float* out, int mode, int swap, int* err) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __half* sP = reinterpret_cast<__half*>(smem);
  __half* sQ = sP + R * Cc;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sQ + R * Cc);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < R * Cc / 8; i += 128) {
    reinterpret_cast<uint4*>(sP)[i] = reinterpret_cast<const uint4*>(Pimg)[i];
    reinterpret_cast<uint4*>(sQ)[i] = reinterpret_cast<const uint4*>(Qimg)[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = *slot;
  if (tid == 0) {
    const bool a_mn = mode == 1, b_mn = mode >= 1;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24) | (a_mn ? (1u << 15) : 0u) |
                           (b_mn ? (1u << 16) : 0u);
    // K-major: LBO = 2048 (between the two 8-column chunks of a K=16 slice), SBO = 128 (between 8-row groups), slice += 4096 B
    // MN-major: SBO = 2048 (between 8-column groups = MN), LBO = 128 (between 8-row groups = K), slice += 256 B
    const uint32_t mn_lbo = swap ? 2048 : 128, mn_sbo = swap ? 128 : 2048;
    for (int s = 0; s < 8; ++s) {
      const uint64_t ad = a_mn ? make_desc(smem_u32(sP) + s * 256, mn_lbo, mn_sbo) : make_desc(smem_u32(sP) + s * 4096, 2048, 128);
      const uint64_t bd = b_mn ? make_desc(smem_u32(sQ) + s * 256, mn_lbo, mn_sbo) : make_desc(smem_u32(sQ) + s * 4096, 2048, 128);
      mma_ss(tb, ad, bd, idesc, s > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  bool ok = false;
  for (int i = 0; i < (1 << 22) && !ok; ++i) ok = mbar_try(bar, 0);
  if (!ok) atomicExch(err, 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < 4; ++c) {
    uint32_t r[32];
    TMEM_LD32(tb + ((uint32_t)(warp * 32) << 16) + c * 32, r);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 128 + c * 32 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(128));
}

int main() {
  std::vector<float> P(R * Cc), Q(R * Cc);
  std::vector<__half> Pi(R * Cc), Qi(R * Cc);
  srand(1);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < Cc; ++c) {
      const float a = (float)(rand() % 17 - 8) / 8.0f, b = (float)(rand() % 13 - 6) / 4.0f;  // exact in fp16
      P[r * Cc + c] = a; Q[r * Cc + c] = b;
      Pi[img_off(r, c)] = __float2half(a); Qi[img_off(r, c)] = __float2half(b);
    }
  __half *dP, *dQ; float* dO; int* dE;
  CK(cudaMalloc(&dP, R * Cc * 2)); CK(cudaMalloc(&dQ, R * Cc * 2)); CK(cudaMalloc(&dO, 128 * 128 * 4)); CK(cudaMalloc(&dE, 4));
  CK(cudaMemcpy(dP, Pi.data(), R * Cc * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dQ, Qi.data(), R * Cc * 2, cudaMemcpyHostToDevice));
  const int smem = 2 * R * Cc * 2 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> out(128 * 128);
  for (int mode = 0; mode < 3; ++mode)
    for (int swap = 0; swap < (mode ? 2 : 1); ++swap) {
      CK(cudaMemset(dE, 0, 4)); CK(cudaMemset(dO, 0, 128 * 128 * 4));
      probe<<<1, 128, smem>>>(dP, dQ, dO, mode, swap, dE);
      CK(cudaDeviceSynchronize());
      int e; CK(cudaMemcpy(&e, dE, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(out.data(), dO, 128 * 128 * 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
          double ref = 0;
          for (int k = 0; k < 128; ++k)
            ref += mode == 0 ? (double)P[m * Cc + k] * Q[n * Cc + k] : mode == 1 ? (double)P[k * Cc + m] * Q[k * Cc + n]
                                                                                 : (double)P[m * Cc + k] * Q[k * Cc + n];
          worst = fmax(worst, fabs(ref - out[m * 128 + n]));
        }
      printf("mode %d (%s) swap %d: timeout %d, max |D - ref| = %.4g\n", mode,
             mode == 0 ? "P Q^T, K-major both" : mode == 1 ? "P^T Q, MN-major both" : "P Q, A K-major / B MN-major", swap, e, worst);
    }
  return 0;
}
