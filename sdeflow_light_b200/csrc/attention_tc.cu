// Single-head attention of the 2-D U-Net's AttentionBlock (QKVAttention, model/unet.py:236-250) on tcgen05.
//
//   qkv (B, 3C, T) fp32  ->  out (B, C, T),  out[:, t] = sum_s softmax_s(q_t . k_s / sqrt(C)) v[:, s]
//
// CTA = one sample x 128 queries.  Both contractions run on the tensor pipe with operands split into fp16 hi + lo
// (three products each: fp32-level parity, as in conv2d_tc.cu):
//   1. S = Q^T K       M = 128 queries, N = T keys (<= 256), K = C channels: accumulator = T TMEM columns;
//   2. softmax on TMEM rows: thread = query row; pass 1 row max, pass 2 e = exp((s - max) / sqrt(C)), row sum, and the
//      unnormalised e split hi/lo written as the A operand of the next product, 64 keys at a time;
//   3. O = P V         M = 128, N = C, K = T: accumulator = C TMEM columns after S; 1 / rowsum applied in the epilogue.
// q and k are channel-major in memory, which IS the K-major core-matrix order after an 8-channel gather (lanes =
// consecutive tokens: coalesced); v[c, s..s+7] is already contiguous.  The P tiles reuse the Q / K operand space.
//   4. (optional) the block's 1x1 output projection and residual in the same launch: the normalised O rows are split and
//      written as an A operand into the Q space, the projection's packed weights are copied into the K | V space, one more
//      product (M = 128, N = C, K = C) lands in the S columns, and the epilogue adds bias and the block input.  Saves the
//      separate conv launch and the round trip of the attention output (11 launches of ~15 us per 2-D U-Net forward).
#include <cuda_fp16.h>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

struct AttnTcParams {
  const float* qkv;
  float* out;
  int C, T;
  float scale2;
  TcFlags flags;
  // fused output projection (AttentionBlock.proj_out + residual, model/unet.py:228-234): out = W_proj a + bias + res.
  // wimg = the 1x1 conv's packed image (conv2d_tc_pack_kernel, one N tile of C columns) or NULL (out = a)
  const unsigned char* wimg;
  const float* pbias;
  const float* res;
};

constexpr int ATC_THREADS = 256;
constexpr int ATC_PBUF = 32768;  // one P chunk: 64 keys x 128 queries, fp16 hi (16 KB) + lo (16 KB)

__global__ void __launch_bounds__(ATC_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams P) {
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  uint64_t* bar_s = reinterpret_cast<uint64_t*>(smem);   // S complete
  uint64_t* bar_p = bar_s + 1;                           // [4] P chunk consumed (its MMAs complete)
  uint64_t* bar_w = bar_s + 5;                           // projection product complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 6);
  const int C = P.C, T = P.T;
  const int QB = C * 128 * 2, KB = C * T * 2;            // bytes of one hi (or lo) plane set
  unsigned char* sQ = smem + 128;                        // [hi|lo][C/8][128][8]
  unsigned char* sK = sQ + 2 * QB;                       // [hi|lo][C/8][T][8]
  unsigned char* sV = sK + 2 * KB;                       // [hi|lo][T/8][C][8]
  unsigned char* sP = sQ;                                // P chunks reuse the Q | K space once S is complete
  const int npbuf = min(3, (2 * QB + 2 * KB) / ATC_PBUF);

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // broadcast form: the role branches are provably warp-uniform
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  const float* q = P.qkv + (size_t)b * 3 * C * T;
  const float* k = q + (size_t)C * T;
  const float* v = k + (size_t)C * T;

  if (tid == 0) {
    mbar_init(bar_s, 1);
    for (int i = 0; i < 4; ++i) mbar_init(bar_p + i, 1);
    mbar_init(bar_w, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ---- stage Q (this tile's queries), K, V: fp32 -> fp16 hi / lo core-matrix planes ----------------------------------
  // ATC_ILP items (8 values each) per thread and step: every global load of a step is issued before the first conversion,
  // so a thread pays one memory round trip per ATC_ILP items instead of one per item (staging was ~half of the kernel).
  constexpr int ATC_ILP = 4;
  for (int e0 = tid; e0 < (C / 8) * 128; e0 += ATC_ILP * ATC_THREADS) {
    float x[ATC_ILP][8];
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS, kc = e >> 7, t = e & 127;
      if (e < (C / 8) * 128 && t0 + t < T) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = __ldg(q + (size_t)(kc * 8 + j) * T + t0 + t);
      }
    }
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS, t = e & 127;
      if (e >= (C / 8) * 128) break;
      uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
      if (t0 + t < T) {
        split2_f16(x[u][0], x[u][1], hi.x, lo.x); split2_f16(x[u][2], x[u][3], hi.y, lo.y);
        split2_f16(x[u][4], x[u][5], hi.z, lo.z); split2_f16(x[u][6], x[u][7], hi.w, lo.w);
      }
      *reinterpret_cast<uint4*>(sQ + e * 16) = hi;
      *reinterpret_cast<uint4*>(sQ + QB + e * 16) = lo;
    }
  }
  for (int e0 = tid; e0 < (C / 8) * T; e0 += ATC_ILP * ATC_THREADS) {
    float x[ATC_ILP][8];
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS;
      if (e < (C / 8) * T) {
        const int kc = e / T, sidx = e % T;
#pragma unroll
        for (int j = 0; j < 8; ++j) x[u][j] = __ldg(k + (size_t)(kc * 8 + j) * T + sidx);
      }
    }
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS;
      if (e >= (C / 8) * T) break;
      uint4 hi, lo;
      split2_f16(x[u][0], x[u][1], hi.x, lo.x); split2_f16(x[u][2], x[u][3], hi.y, lo.y);
      split2_f16(x[u][4], x[u][5], hi.z, lo.z); split2_f16(x[u][6], x[u][7], hi.w, lo.w);
      *reinterpret_cast<uint4*>(sK + e * 16) = hi;
      *reinterpret_cast<uint4*>(sK + KB + e * 16) = lo;
    }
  }
  for (int e0 = tid; e0 < (T / 8) * C; e0 += ATC_ILP * ATC_THREADS) {
    float4 xa[ATC_ILP], xb[ATC_ILP];
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS;
      if (e < (T / 8) * C) {
        const int sc = e / C, c = e % C;
        xa[u] = __ldg(reinterpret_cast<const float4*>(v + (size_t)c * T + sc * 8));
        xb[u] = __ldg(reinterpret_cast<const float4*>(v + (size_t)c * T + sc * 8 + 4));
      }
    }
#pragma unroll
    for (int u = 0; u < ATC_ILP; ++u) {
      const int e = e0 + u * ATC_THREADS;
      if (e >= (T / 8) * C) break;
      uint4 hi, lo;
      split2_f16(xa[u].x, xa[u].y, hi.x, lo.x); split2_f16(xa[u].z, xa[u].w, hi.y, lo.y);
      split2_f16(xb[u].x, xb[u].y, hi.z, lo.z); split2_f16(xb[u].z, xb[u].w, hi.w, lo.w);
      *reinterpret_cast<uint4*>(sV + e * 16) = hi;
      *reinterpret_cast<uint4*>(sV + KB + e * 16) = lo;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  // ---- S = Q^T K --------------------------------------------------------------------------------------------------------
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(128, T);
    const uint32_t qh = smem_u32(sQ), kh = smem_u32(sK);
    for (int kk = 0; kk < C / 16; ++kk) {
      const uint64_t dQh = umma_desc(qh + kk * 4096, 2048, 128), dQl = umma_desc(qh + QB + kk * 4096, 2048, 128);
      const uint64_t dKh = umma_desc(kh + kk * 2 * T * 16, T * 16, 128), dKl = umma_desc(kh + KB + kk * 2 * T * 16, T * 16, 128);
      umma_ss(tbase, dQh, dKh, idesc, kk > 0 ? 1u : 0u, 0);
      umma_ss(tbase, dQl, dKh, idesc, 1u, 0);
      umma_ss(tbase, dQh, dKl, idesc, 1u, 0);
    }
    umma_commit(bar_s, 0);
    __syncwarp();
  }

  bool ok = true;
  if (warp < 4) {
    // ---- softmax rows + O = P V -----------------------------------------------------------------------------------------
    ok = mbar_wait(bar_s, 0, P.flags);
    tc_fence_after();
    const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
    const int row = tid;  // query row == TMEM lane
    float mx = -INFINITY;
    for (int c0 = 0; c0 < T; c0 += 32) {
      uint32_t r[32];
      TMEM_LD32(trow + c0, r);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
    }
    float sum = 0.0f;
    const uint32_t idesc_o = umma_idesc_f16(128, C);
    const uint32_t ocol = tbase + 256;
    const int nchunk = T / 64;
    for (int ch = 0; ch < nchunk && ok; ++ch) {
      const int pb = ch % npbuf;
      if (ch >= npbuf) {  // the buffer's previous chunk must have been consumed by its MMAs
        ok = mbar_wait(bar_p + pb, (uint32_t)(((ch / npbuf) - 1) & 1), P.flags);
        tc_fence_after();
      }
      unsigned char* pbuf = sP + pb * ATC_PBUF;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        TMEM_LD32(trow + ch * 64 + h * 32, r);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 8 keys -> one 16-byte row of k-chunk (h * 4 + g)
          float e[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            e[j] = __expf((__uint_as_float(r[g * 8 + j]) - mx) * P.scale2);
            sum += e[j];
          }
          uint4 hi, lo;
          split2_f16(e[0], e[1], hi.x, lo.x); split2_f16(e[2], e[3], hi.y, lo.y);
          split2_f16(e[4], e[5], hi.z, lo.z); split2_f16(e[6], e[7], hi.w, lo.w);
          *reinterpret_cast<uint4*>(pbuf + (h * 4 + g) * 2048 + row * 16) = hi;
          *reinterpret_cast<uint4*>(pbuf + 16384 + (h * 4 + g) * 2048 + row * 16) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four row warps
      if (warp == 0) {
        tc_fence_after();
        const uint32_t ph = smem_u32(pbuf), vh = smem_u32(sV);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t dPh = umma_desc(ph + ks * 4096, 2048, 128), dPl = umma_desc(ph + 16384 + ks * 4096, 2048, 128);
          const uint32_t voff = (uint32_t)((ch * 8 + ks * 2) * C * 16);
          const uint64_t dVh = umma_desc(vh + voff, C * 16, 128), dVl = umma_desc(vh + KB + voff, C * 16, 128);
          umma_ss(ocol, dPh, dVh, idesc_o, (ch > 0 || ks > 0) ? 1u : 0u, 0);
          umma_ss(ocol, dPl, dVh, idesc_o, 1u, 0);
          umma_ss(ocol, dPh, dVl, idesc_o, 1u, 0);
        }
        umma_commit(bar_p + pb, 0);
        __syncwarp();
      }
    }
    // every chunk's MMAs complete (commits are ordered: the last one covers all earlier tcgen05 operations)
    if (ok) {
      const int last = nchunk - 1;
      ok = mbar_wait(bar_p + last % npbuf, (uint32_t)((last / npbuf) & 1), P.flags);
      tc_fence_after();
    }
    const float inv = 1.0f / sum;
    const bool live = ok && t0 + row < T;
    if (!P.wimg) {
      float* o = P.out + (size_t)b * C * T + t0 + row;
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        TMEM_LD32(trow + 256 + c0, r);
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)(c0 + j) * T] = __uint_as_float(r[j]) * inv;
        }
      }
    } else {
      // normalised attention output of this row -> A operand of the projection, [hi|lo][C/8][128][8] in the Q space (every
      // product that read Q, K, V or a P chunk has completed: the last commit covers all earlier tcgen05 operations)
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        TMEM_LD32(trow + 256 + c0, r);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 hi, lo;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(r[g * 8 + j]) * inv;
          split2_f16(f[0], f[1], hi.x, lo.x); split2_f16(f[2], f[3], hi.y, lo.y);
          split2_f16(f[4], f[5], hi.z, lo.z); split2_f16(f[6], f[7], hi.w, lo.w);
          unsigned char* dst = sQ + ((c0 >> 3) + g) * 2048 + row * 16;
          *reinterpret_cast<uint4*>(dst) = hi;
          *reinterpret_cast<uint4*>(dst + QB) = lo;
        }
      }
    }
  }
  if (P.wimg) {  // launch-uniform
    __syncthreads();  // every warp now knows that K and V are dead
    {
      const uint4* src = reinterpret_cast<const uint4*>(P.wimg);
      uint4* dst = reinterpret_cast<uint4*>(sK);
      for (int e = tid; e < C * C / 4; e += ATC_THREADS) dst[e] = __ldg(src + e);  // 4 C^2 bytes: [chunk][hi|lo][kc][C][8]
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      const uint32_t idesc_c = umma_idesc_f16(128, C);
      const uint32_t ah = smem_u32(sQ), wh = smem_u32(sK);
      for (int kk = 0; kk < C / 16; ++kk) {
        const uint64_t dAh = umma_desc(ah + kk * 4096, 2048, 128), dAl = umma_desc(ah + QB + kk * 4096, 2048, 128);
        const uint32_t wk = wh + (uint32_t)(kk * 64 * C);
        const uint64_t dWh = umma_desc(wk, C * 16, 128), dWl = umma_desc(wk + 32 * C, C * 16, 128);
        umma_ss(tbase, dAh, dWh, idesc_c, kk > 0 ? 1u : 0u, 0);
        umma_ss(tbase, dAl, dWh, idesc_c, 1u, 0);
        umma_ss(tbase, dAh, dWl, idesc_c, 1u, 0);
      }
      umma_commit(bar_w, 0);
      __syncwarp();
    }
    if (warp < 4) {
      ok = mbar_wait(bar_w, 0, P.flags) && ok;
      tc_fence_after();
      const int row = tid;
      const bool live = ok && t0 + row < T;
      const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
      const size_t base = (size_t)b * C * T + t0 + row;
      for (int c0 = 0; c0 < C; c0 += 32) {
        uint32_t r[32];
        TMEM_LD32(trow + c0, r);
        tc_wait_ld();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]);
            if (P.pbias) v += __ldg(P.pbias + c0 + j);
            if (P.res) v += __ldg(P.res + base + (size_t)(c0 + j) * T);
            P.out[base + (size_t)(c0 + j) * T] = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

bool attention_tc_ok(int C, int T) {
  if (C % 32 || C < 32 || C > 128 || T % 64 || T < 64 || T > 256) return false;
  return 128 + 128 + 4 * (size_t)C * (128 + 2 * T) <= 227 * 1024 && (4 * C * 128 + 4 * C * T) >= ATC_PBUF;
}

int conv2d_tc_nout(int Cout, int taps);  // conv2d_tc.cu

// the fused projection needs the 1x1 conv's packed image to be ONE N tile of C columns
bool attention_proj_tc_ok(int C, int T) { return attention_tc_ok(C, T) && conv2d_tc_nout(C, 1) == C && C <= 2 * T; }

int attention_tc(msgm_ctx* ctx, const float* qkv, float* out, int B, int C, int T, cudaStream_t stream, const void* wimg,
                 const float* pbias, const float* res) {
  AttnTcParams P{qkv, out, C, T, 1.0f / sqrtf((float)C), next_tc_flags(ctx), reinterpret_cast<const unsigned char*>(wimg), pbias, res};
  const size_t smem = 128 + 128 + 4 * (size_t)C * (128 + 2 * T);
  MSGM_CUDA_TRY(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<<<dim3((T + 127) / 128, B), ATC_THREADS, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
