// 2-D U-Net convolutions on tcgen05 (model/unet.py:40-250: ResBlock / AttentionBlock / Upsample / Downsample convs).
//
// The conv is evaluated as a "shift GEMM" over a zero-padded LINEAR index space.  Every image is laid out with its
// padding ring, (Hp, Wp) = (Hi + 2 PAD, Wi + 2 PAD), and all images are stacked: p = (b Hp + r) Wp + c.  For an interior
// position p the 3x3 tap (ky, kx) reads position p + (ky-1) Wp + (kx-1) of the same image, so
//     D[p, co] = sum_tap sum_ci A[p + off_tap, ci] W[co, ci, tap]
// is nine GEMMs whose A operands are the SAME staged tile read at nine start addresses: in the canonical no-swizzle
// K-major core-matrix layout with SBO = 128 B a row is 16 bytes, so a tap shift is a 16-byte multiple added to the
// descriptor's start address.  No im2col copy exists anywhere.  Positions on the padding ring produce garbage rows that
// the epilogue skips (efficiency Hi Wi / (Hp Wp): 0.89 at 32x32).
//
//   * CTA = MB x 128 consecutive positions x NOUT output channels; accumulators (MB x NOUT fp32 columns) in TMEM;
//   * K loop over chunks of 16 input channels, double-buffered: 8 stager warps read the chunk's halo tile from global
//     memory (NCHW fp32, the concat [x1, x2] in place, nearest x2 upsampling folded into the index), apply the
//     GroupNorm scale/shift (+ SiLU) ON THE FLY, split each value into fp16 hi + lo and write the four planes
//     [hi|lo][8-channel k-chunk][position][8] to shared memory; the chunk's packed weights (hi + lo, every tap) arrive by
//     one TMA bulk copy; a ninth warp issues MB x taps x 3 tcgen05.mma (Ahi Whi + Alo Whi + Ahi Wlo: the fp16 x 3
//     split keeps ~22 mantissa bits, i.e. fp32-level parity with the reference, at tensor-pipe speed) and commits to
//     the buffer's "empty" mbarrier;
//   * epilogue: tcgen05.ld, + bias[co] + ebias[b, co] (ResBlock embedding term) + residual, NCHW stores (lanes =
//     consecutive positions -> coalesced); stride 2 (Downsample) keeps every second row / column.
#include <cuda_fp16.h>

#include <algorithm>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

struct ConvTcParams {
  const float* x1; int C1;
  const float* x2; int C2;
  const unsigned char* wimg;  // packed by conv2d_tc_pack_kernel: [ntile][chunk][hi|lo][tap][kc][NOUT][8] fp16
  const float* bias;          // (Cout) or NULL
  const float* ebias;         // (B, Cout) or NULL
  const float* res;           // (B, Cout, Ho, Wo) or NULL
  const float* ss;            // (B, Cin, 2) GroupNorm scale / shift or NULL
  const float* etab;          // 1-D only: (B, Cout, NT) folded embedding channels per tap (NNUnet1D.py:156-175) or NULL
  float* out;
  int silu;                   // apply SiLU after the affine normalisation
  int gelu;                   // 1-D only: exact GELU in the epilogue (ConvBlock1D, NNUnet1D.py:13-33)
  int convt;                  // 1-D only: transposed conv k4 s2 p1 as a 3-tap conv with 2 x convt output columns
                              // (n = parity * convt + co -> out[b, co, 2 c + parity]); convt = real Cout, else 0
  int B, Cout, stride, up, Hs, Ws;  // (Hs, Ws): stored input size; the conv sees (Hs up, Ws up)
  int Hi, Wi, Hp, Wp, Ho, Wo;
  int halo, SL, MB, NC;       // halo = PADH Wp + PADR; SL = 128 MB + 2 halo staged positions; NC = Cin / 16 chunks
  long long total;            // B Hp Wp
  uint32_t mul_img, shr_img, mul_row, shr_row;  // magic numbers: n / (Hp Wp) and n / Wp for n < 2^31
  int tmem_cols;
  int fast;                   // 1: single fp16 product (hi planes only; ~1e-3 relative), 0: three split products (fp32-level)
  const unsigned int* in_amax;  // NULL or max|input| as float bits: the input is staged times 2^k (max -> [2^14, 2^15)) and the
                              // accumulators are scaled back: data gradients of cotangents far below the fp16 range
  TcFlags flags;
};

__device__ __forceinline__ float silu_acc(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

constexpr int CTC_STAGERS = 256;
// staged items (position x 8-channel k-chunk) whose global loads a stager thread keeps in flight together (1, 2, 3 or 6)
#ifndef CTC_ILP
#define CTC_ILP 2
#endif

// power of two that brings max|x| into [2^14, 2^15) (exact scaling, the top of the fp16 range; 1 when no range word is given)
__device__ __forceinline__ float ctc_range_scale(const unsigned int* amax_bits) {
  if (!amax_bits) return 1.0f;
  const float m = __uint_as_float(*amax_bits);
  if (!(m > 0.0f)) return 1.0f;
  int e;
  frexpf(m, &e);
  return ldexpf(1.0f, max(-120, min(120, 15 - e)));
}

// n / d for n < 2^31 with the precomputed (mul, shr) of find_divisor below (d == 1: mul = 0)
__device__ __forceinline__ int fast_div(int n, uint32_t mul, uint32_t shr) {
  return mul ? (int)(__umulhi((uint32_t)n, mul) >> shr) : n;
}

// CONST_BASE: the dynamic shared memory block starts at shared-window address 1024 (probed by the host, verified
// here), so every MMA descriptor is computed from kernel parameters and constants only and ptxas keeps it in uniform
// registers -- no per-MMA R2UR / vote sequences on the issue path.
// NT = taps: 9 (3x3, padding 1), 1 (1x1), 3 (1-D k3 p1), 4 (1-D k4 p1, used with stride 2: taps at -1..+2).
template <int NT>
struct TapGeom {
  static constexpr int PADH = NT == 9 ? 1 : 0;
  static constexpr int PADL = NT == 1 ? 0 : 1;
  static constexpr int PADR = NT == 4 ? 2 : PADL;
};

template <int NOUT, int NT, bool CONST_BASE>
__global__ void __launch_bounds__(CTC_STAGERS + 32, 2) conv2d_tc_kernel(const __grid_constant__ ConvTcParams P) {
  constexpr int PADH = TapGeom<NT>::PADH, PADL = TapGeom<NT>::PADL;
  constexpr int WCHUNK = NT * 2 * 2 * NOUT * 16;  // bytes of packed weights per 16-channel chunk: [hi|lo][tap][kc][NOUT][8]
  const int WSTAGE = P.fast ? WCHUNK / 2 : WCHUNK;  // the single-product mode copies the hi half only
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  // carve: [barriers 128 B][A stage 0][A stage 1][W stage 0][W stage 1], 128-byte aligned
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint64_t* bar_empty = bar_full + 2;                      // [2]
  uint64_t* bar_done = bar_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 5);
  const int PS = P.SL * 16;                                // plane stride (bytes)
  const int ASTAGE = (P.fast ? 2 : 4) * PS;        // planes [hi|lo][k-chunk]; no lo planes in the single-product mode
  unsigned char* sA = smem + 128;
  unsigned char* sW = sA + 2 * ASTAGE;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // broadcast form: lets ptxas treat the role branch as warp-uniform
  const long long p0 = (long long)blockIdx.x * (128 * P.MB);
  const int co0 = blockIdx.y * NOUT;
  const int HpWp = P.Hp * P.Wp;
  const int Cin = P.C1 + P.C2;
  const int HWs = P.Hs * P.Ws;

  if (tid == CTC_STAGERS) {
    mbar_init(bar_full + 0, CTC_STAGERS + 1);
    mbar_init(bar_full + 1, CTC_STAGERS + 1);
    mbar_init(bar_empty + 0, 1);
    mbar_init(bar_empty + 1, 1);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == CTC_STAGERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == CTC_STAGERS / 32) {
    // ================================================ MMA issuer ==================================================
    // Everything on the issue path is warp-uniform by construction (kernel parameters, constants, vote results).
    const uint32_t idesc = umma_idesc_f16(128, NOUT);
    const uint32_t sbase = CONST_BASE ? 1024u : smem_u32(smem);
    if (sbase != smem_u32(smem)) {
      if (lane == 0) tc_raise(P.flags, 2);
    } else {
      const uint32_t a_base0 = sbase + 128u, w_base0 = a_base0 + 2u * (uint32_t)ASTAGE;
      for (int k = 0; k < P.NC; ++k) {
        const int buf = k & 1;
        if (!__all_sync(0xffffffffu, mbar_wait(bar_full + buf, (uint32_t)((k >> 1) & 1), P.flags))) break;
        tc_fence_after();
        const uint32_t a_base = a_base0 + (uint32_t)(buf * ASTAGE);
        const uint32_t w_base = w_base0 + (uint32_t)(buf * WSTAGE);
        for (int mb = 0; mb < P.MB; ++mb) {
          const uint32_t dcol = tbase + (uint32_t)(mb * NOUT);
#pragma unroll
          for (int t = 0; t < NT; ++t) {
            const int toff = NT == 9 ? (t / 3 - 1) * P.Wp + (t % 3 - 1) : (NT == 1 ? 0 : t - 1);
            const uint32_t a_hi = a_base + (uint32_t)((mb * 128 + P.halo + toff) * 16);
            const uint32_t a_lo = a_hi + 2u * (uint32_t)PS;
            const uint32_t w_hi = w_base + (uint32_t)(t * 2 * NOUT * 16);
            const uint32_t w_lo = w_hi + (uint32_t)(NT * 2 * NOUT * 16);
            const uint64_t dAh = umma_desc(a_hi, PS, 128), dAl = umma_desc(a_lo, PS, 128);
            const uint64_t dWh = umma_desc(w_hi, NOUT * 16, 128), dWl = umma_desc(w_lo, NOUT * 16, 128);
            umma_ss(dcol, dAh, dWh, idesc, (k > 0 || t > 0) ? 1u : 0u, 0);
            if (!P.fast) {
              umma_ss(dcol, dAl, dWh, idesc, 1u, 0);
              umma_ss(dcol, dAh, dWl, idesc, 1u, 0);
            }
          }
        }
        umma_commit(bar_empty + buf, 0);
      }
    }
    umma_commit(bar_done, 0);
    __syncwarp();
  } else {
    // ================================================== stagers ===================================================
    bool ok = true;
    // The items a thread stages (position x 8-channel k-chunk) are the same for every 16-channel chunk: decode them once
    // per tile.  it_b = sample index (-1: padding ring / outside, -2: no such item), it_o = element offset of the item's
    // first channel inside the chunk's 16 stored channel planes, it_s = k-chunk << 30 | staged position.
    constexpr int NI_MAX = 6;  // 6 x 256 items = 768 staged positions (the host guarantees SL <= 768)
    const float in_scale = ctc_range_scale(P.in_amax);
    const int nitem = 2 * P.SL, upsh = P.up == 2 ? 1 : 0;
    int it_b[NI_MAX], it_o[NI_MAX], it_s[NI_MAX];
#pragma unroll
    for (int i = 0; i < NI_MAX; ++i) {
      const int e = tid + i * CTC_STAGERS;
      it_b[i] = -2; it_o[i] = 0; it_s[i] = 0;
      if (e < nitem) {
        const int kc = e >= P.SL ? 1 : 0, sp = e - kc * P.SL;
        it_s[i] = (kc << 30) | sp;
        it_b[i] = -1;
        const long long q = p0 - P.halo + sp;
        if (q >= 0 && q < P.total) {
          const int qi = (int)q, bq = fast_div(qi, P.mul_img, P.shr_img), rem = qi - bq * HpWp;
          const int rr = fast_div(rem, P.mul_row, P.shr_row), r = rr - PADH, c = rem - rr * P.Wp - PADL;
          if (r >= 0 && r < P.Hi && c >= 0 && c < P.Wi) {
            it_b[i] = bq;
            it_o[i] = kc * 8 * HWs + (r >> upsh) * P.Ws + (c >> upsh);
          }
        }
      }
    }
    for (int k = 0; k < P.NC && ok; ++k) {
      const int buf = k & 1;
      if (k >= 2) {
        ok = mbar_wait(bar_empty + buf, (uint32_t)(((k >> 1) - 1) & 1), P.flags);
        tc_fence_after();
      }
      unsigned char* wdst = sW + buf * WSTAGE;
      if (tid == 0) {
        mbar_expect_tx(bar_full + buf, (uint32_t)WSTAGE);
        tma_bulk_g2s(wdst, P.wimg + ((size_t)blockIdx.y * P.NC + k) * WCHUNK, (uint32_t)WSTAGE, bar_full + buf);
      }
      unsigned char* adst = sA + buf * ASTAGE;
      // 16 input channels of this chunk come from x1 or from x2 (C1 % 16 == 0): one base pointer + per-sample stride
      const int ch0 = k * 16;
      const bool from1 = ch0 < P.C1;
      const float* xb = from1 ? P.x1 + (size_t)ch0 * HWs : P.x2 + (size_t)(ch0 - P.C1) * HWs;
      const size_t bstride = (size_t)(from1 ? P.C1 : P.C2) * HWs;
      // ILP items (position x 8-channel k-chunk) per step: 8 ILP independent global loads in flight per thread
      constexpr int ILP = CTC_ILP;
#pragma unroll
      for (int i0 = 0; i0 < NI_MAX; i0 += ILP) {
        if (tid + i0 * CTC_STAGERS >= nitem) break;
        float v[ILP][8];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          if (i0 + u < NI_MAX && it_b[i0 + u] >= 0) {
            const float* src = xb + (size_t)it_b[i0 + u] * bstride + it_o[i0 + u];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = __ldg(src + (size_t)j * HWs);
          }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          if (i0 + u >= NI_MAX || it_b[i0 + u] == -2) break;
          const int bcur = it_b[i0 + u], kcc = it_s[i0 + u] >> 30, spp = it_s[i0 + u] & 0x3fffffff;
          uint4 hi4 = make_uint4(0, 0, 0, 0), lo4 = make_uint4(0, 0, 0, 0);
          if (bcur >= 0) {
            if (P.in_amax) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[u][j] *= in_scale;
            }
            if (P.ss) {
              const float4* ssp = reinterpret_cast<const float4*>(P.ss + ((size_t)bcur * Cin + ch0 + kcc * 8) * 2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 a = __ldg(ssp + j);
                v[u][2 * j] = fmaf(v[u][2 * j], a.x, a.y);
                v[u][2 * j + 1] = fmaf(v[u][2 * j + 1], a.z, a.w);
              }
              if (P.silu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = silu_acc(v[u][j]);
              }
            }
            split2_f16(v[u][0], v[u][1], hi4.x, lo4.x);
            split2_f16(v[u][2], v[u][3], hi4.y, lo4.y);
            split2_f16(v[u][4], v[u][5], hi4.z, lo4.z);
            split2_f16(v[u][6], v[u][7], hi4.w, lo4.w);
          }
          *reinterpret_cast<uint4*>(adst + kcc * PS + spp * 16) = hi4;
          if (!P.fast) *reinterpret_cast<uint4*>(adst + (2 + kcc) * PS + spp * 16) = lo4;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + buf);
    }

    // ================================================== epilogue ==================================================
    ok = ok && mbar_wait(bar_done, 0, P.flags);
    tc_fence_after();
    const int q4 = warp & 3, half = warp >> 2;
    constexpr int NH = NOUT / 2;
    const float out_scale = 1.0f / in_scale;
    const int HWo = P.Ho * P.Wo;
    for (int mb = 0; mb < P.MB; ++mb) {
      const long long p = p0 + mb * 128 + q4 * 32 + lane;
      bool valid = ok && p < P.total;
      int b = 0, oy = 0, ox = 0, ccol = 0;
      if (valid) {
        b = fast_div((int)p, P.mul_img, P.shr_img);
        const int rem = (int)p - b * HpWp;
        const int rr = fast_div(rem, P.mul_row, P.shr_row), r = rr - PADH, c = rem - rr * P.Wp - PADL;
        valid = r >= 0 && r < P.Hi && c >= 0 && c < P.Wi;
        ccol = c;
        if (P.stride == 2) {
          valid = valid && !(r & 1) && !(c & 1);
          oy = r >> 1; ox = c >> 1;
          valid = valid && oy < P.Ho && ox < P.Wo;
        } else {
          oy = r; ox = c;
        }
      }
      const uint32_t taddr = tbase + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(mb * NOUT + half * NH);
#pragma unroll 1
      for (int cc = 0; cc < NH; cc += 16) {
        uint32_t rr[16];
        TMEM_LD16(taddr + cc, rr);
        tc_wait_ld();
        if (valid) {
          // 16 consecutive output channels of one position.  Everything that does not depend on the channel is hoisted: one
          // base pointer per tensor and a constant stride per channel, feature switches tested once per 16 values, tap masks
          // as multipliers (fma(1, e, v) = v + e exactly) -- the epilogue was 78 % of the kernel's executed instructions.
          const int cbase = co0 + half * NH + cc;
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
          if (P.in_amax) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] *= out_scale;
          }
          bool done = false;
          if constexpr (NT == 3) {
            if (P.convt) {  // ConvTranspose1d: even / odd outputs are the two halves of the N dimension (convt % 16 == 0)
              const int par = cbase >= P.convt ? 1 : 0, cb = cbase - par * P.convt;
              if (P.bias) {
                const float* bp = P.bias + cb;
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += __ldg(bp + j);
              }
              float* op = P.out + ((size_t)b * P.convt + cb) * P.Wo + 2 * ccol + par;
#pragma unroll
              for (int j = 0; j < 16; ++j) op[(size_t)j * P.Wo] = v[j];
              done = true;
            }
          }
          if (!done) {
            if (P.bias) {
              const float* bp = P.bias + cbase;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(bp + j);
            }
            if (P.ebias) {
              const float* ep = P.ebias + (size_t)b * P.Cout + cbase;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(ep + j);
            }
            const size_t o = ((size_t)b * P.Cout + cbase) * HWo + oy * P.Wo + ox;
            if (P.res) {
              const float* rp = P.res + o;
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += __ldg(rp + (size_t)j * HWo);
            }
            if constexpr (NT == 3 || NT == 4) {
              if (P.etab) {  // embedding channels are constant along the signal: a tap contributes where it reads inside it
                const float* et = P.etab + ((size_t)b * P.Cout + cbase) * NT;
                float m[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) m[t] = (ccol + t - 1 >= 0 && ccol + t - 1 < P.Wi) ? 1.0f : 0.0f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
#pragma unroll
                  for (int t = 0; t < NT; ++t) v[j] = fmaf(m[t], __ldg(et + j * NT + t), v[j]);
                }
              }
              if (P.gelu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752440f));
              }
            }
            float* op = P.out + o;
#pragma unroll
            for (int j = 0; j < 16; ++j) op[(size_t)j * HWo] = v[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CTC_STAGERS / 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(P.tmem_cols));
}

// ---- weight packing: (Cout, Cin, K, K) fp32 -> [ntile][chunk][hi|lo][tap][kc][NOUT][8] fp16 ---------------------------
// dgrad = 1: W is the FORWARD conv's weight (Cin rows of the image come from its output axis): the image is the one of the
// data-gradient conv, whose weight is W with the channel roles swapped and the taps flipped, Wd[ci_f][co_f][KK-1-t] -- so that
// no transposed / flipped copy of the weight has to be made first.  Cw is always the row length of W's input axis.
__global__ void conv2d_tc_pack_kernel(const float* __restrict__ W, int Cout, int Cw, int Cin, int KK, int NOUT,
                                      __half* __restrict__ img, long long nel, int dgrad) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  const int j = (int)(e % 8);
  long long r = e / 8;
  const int n = (int)(r % NOUT); r /= NOUT;
  const int kc = (int)(r % 2); r /= 2;
  const int t = (int)(r % KK); r /= KK;
  const int hl = (int)(r % 2); r /= 2;
  const int NC = Cin / 16;
  const int k = (int)(r % NC);
  const int nt = (int)(r / NC);
  const int co = nt * NOUT + n, ci = k * 16 + kc * 8 + j;
  // Cw >= Cin: the weight tensor may carry extra (folded) input channels
  const float v = dgrad ? W[((size_t)ci * Cw + co) * KK + (KK - 1 - t)] : W[((size_t)co * Cw + ci) * KK + t];
  const __half hi = __float2half_rn(v);
  img[e] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
}

// ConvTranspose1d(Cin, Cout, k4, s2, p1), W (Cin, Cout, 4): out[2m] = x[m] W1 + x[m-1] W3, out[2m+1] = x[m+1] W0 + x[m] W2
// -> a 3-tap conv (offsets -1, 0, +1) with 2 Cout output columns n = parity Cout + co; one third of the image is zeros.
__global__ void convt1d_tc_pack_kernel(const float* __restrict__ W, int Cout, int Cin, int NOUT, __half* __restrict__ img,
                                       long long nel) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nel) return;
  const int j = (int)(e % 8);
  long long r = e / 8;
  const int n = (int)(r % NOUT); r /= NOUT;
  const int kc = (int)(r % 2); r /= 2;
  const int t = (int)(r % 3); r /= 3;
  const int hl = (int)(r % 2); r /= 2;
  const int NC = Cin / 16;
  const int k = (int)(r % NC);
  const int nt = (int)(r / NC);
  const int nn = nt * NOUT + n, par = nn >= Cout ? 1 : 0, co = nn - par * Cout, ci = k * 16 + kc * 8 + j;
  const int wk = par == 0 ? (t == 0 ? 3 : (t == 1 ? 1 : -1)) : (t == 0 ? -1 : (t == 1 ? 2 : 0));
  const float v = wk >= 0 ? W[((size_t)ci * Cout + co) * 4 + wk] : 0.0f;
  const __half hi = __float2half_rn(v);
  img[e] = hl ? __float2half_rn(v - __half2float(hi)) : hi;
}

// ---- GroupNorm as a per-(sample, channel) affine map: ss[b, c] = (rstd gamma_c, beta_c - mean rstd gamma_c) -------------
// (GroupNorm32, model/nn_utils.py:39-41,107-114; statistics over channels [g cpg, (g+1) cpg) x HW of the concat [x1, x2])
// One warp per (sample, group): 16-byte loads when HW % 4 == 0, shuffle reduction, no block barrier.
__global__ void __launch_bounds__(256) gn_scale_shift_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                             int C2, int HW, int G, int BG, float eps,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             float* __restrict__ ss) {
  const int wg = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (wg >= BG) return;
  const int b = wg / G, g = wg % G, C = C1 + C2, cpg = C / G;
  float s = 0.0f, q = 0.0f;
  for (int ci = 0; ci < cpg; ++ci) {
    const int c = g * cpg + ci;
    const float* src = c < C1 ? x1 + ((size_t)b * C1 + c) * HW : x2 + ((size_t)b * C2 + (c - C1)) * HW;
    if ((HW & 3) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int e = lane; e < HW / 4; e += 32) {
        const float4 v = __ldg(s4 + e);
        s += (v.x + v.y) + (v.z + v.w);
        q = fmaf(v.x, v.x, q); q = fmaf(v.y, v.y, q); q = fmaf(v.z, v.z, q); q = fmaf(v.w, v.w, q);
      }
    } else {
      for (int e = lane; e < HW; e += 32) {
        const float v = __ldg(src + e);
        s += v;
        q = fmaf(v, v, q);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
  const float n = (float)cpg * (float)HW;
  const float mean = s / n, var = fmaxf(q / n - mean * mean, 0.0f), rstd = rsqrtf(var + eps);
  for (int ci = lane; ci < cpg; ci += 32) {
    const int c = g * cpg + ci;
    const float sc = rstd * gamma[c];
    ss[((size_t)b * C + c) * 2] = sc;
    ss[((size_t)b * C + c) * 2 + 1] = fmaf(-mean, sc, beta[c]);
  }
}

// ---- host dispatch ---------------------------------------------------------------------------------------------------
// output-channel tile: 192 serves the attention qkv convs (Cout = 3 C = 192 / 384) with one / two tiles instead of three
// output-channel tile: 192 only where two weight stages of that width fit shared memory (up to 4 taps: 1x1 and the 1-D convs);
// a 3x3 conv with 192 k output channels (data gradients of the 192-input decoder convs) runs as 128- or 64-wide tiles
int conv2d_tc_nout(int Cout, int taps) {
  return (Cout % 192 == 0 && taps <= 4) ? 192 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32));
}

size_t conv2d_tc_pack_bytes(int Cout, int Cin, int KK) { return (size_t)Cout * Cin * KK * 2 * 2; }

// W is (Cout, Cw, taps) with the first Cin input channels packed (Cw > Cin: trailing channels are handled elsewhere).
int conv2d_tc_pack(msgm_ctx* ctx, const float* W, int Cout, int Cw, int Cin, int KK, void* img, cudaStream_t stream, int dgrad) {
  const long long nel = (long long)Cout * Cin * KK * 2;
  conv2d_tc_pack_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, stream>>>(W, Cout, Cw, Cin, KK, conv2d_tc_nout(Cout, KK),
                                                                          reinterpret_cast<__half*>(img), nel, dgrad);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int gn_scale_shift(msgm_ctx* ctx, const float* x1, int C1, const float* x2, int C2, int HW, int G, int B, const float* gamma,
                   const float* beta, float* ss, cudaStream_t stream) {
  if (B == 0) return MSGM_OK;
  gn_scale_shift_kernel<<<(B * G + 7) / 8, 256, 0, stream>>>(x1, C1, x2, x2 ? C2 : 0, HW, G, B * G, 1e-5f, gamma, beta, ss);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

// CUTLASS-style magic division: n / d == umulhi(n, mul) >> shr for 0 <= n < 2^31
static void find_divisor(uint32_t d, uint32_t* mul, uint32_t* shr) {
  if (d <= 1) { *mul = 0; *shr = 0; return; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const uint32_t p = 31 + l;
  *mul = (uint32_t)(((1ull << p) + d - 1) / d);
  *shr = p - 32;
}

template <int NOUT, int NT>
static int launch_conv_tc(msgm_ctx* ctx, ConvTcParams& P, cudaStream_t stream) {
  const int WSTAGE = NT * (P.fast ? 1 : 2) * 2 * NOUT * 16;
  const int nplane = P.fast ? 2 : 4;
  P.Hp = P.Hi + 2 * TapGeom<NT>::PADH;
  P.Wp = P.Wi + TapGeom<NT>::PADL + TapGeom<NT>::PADR;
  P.halo = TapGeom<NT>::PADH * P.Wp + TapGeom<NT>::PADR;
  P.total = (long long)P.B * P.Hp * P.Wp;
  find_divisor((uint32_t)(P.Hp * P.Wp), &P.mul_img, &P.shr_img);
  find_divisor((uint32_t)P.Wp, &P.mul_row, &P.shr_row);
  if (P.total >= (1LL << 31) - 4096) {
    set_error("msgm_conv_tc: more than 2^31 padded positions in one call; split the batch");
    return MSGM_ERR_UNSUPPORTED;
  }
  // M blocks (128 positions) per CTA, 1..4 (TMEM: MB x NOUT <= 512 columns): the choice that minimises
  // waves x (MB + 1.5 blocks of per-tile overhead), e.g. 324 blocks on 148 SMs run as one wave of 108 three-block tiles rather than
  // two waves of two-block tiles.  A wave holds num_sms x (CTAs that fit one SM's shared memory / register file).
  const long long nblk = (P.total + 127) / 128;
  const int reg_limit = 2;
  int MB = 0;
  size_t smem = 0;
  double best = 1e300;
  for (int mb = 1; mb <= 4; ++mb) {
    const int SL = 128 * mb + 2 * P.halo;
    const size_t sm = 128 + 128 + 2 * (size_t)(nplane * SL * 16) + 2 * (size_t)WSTAGE;
    if (sm > 227 * 1024 || SL > 768 || mb * NOUT > 512) continue;  // shared memory, item table, TMEM columns
    const long long tiles = ((nblk + mb - 1) / mb) * (P.Cout / NOUT);
    const int occ = (int)std::max<size_t>(1, std::min<size_t>(reg_limit, (size_t)(228 * 1024) / (sm + 1024)));
    const long long slots = (long long)ctx->num_sms * occ;
    // few waves: whole waves count (the tail wave costs a full tile time); many waves: CTAs drift apart and the
    // scheduler fills the gaps, so the fractional count is the better model
    const double wexact = (double)tiles / (double)slots;
    const double waves = wexact >= 4.0 ? wexact : (double)((tiles + slots - 1) / slots);
    const double cost = waves * (mb + 1.5);  // 1.5 blocks: halo staging (2 halo / 128), pipeline fill, prologue, epilogue
    if (cost <= best) { best = cost; MB = mb; smem = sm; }
  }
  if (MB == 0) {
    set_error("msgm_conv_tc: tile does not fit shared memory / the stager's item table (image too wide)");
    return MSGM_ERR_UNSUPPORTED;
  }
  P.MB = MB;
  P.SL = 128 * MB + 2 * P.halo;
  int cols = 32;
  while (cols < MB * NOUT) cols <<= 1;
  P.tmem_cols = cols;
  uint32_t sb = 0;
  int rc = dyn_smem_base(ctx, stream, &sb);
  if (rc) return rc;
  auto kern = sb == 1024u ? conv2d_tc_kernel<NOUT, NT, true> : conv2d_tc_kernel<NOUT, NT, false>;
  MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((nblk + MB - 1) / MB), (unsigned)(P.Cout / NOUT));
  kern<<<grid, CTC_STAGERS + 32, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int NT>
static int launch_conv_tc_n(msgm_ctx* ctx, ConvTcParams& P, cudaStream_t stream) {
  const int nout = conv2d_tc_nout(P.Cout, NT);
  if (nout == 192) return launch_conv_tc<192, NT>(ctx, P, stream);
  if (nout == 128) return launch_conv_tc<128, NT>(ctx, P, stream);
  if (nout == 64) return launch_conv_tc<64, NT>(ctx, P, stream);
  return launch_conv_tc<32, NT>(ctx, P, stream);
}

int conv2d_tc(msgm_ctx* ctx, const msgm_conv2d_tc_desc* D, cudaStream_t stream) {
  if (D->B == 0) return MSGM_OK;
  ConvTcParams P{};
  P.x1 = D->x1; P.C1 = D->C1; P.x2 = D->x2; P.C2 = D->x2 ? D->C2 : 0;
  P.wimg = reinterpret_cast<const unsigned char*>(D->wimg);
  P.bias = D->bias; P.ebias = D->ebias; P.res = D->res; P.ss = D->ss; P.out = D->out;
  P.silu = D->prologue == 2;
  P.fast = D->fast ? 1 : 0;
  P.B = D->B; P.Cout = D->Cout; P.stride = D->stride; P.up = D->up; P.Hs = D->Hs; P.Ws = D->Ws;
  const int pad = D->K / 2;
  P.Hi = D->Hs * D->up; P.Wi = D->Ws * D->up;
  P.Ho = (P.Hi + 2 * pad - D->K) / D->stride + 1;
  P.Wo = (P.Wi + 2 * pad - D->K) / D->stride + 1;
  P.NC = (P.C1 + P.C2) / 16;
  P.in_amax = ctx->tc_in_amax;
  ctx->tc_in_amax = nullptr;
  P.flags = next_tc_flags(ctx);
  return D->K == 3 ? launch_conv_tc_n<9>(ctx, P, stream) : launch_conv_tc_n<1>(ctx, P, stream);
}

// nn.Conv1d k3 (stride 1) / k4 (stride 2) / k1, padding 1 (0 for k1), as the H = 1 case of the same kernel
int conv1d_tc(msgm_ctx* ctx, const msgm_conv1d_tc_desc* D, cudaStream_t stream) {
  if (D->B == 0) return MSGM_OK;
  ConvTcParams P{};
  P.x1 = D->x1; P.C1 = D->C1; P.x2 = D->x2; P.C2 = D->x2 ? D->C2 : 0;
  P.wimg = reinterpret_cast<const unsigned char*>(D->wimg);
  P.bias = D->bias; P.etab = D->E; P.out = D->out; P.gelu = D->gelu;
  P.fast = D->fast ? 1 : 0;
  P.B = D->B; P.Cout = D->Cout; P.stride = D->stride; P.up = 1; P.Hs = 1; P.Ws = D->Lin;
  P.Hi = 1; P.Wi = D->Lin; P.Ho = 1;
  const int pad = D->K == 1 ? 0 : 1;
  P.Wo = (D->Lin + 2 * pad - D->K) / D->stride + 1;
  P.NC = (P.C1 + P.C2) / 16;
  P.in_amax = ctx->tc_in_amax;
  ctx->tc_in_amax = nullptr;
  P.flags = next_tc_flags(ctx);
  if (D->K == 3) return launch_conv_tc_n<3>(ctx, P, stream);
  if (D->K == 4) return launch_conv_tc_n<4>(ctx, P, stream);
  return launch_conv_tc_n<1>(ctx, P, stream);
}

int convt1d_tc_pack(msgm_ctx* ctx, const float* W, int Cout, int Cin, void* img, cudaStream_t stream) {
  const long long nel = (long long)2 * Cout * Cin * 3 * 2;
  convt1d_tc_pack_kernel<<<(unsigned)((nel + 255) / 256), 256, 0, stream>>>(W, Cout, Cin, conv2d_tc_nout(2 * Cout, 3),
                                                                           reinterpret_cast<__half*>(img), nel);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

// out (B, Cout, Lout) must be zero beyond 2 Lin (the reference right-pads with zeros, NNUnet1D.py:165-169)
int convt1d_tc(msgm_ctx* ctx, const float* x, const void* wimg, const float* bias, float* out, int B, int Cin, int Cout,
               int Lin, int Lout, int fast, cudaStream_t stream) {
  if (B == 0) return MSGM_OK;
  ConvTcParams P{};
  P.x1 = x; P.C1 = Cin;
  P.wimg = reinterpret_cast<const unsigned char*>(wimg);
  P.bias = bias; P.out = out; P.convt = Cout; P.fast = fast ? 1 : 0;
  P.B = B; P.Cout = 2 * Cout; P.stride = 1; P.up = 1; P.Hs = 1; P.Ws = Lin;
  P.Hi = 1; P.Wi = Lin; P.Ho = 1; P.Wo = Lout;
  P.NC = Cin / 16;
  P.in_amax = ctx->tc_in_amax;
  ctx->tc_in_amax = nullptr;
  P.flags = next_tc_flags(ctx);
  return launch_conv_tc_n<3>(ctx, P, stream);
}

}  // namespace msgm
