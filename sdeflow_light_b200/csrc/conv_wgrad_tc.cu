// Weight gradient of the U-Net convolutions on tcgen05 (training path of NNUnet1D.py:110-179 / model/unet.py:101-250 under
// PluginReverseSDE.ssm_loss, SDEs.py:616-646):
//     gW[co][ci][ky][kx] = sum_{n,y,x} cot[n][co][y][x] in[n][ci][s y + ky - p][s x + kx - p]      (stride s = 1 or 2)
// is a product over POSITIONS: D[co, (tap, ci)] = Cot^T In_tap with K = all positions of all 2B samples of the primal /
// tangent pair.  Both tensors are NCHW, i.e. the contraction index is the contiguous one, and the staged tile layout of the
// forward conv (conv2d_tc.cu: planes [8-channel chunk][position][8 channels], fp16 hi + lo) read MN-MAJOR is exactly that
// product's operand layout (validated for ssm_tc.cu by tools/mn_probe.cu: LBO = 128 B between 8-position groups, SBO = plane
// chunk stride between 8-channel groups).  As in the forward conv, all images live in one zero-padded LINEAR position space
// p = (n Hp + r) Wp + c, so a tap is a shift of the input operand's start address by whole 16-byte rows and no im2col exists.
//
//   * CTA = 128 output channels (M, zero-padded) x NCI input channels (N) x the NTG taps of one kernel row (3x3: grid.y also
//     enumerates the kernel rows, whose shift (ky - 1) Wp is folded into the staged input window, so its halo is one position)
//     x a contiguous slice of the positions; accumulators NTG x NCI fp32 TMEM columns;
//   * position chunks of 64 (4 K = 16 slices), double-buffered: 8 stager warps read cot and the input window from global
//     memory (lanes = consecutive positions: coalesced), scale the cotangent by a power of two into the fp16 range (deep-layer
//     cotangents are ~1e-7: the fp16 x 3 split would lose its low part in the subnormals), split to fp16 hi + lo and write the
//     planes; a ninth warp issues 4 x NTG x 3 tcgen05.mma (hi hi + lo hi + hi lo: fp32-level parity) and commits;
//   * epilogue: tcgen05.ld, unscale, the CTA's partial tile goes to a scratch slice with coalesced plain stores (lane = output
//     channel); wgrad_reduce_kernel sums the position slices into gW in a fixed order (no atomics: deterministic).
#include <cuda_fp16.h>

#include <algorithm>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

constexpr int WG_KT = 64;          // positions per staged chunk
constexpr int WG_STAGERS = 256;
constexpr uint32_t WG_IDESC_MN = (1u << 15) | (1u << 16);  // both operands MN-major ("transposed")

// Tap structure of one CTA.  MODE 0: 1x1.  MODE 1: the three taps of a kernel row, stride 1 (3x3 pad 1: grid.y enumerates the
// rows; 1-D k3 pad 1).  MODE 2: 1-D k4 stride 2 pad 1 (also the weight gradient of ConvTranspose1d(k4, s2, p1) with the roles
// of the two tensors swapped).  MODE 3: the three taps of a kernel row of a 3x3 stride-2 pad-1 conv.
// With stride 2 consecutive cotangent positions read every second input column, so the input window is staged as two PHASE
// windows, E[x] = in[2x] and O[x] = in[2x + 1]; tap kx reads in[2x - 1 + kx] = O[x-1], E[x], O[x], E[x+1].  A window row r holds
// cotangent position (chunk start + r - HALO); a tap is (phase window, row offset).
template <int MODE>
struct WgTaps {
  static constexpr int NTG = MODE == 0 ? 1 : (MODE == 2 ? 4 : 3);
  static constexpr int NPH = MODE >= 2 ? 2 : 1;
  static constexpr int HALO = MODE == 0 ? 0 : 1;
  static constexpr int STRIDE = MODE >= 2 ? 2 : 1;
  __host__ __device__ static constexpr int phase(int t) { return MODE >= 2 ? ((t & 1) ? 0 : 1) : 0; }
  __host__ __device__ static constexpr int off(int t) { return MODE >= 2 ? (t + 1) / 2 : t; }
};

struct WgradTcParams {
  const float* cot;  // (N, Cout, Ho, Wo)
  const float* x1;   // (N, C1, Hs, Ws)
  const float* x2;   // (N, C2, Hs, Ws) or NULL: channel concat
  float* gW;         // (Cout, Cw, KH, KW); input channels [coff, coff + C1 + C2)
  float* scratch;    // per-CTA partial tiles [cta][tap * NCI + ci][128 co] (plain coalesced stores; reduced by wgrad_reduce_kernel)
  const unsigned int* amax_bits;  // max |cot| as float bits (msgm_amax) or NULL: no range scaling
  const unsigned int* amax_in_bits;  // the same for the input operand (ConvTranspose: the "input" is the cotangent)
  int C1, C2, Cout, Cw, coff, KH, KW;
  int N, Ho, Wo, Wpo;         // cotangent grid; Wpo = Wo + 2 HALO: one zero ring column on either side of every row
  int Hin, Win, Hs, Ws, upsh; // the conv sees the input as (Hin, Win) = (Hs << upsh, Ws << upsh)
  int pad_h;                  // rows: input row = STRIDE * y + ky - pad_h
  long long total;            // N Ho Wpo padded cotangent positions
  int nchunks, chunks_per_cta;
  int NCI, NG;                // input-channel tile width; kernel rows enumerated by grid.y (KH)
  uint32_t mul_img, shr_img, mul_row, shr_row;
  int tmem_cols;
  TcFlags flags;
};

__device__ __forceinline__ int wg_div(int n, uint32_t mul, uint32_t shr) { return mul ? (int)(__umulhi((uint32_t)n, mul) >> shr) : n; }

__device__ __forceinline__ float wg_pow2(const unsigned int* amax_bits, int target_exp) {
  if (!amax_bits) return 1.0f;
  const float m = __uint_as_float(*amax_bits);
  if (!(m > 0.0f)) return 1.0f;
  int e;
  frexpf(m, &e);
  return ldexpf(1.0f, max(-120, min(120, target_exp - e)));
}

template <int MODE, bool CONST_BASE>
__global__ void __launch_bounds__(WG_STAGERS + 32, 1) conv_wgrad_tc_kernel(const __grid_constant__ WgradTcParams P) {
  using TP = WgTaps<MODE>;
  constexpr int KT = WG_KT, NTG = TP::NTG, NPH = TP::NPH, HALO = TP::HALO, SLI = KT + 2 * HALO;
  constexpr int APLANE = 16 * KT * 16;  // 128 channels x KT positions, fp16
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);  // [2]
  uint64_t* bar_empty = bar_full + 2;                      // [2]
  uint64_t* bar_done = bar_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 5);
  const int BPLANE = (P.NCI / 8) * SLI * 16;               // one plane of one phase window
  const int ASTAGE = 2 * APLANE, BSTAGE = NPH * 2 * BPLANE;  // [hi|lo] / [phase][hi|lo]
  unsigned char* sA = smem + 128;
  unsigned char* sB = sA + 2 * ASTAGE;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int g = blockIdx.y % P.NG, cit = blockIdx.y / P.NG;  // g = kernel row ky
  const int co0 = blockIdx.z * 128, ci0 = cit * P.NCI;
  const int Cin = P.C1 + P.C2;
  const int first_chunk = blockIdx.x * P.chunks_per_cta;
  const int my_chunks = min(P.chunks_per_cta, P.nchunks - first_chunk);
  const int HoWpo = P.Ho * P.Wpo;

  if (tid == WG_STAGERS) {
    mbar_init(bar_full + 0, WG_STAGERS);
    mbar_init(bar_full + 1, WG_STAGERS);
    mbar_init(bar_empty + 0, 1);
    mbar_init(bar_empty + 1, 1);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_STAGERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // channel chunks the stagers never write (Cout < 128, input channels past Cin) must read as zeros
  const int nch_a = min(16, (P.Cout - co0 + 7) / 8);
  const int nch_b = min(P.NCI / 8, max(0, (Cin - ci0) / 8));
  if (nch_a < 16 || nch_b < P.NCI / 8) {
    uint4* z = reinterpret_cast<uint4*>(sA);
    const int n16 = (2 * ASTAGE + 2 * BSTAGE) / 16;
    for (int e = tid; e < n16; e += WG_STAGERS + 32) z[e] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == WG_STAGERS / 32) {
    // ================================================ MMA issuer ==================================================
    const uint32_t idesc = umma_idesc_f16(128, P.NCI) | WG_IDESC_MN;
    const uint32_t sbase = CONST_BASE ? 1024u : smem_u32(smem);
    if (sbase != smem_u32(smem)) {
      if (lane == 0) tc_raise(P.flags, 2);
    } else {
      const uint32_t a_base0 = sbase + 128u, b_base0 = a_base0 + 2u * (uint32_t)ASTAGE;
      for (int k = 0; k < my_chunks; ++k) {
        const int buf = k & 1;
        if (!__all_sync(0xffffffffu, mbar_wait(bar_full + buf, (uint32_t)((k >> 1) & 1), P.flags))) break;
        tc_fence_after();
        const uint32_t a_base = a_base0 + (uint32_t)(buf * ASTAGE), b_base = b_base0 + (uint32_t)(buf * BSTAGE);
#pragma unroll
        for (int s = 0; s < KT / 16; ++s) {
          // MN-major descriptors: 8-position groups 128 B apart (LBO), 8-channel groups one plane chunk apart (SBO)
          const uint64_t dAh = umma_desc(a_base + (uint32_t)(s * 256), 128, KT * 16);
          const uint64_t dAl = umma_desc(a_base + (uint32_t)(APLANE + s * 256), 128, KT * 16);
#pragma unroll
          for (int t = 0; t < NTG; ++t) {
            const uint32_t b_hi = b_base + (uint32_t)(TP::phase(t) * 2 * BPLANE) + (uint32_t)((TP::off(t) + 16 * s) * 16);
            const uint64_t dBh = umma_desc(b_hi, 128, SLI * 16), dBl = umma_desc(b_hi + (uint32_t)BPLANE, 128, SLI * 16);
            const uint32_t dcol = tbase + (uint32_t)(t * P.NCI);
            umma_ss(dcol, dAh, dBh, idesc, (k > 0 || s > 0) ? 1u : 0u, 0);
            umma_ss(dcol, dAl, dBh, idesc, 1u, 0);
            umma_ss(dcol, dAh, dBl, idesc, 1u, 0);
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
    const float scale = wg_pow2(P.amax_bits, 15), scale_in = wg_pow2(P.amax_in_bits, 15);
    const long long HWo = (long long)P.Ho * P.Wo, HWs = (long long)P.Hs * P.Ws;
    for (int k = 0; k < my_chunks && ok; ++k) {
      const int buf = k & 1;
      if (k >= 2) {
        ok = mbar_wait(bar_empty + buf, (uint32_t)(((k >> 1) - 1) & 1), P.flags);
        tc_fence_after();
      }
      const long long pbase = (long long)(first_chunk + k) * KT;
      unsigned char* adst = sA + buf * ASTAGE;
      unsigned char* bdst = sB + buf * BSTAGE;
      // Items = (8-channel chunk, position).  A thread's items are handled ILP at a time: all their global loads are issued
      // before the first is consumed (one round trip per batch instead of one per item).
      constexpr int ILP = 5;
      // ---- cotangent: nch_a channel chunks x KT positions ----
      for (int e0 = 0; e0 < nch_a * KT; e0 += ILP * WG_STAGERS) {
        float v[ILP][8];
        int dst[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          const int e = e0 + tid + u * WG_STAGERS;
          dst[u] = -1;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.0f;
          if (e < nch_a * KT) {
            const int row = e & (KT - 1), c = e / KT;
            dst[u] = (c * KT + row) * 16;
            const long long q = pbase + row;
            if (q < P.total) {
              const int qi = (int)q, n = wg_div(qi, P.mul_img, P.shr_img), rem = qi - n * HoWpo;
              const int y = wg_div(rem, P.mul_row, P.shr_row), x = rem - y * P.Wpo - HALO;
              if (x >= 0 && x < P.Wo) {
                const int ch = co0 + 8 * c;
                const float* src = P.cot + ((long long)n * P.Cout + ch) * HWo + (long long)y * P.Wo + x;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  if (ch + j < P.Cout) v[u][j] = __ldg(src + j * HWo);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          if (dst[u] >= 0) {
            uint4 hi4, lo4;
            split2_f16(v[u][0] * scale, v[u][1] * scale, hi4.x, lo4.x);
            split2_f16(v[u][2] * scale, v[u][3] * scale, hi4.y, lo4.y);
            split2_f16(v[u][4] * scale, v[u][5] * scale, hi4.z, lo4.z);
            split2_f16(v[u][6] * scale, v[u][7] * scale, hi4.w, lo4.w);
            *reinterpret_cast<uint4*>(adst + dst[u]) = hi4;
            *reinterpret_cast<uint4*>(adst + APLANE + dst[u]) = lo4;
          }
        }
      }
      // ---- input: NPH phase windows x nch_b channel chunks x SLI positions ----
      const int nb = NPH * nch_b * SLI;
      for (int e0 = 0; e0 < nb; e0 += ILP * WG_STAGERS) {
        float v[ILP][8];
        int dst[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          const int e = e0 + tid + u * WG_STAGERS;
          dst[u] = -1;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.0f;
          if (e < nb) {
            const int ph = NPH == 2 ? (e >= nch_b * SLI ? 1 : 0) : 0;
            const int e1 = e - ph * nch_b * SLI, c = e1 / SLI, row = e1 - c * SLI;
            dst[u] = ph * 2 * BPLANE + (c * SLI + row) * 16;
            const long long q = pbase + row - HALO;
            if (q >= 0 && q < P.total) {
              const int qi = (int)q, n = wg_div(qi, P.mul_img, P.shr_img), rem = qi - n * HoWpo;
              const int y = wg_div(rem, P.mul_row, P.shr_row), x = rem - y * P.Wpo - HALO;
              // ring columns (x = -1, Wo) map outside the input for every tap that reads them
              const int iy = TP::STRIDE * y + g - P.pad_h, ix = TP::STRIDE * x + ph;
              if (x >= -1 && x <= P.Wo && iy >= 0 && iy < P.Hin && ix >= 0 && ix < P.Win) {
                const int ch = ci0 + 8 * c;  // C1 % 8 == 0: a chunk never straddles the concat
                const long long off = (long long)(iy >> P.upsh) * P.Ws + (ix >> P.upsh);
                const float* src = ch < P.C1 ? P.x1 + ((long long)n * P.C1 + ch) * HWs + off
                                             : P.x2 + ((long long)n * P.C2 + ch - P.C1) * HWs + off;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = __ldg(src + j * HWs);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
          if (dst[u] >= 0) {
            uint4 hi4, lo4;
            split2_f16(v[u][0] * scale_in, v[u][1] * scale_in, hi4.x, lo4.x);
            split2_f16(v[u][2] * scale_in, v[u][3] * scale_in, hi4.y, lo4.y);
            split2_f16(v[u][4] * scale_in, v[u][5] * scale_in, hi4.z, lo4.z);
            split2_f16(v[u][6] * scale_in, v[u][7] * scale_in, hi4.w, lo4.w);
            *reinterpret_cast<uint4*>(bdst + dst[u]) = hi4;
            *reinterpret_cast<uint4*>(bdst + BPLANE + dst[u]) = lo4;
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(bar_full + buf);
    }

    // ================================================== epilogue ==================================================
    ok = ok && mbar_wait(bar_done, 0, P.flags);
    tc_fence_after();
    if (warp < 4) {
      // partial tile -> scratch: lane = output channel, so a warp store covers 128 contiguous bytes; no atomics, fixed order
      const float inv = 1.0f / (scale * scale_in);
      const size_t cta = ((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      float* dst = P.scratch + cta * (size_t)(NTG * P.NCI) * 128 + warp * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < NTG * P.NCI; cc += 16) {
        uint32_t rr[16];
        TMEM_LD16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)cc, rr);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[(size_t)(cc + j) * 128] = (ok && my_chunks > 0) ? __uint_as_float(rr[j]) * inv : 0.0f;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WG_STAGERS / 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(P.tmem_cols));
}

// gW[co][coff + ci][ky][kx] (+)= sum over the position slices of the partial tiles (threads run along co: coalesced reads)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ gW, int gx, int gy,
                                                           int gz, int cols, int NCI, int NG, int Cout, int Cin, int Cw, int coff,
                                                           int KH, int KW, int accumulate) {
  const long long total = (long long)gz * gy * cols * 128;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i & 127);
    long long r = i >> 7;
    const int col = (int)(r % cols); r /= cols;
    const int y = (int)(r % gy), z = (int)(r / gy);
    const int co = z * 128 + row, g = y % NG, cit = y / NG, t = col / NCI, ci = cit * NCI + col % NCI;
    if (co >= Cout || ci >= Cin) continue;
    const float* src = scratch + (((size_t)z * gy + y) * gx) * (size_t)cols * 128 + (size_t)col * 128 + row;
    float sum = 0.0f;
    for (int x = 0; x < gx; ++x) sum += src[(size_t)x * cols * 128];
    float* o = gW + (((size_t)co * Cw + coff + ci) * KH + g) * KW + t;
    *o = accumulate ? *o + sum : sum;
  }
}

static void wg_find_divisor(uint32_t d, uint32_t* mul, uint32_t* shr) {
  if (d <= 1) { *mul = 0; *shr = 0; return; }
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const uint32_t p = 31 + l;
  *mul = (uint32_t)(((1ull << p) + d - 1) / d);
  *shr = p - 32;
}

// tap structure of a convolution, or -1 when the tensor-core kernel does not take it
static int wg_mode(int KH, int KW, int stride, int pad, int up, int Hs, int Ws) {
  if (stride == 1 && (up == 1 || up == 2)) {
    if (KH == 1 && KW == 1 && pad == 0) return 0;
    if ((KH == 3 || KH == 1) && KW == 3 && pad == 1) return 1;
  }
  if (stride == 2 && up == 1 && pad == 1) {
    if (KH == 1 && KW == 4 && Ws % 2 == 0) return 2;
    if (KH == 3 && KW == 3 && Hs % 2 == 0 && Ws % 2 == 0) return 3;
  }
  return -1;
}

struct WgGeom {
  int mode, ntg, nph, halo, NG, nci, ci_tiles, co_tiles, nchunks, chunks_per_cta, slices, cols, Ho, Wo, Wpo;
  long long total;
};

static WgGeom wg_geometry(const msgm_ctx* ctx, int N, int Cout, int Cin, int KH, int KW, int stride, int pad, int up, int Hs, int Ws) {
  WgGeom g{};
  g.mode = wg_mode(KH, KW, stride, pad, up, Hs, Ws);
  g.ntg = g.mode == 0 ? 1 : (g.mode == 2 ? 4 : 3);
  g.nph = g.mode >= 2 ? 2 : 1;
  g.halo = g.mode == 0 ? 0 : 1;
  g.NG = KH;
  g.Ho = stride == 2 ? (KH == 1 ? 1 : Hs / 2) : Hs * up;
  g.Wo = stride == 2 ? Ws / 2 : Ws * up;
  g.Wpo = g.Wo + 2 * g.halo;
  g.total = (long long)N * g.Ho * g.Wpo;
  // input-channel tile: as wide as the TMEM columns (ntg x NCI <= 512) and the channel count allow, a multiple of 16
  int nci = std::min(128, Cin);
  if (Cin > 128) {  // balance the tiles: 192 -> 2 x 96, 256 -> 2 x 128, 384 -> 3 x 128
    const int tiles = (Cin + 127) / 128;
    nci = ((Cin + tiles - 1) / tiles + 15) / 16 * 16;
  }
  g.nci = nci;
  g.ci_tiles = (Cin + nci - 1) / nci;
  g.co_tiles = (Cout + 127) / 128;
  g.cols = g.ntg * nci;
  g.nchunks = (int)((g.total + WG_KT - 1) / WG_KT);
  // position slices: about one CTA per SM over the whole grid, at least 4 chunks each
  const int other = g.NG * g.ci_tiles * g.co_tiles;
  int slices = std::max(1, std::min(g.nchunks / 4 + 1, (ctx->num_sms + other - 1) / other));
  g.chunks_per_cta = (g.nchunks + slices - 1) / slices;
  g.slices = (g.nchunks + g.chunks_per_cta - 1) / g.chunks_per_cta;
  return g;
}

// 1 when conv_wgrad_tc takes this convolution (the caller falls back to the CUDA-core kernel otherwise)
int conv_wgrad_tc_supported(int N, int Cout, int C1, int C2, int KH, int KW, int stride, int pad, int up, int Hs, int Ws) {
  const int Cin = C1 + C2;
  if (wg_mode(KH, KW, stride, pad, up, Hs, Ws) < 0 || Cin % 16 || C1 % 16 || Cout < 1) return 0;
  const long long padded = (long long)N * (Hs * up) * (Ws * up + 2);
  return padded < (1LL << 31) - 4096;
}

size_t conv_wgrad_tc_scratch_bytes(const msgm_ctx* ctx, int N, int Cout, int Cin, int KH, int KW, int stride, int pad, int up, int Hs,
                                   int Ws) {
  const WgGeom g = wg_geometry(ctx, N, Cout, Cin, KH, KW, stride, pad, up, Hs, Ws);
  return (size_t)g.slices * g.NG * g.ci_tiles * g.co_tiles * g.cols * 128 * sizeof(float);
}

int conv_wgrad_tc(msgm_ctx* ctx, const float* cot, const float* x1, const float* x2, float* gW, const unsigned int* amax_bits,
                  const unsigned int* amax_in_bits, float* scratch, int N, int Cout, int C1, int C2, int Cw, int coff, int KH, int KW, int stride, int pad, int up,
                  int Hs, int Ws, int accumulate, cudaStream_t stream) {
  WgradTcParams P{};
  P.cot = cot; P.x1 = x1; P.x2 = x2; P.gW = gW; P.amax_bits = amax_bits; P.amax_in_bits = amax_in_bits; P.scratch = scratch;
  P.C1 = C1; P.C2 = x2 ? C2 : 0; P.Cout = Cout; P.Cw = Cw; P.coff = coff; P.KH = KH; P.KW = KW;
  const int Cin = P.C1 + P.C2;
  const WgGeom g = wg_geometry(ctx, N, Cout, Cin, KH, KW, stride, pad, up, Hs, Ws);
  P.N = N; P.Hs = Hs; P.Ws = Ws; P.upsh = up == 2 ? 1 : 0; P.Hin = Hs * up; P.Win = Ws * up;
  P.Ho = g.Ho; P.Wo = g.Wo; P.Wpo = g.Wpo; P.total = g.total;
  P.pad_h = KH == 3 ? 1 : 0;
  wg_find_divisor((uint32_t)(P.Ho * P.Wpo), &P.mul_img, &P.shr_img);
  wg_find_divisor((uint32_t)P.Wpo, &P.mul_row, &P.shr_row);
  P.NG = g.NG; P.NCI = g.nci; P.nchunks = g.nchunks; P.chunks_per_cta = g.chunks_per_cta;
  int cols = 32;
  while (cols < g.cols) cols <<= 1;
  P.tmem_cols = cols;
  P.flags = next_tc_flags(ctx);
  const int sli = WG_KT + 2 * g.halo;
  const size_t smem = 128 + 128 + 2 * (size_t)(2 * 16 * WG_KT * 16) + 2 * (size_t)(g.nph * 2 * (g.nci / 8) * sli * 16);
  uint32_t sb = 0;
  int rc = dyn_smem_base(ctx, stream, &sb);
  if (rc) return rc;
  const dim3 grid((unsigned)g.slices, (unsigned)(g.NG * g.ci_tiles), (unsigned)g.co_tiles);
#define MSGM_WG_LAUNCH(MODE_)                                                                                              \
  {                                                                                                                        \
    auto kern = sb == 1024u ? conv_wgrad_tc_kernel<MODE_, true> : conv_wgrad_tc_kernel<MODE_, false>;                      \
    MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                     \
    kern<<<grid, WG_STAGERS + 32, smem, stream>>>(P);                                                                      \
  }
  if (g.mode == 0) MSGM_WG_LAUNCH(0) else if (g.mode == 1) MSGM_WG_LAUNCH(1) else if (g.mode == 2) MSGM_WG_LAUNCH(2) else MSGM_WG_LAUNCH(3)
#undef MSGM_WG_LAUNCH
  MSGM_CUDA_TRY(cudaGetLastError());
  const long long nel = (long long)grid.z * grid.y * g.cols * 128;
  wgrad_reduce_kernel<<<(unsigned)std::min<long long>((nel + 255) / 256, (long long)ctx->num_sms * 8), 256, 0, stream>>>(
      scratch, gW, (int)grid.x, (int)grid.y, (int)grid.z, g.cols, g.nci, g.NG, Cout, Cin, Cw, coff, KH, KW, accumulate);
  ctx->launches += 2;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm

using namespace msgm;

extern "C" {

int msgm_conv_wgrad_tc_ok(int32_t N, int32_t Cout, int32_t C1, int32_t C2, int32_t KH, int32_t KW, int32_t stride, int32_t pad,
                          int32_t up, int32_t Hs, int32_t Ws) {
  if (N < 1 || Cout < 1 || C1 < 1 || C2 < 0 || Hs < 1 || Ws < 1) return 0;
  return conv_wgrad_tc_supported(N, Cout, C1, C2, KH, KW, stride, pad, up, Hs, Ws);
}

uint64_t msgm_conv_wgrad_tc_scratch_bytes(const msgm_ctx* ctx, int32_t N, int32_t Cout, int32_t Cin, int32_t KH, int32_t KW,
                                          int32_t stride, int32_t pad, int32_t up, int32_t Hs, int32_t Ws) {
  if (!ctx || !msgm_conv_wgrad_tc_ok(N, Cout, Cin, 0, KH, KW, stride, pad, up, Hs, Ws)) return 0;
  return conv_wgrad_tc_scratch_bytes(ctx, N, Cout, Cin, KH, KW, stride, pad, up, Hs, Ws);
}

int msgm_conv_wgrad_tc(msgm_ctx* ctx, const float* cot, const float* in1, const float* in2, float* gW_accumulate,
                       const float* amax_or_null, const float* amax_in_or_null, void* scratch, int32_t N, int32_t Cout, int32_t C1,
                       int32_t C2, int32_t Cw, int32_t coff, int32_t KH, int32_t KW, int32_t stride, int32_t pad, int32_t up, int32_t Hs, int32_t Ws,
                       int32_t accumulate, void* stream) {
  if (!ctx || !cot || !in1 || !gW_accumulate || !scratch || N < 1 || Cout < 1 || C1 < 1 || C2 < 0 || (C2 > 0 && !in2) ||
      coff < 0 || coff + C1 + C2 > Cw || Hs < 1 || Ws < 1) {
    set_error("msgm_conv_wgrad_tc: bad argument");
    return MSGM_ERR_INVALID;
  }
  if (!msgm_conv_wgrad_tc_ok(N, Cout, C1, C2, KH, KW, stride, pad, up, Hs, Ws)) {
    set_error("msgm_conv_wgrad_tc: shape not taken by the tensor-core kernel (3x3 / 1x3 / 1x1 stride 1, 1x4 / 3x3 stride 2; channels % 16)");
    return MSGM_ERR_UNSUPPORTED;
  }
  MSGM_CUDA_TRY(cudaSetDevice(ctx->device));
  return conv_wgrad_tc(ctx, cot, in1, in2, gW_accumulate, reinterpret_cast<const unsigned int*>(amax_or_null),
                       reinterpret_cast<const unsigned int*>(amax_in_or_null), reinterpret_cast<float*>(scratch), N, Cout, C1, C2, Cw, coff, KH, KW, stride, pad, up, Hs, Ws, accumulate,
                       (cudaStream_t)stream);
}

}  // extern "C"
