// 1-D U-Net convolutions, TMA-fed (NNUnet1D.py:13-33 ConvBlock1D, :81-102 encoder / decoder convs, :165-169 ConvTranspose1d).
//
// conv2d_tc.cu evaluates a conv as a shift GEMM whose A operand is staged from an fp32 NCL tensor by 256 threads (load,
// split into fp16 hi + lo, store to shared memory): at the 1-D U-Net's sizes that staging, not the tensor pipe, sets the
// time (ncu: tensor pipe 17-31 % active, issue slots 50 %).  The 1-D net has no normalisation between its convs
// (conv -> GELU -> conv -> GELU), so nothing has to happen to an activation between the epilogue that produces it and the
// MMA that consumes it.  Here the activations therefore LIVE in the operand format ("planes"):
//
//   planes of a (B, C, L) tensor = fp16 [hi | lo][C / 8][R][8],  R = 2 GUARD + B (L + 3) rows of 16 bytes;
//   row of (b, l) = GUARD + b (L + 3) + 1 + l: one zero row left of every signal, two right of it (the padding of k3 p1 and
//   k4 s2 p1 convs alike), GUARD zero rows at both ends of a plane so that a tile may overhang.
//
// With that layout the A tile of a 16-channel chunk is four contiguous runs of SL rows (hi / lo x two 8-channel groups):
// four `cp.async.bulk` copies put it in shared memory exactly as the canonical no-swizzle K-major core-matrix layout wants
// it (a tap shift is a 16-byte multiple of the start address, as in conv2d_tc.cu), the packed weights are a fifth copy, and
// no thread touches an operand.  The kernel is persistent and warp-specialised:
//
//   warp 16     producer: per chunk, waits for a free stage and issues the five bulk copies onto the stage's mbarrier;
//   warp 17     MMA issuer: MB x taps x 3 tcgen05.mma per chunk (A_hi W_hi + A_lo W_hi + A_hi W_lo: fp32-level parity),
//               tcgen05.commit to the stage's "empty" barrier; accumulators are double-buffered in TMEM (2 x MB x NOUT
//               columns), so the next tile's products run while
//   warps 0-15  drain the previous accumulator (warp = TMEM lane quarter x column half x every second 128-row block):
//               + bias + folded embedding taps, exact GELU, split into hi + lo and store the OUTPUT planes (16-byte
//               stores, consecutive rows per lane) and / or an fp32 NCL tensor.  The epilogue is the larger half of the
//               work for the narrow layers (37 -> 24 instructions per output value with the rational erf below), hence
//               four epilogue warps per SM sub-partition.
//
// Zero padding is never written: plane buffers are zero-filled once when they are allocated and the epilogue only ever
// stores rows of real positions, so ring and guard rows stay zero for the life of the buffer.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

int conv2d_tc_nout(int Cout, int taps);  // conv2d_tc.cu: output-channel tile the packed weight image was built for

constexpr int PL_GUARD = 640;  // zero rows at both ends of a plane (>= 4 x 128 rows of tile overhang + halo)
constexpr int PL_PAD = 3;      // 1 left + 2 right zero rows per signal
constexpr int PL_HALO = 2;     // rows staged on both sides of a tile (taps reach -1 .. +2)
constexpr int TCP_EPI_WARPS = 16;
constexpr int TCP_THREADS = 32 * TCP_EPI_WARPS + 64;

static inline long long planes_rows(long long B, int L) { return 2LL * PL_GUARD + B * (long long)(L + PL_PAD); }

struct TcpParams {
  const unsigned char* x1; int C1;  // input planes (the concat [x1, x2] is read in place)
  const unsigned char* x2; int C2;
  const unsigned char* wimg;        // conv2d_tc_pack_kernel / convt1d_tc_pack_kernel image
  const float* bias;                // (CoutT) or NULL
  const float* etab;                // (B, Cout, NT) folded embedding channels per tap or NULL
  unsigned char* outp;              // output planes or NULL
  float* outf;                      // fp32 (B, CoutT, Lout) or NULL
  int gelu, convt, stride, fast;
  int B, Cout, CoutT, Lin, Lout, Wp, Wpo;  // Cout: N of the product (2 CoutT for a transposed conv); Wp = Lin + 3
  long long Rin, Rout, total;              // rows per 8-channel group of the input / output planes; B Wp
  int MB, NC, SL, S;                       // 128-row blocks per tile, 16-channel chunks, staged rows, pipeline stages
  int ntile_n;
  long long ntiles;
  uint32_t mul_row, shr_row;               // n / Wp for n < 2^31
  int tmem_cols;
  TcFlags flags;
  long long* prof;  // NULL, or 24 cycle counters of CTA 0 (MSGM_TCP_PROF=1; msgm_debug_counters): [0] epilogue warp 0 waits for
                    // an accumulator, [1] drains it; [8] MMA warp waits for a free accumulator, [9] for a stage, [10] issues;
                    // [16] producer waits for a free stage, [17] issues copies; [2] / [11] / [18] = tiles / chunks seen
};

struct TcpProf {
  long long* c;
  long long t;
  __device__ __forceinline__ void start() { if (c) t = clock64(); }
  __device__ __forceinline__ void tick(int slot) {
    if (c) { const long long n = clock64(); c[slot] += n - t; t = n; }
  }
  __device__ __forceinline__ void count(int slot) { if (c) c[slot] += 1; }
};

__device__ __forceinline__ int tcp_div(int n, uint32_t mul, uint32_t shr) { return (int)(__umulhi((uint32_t)n, mul) >> shr); }

// Exact (erf) GELU, nn.GELU() of ConvBlock1D (NNUnet1D.py:13-33), with erf as the branch-free rational minimax
// approximation x P(x^2) / Q(x^2) on [-4, 4] (degree 6 / 4 in x^2; the coefficient set XLA and Eigen use for fp32 erf):
// |erf error| <= 4.5e-7, |GELU error| <= 2.6e-7 |v| (checked against float64 on 2e6 points) -- fp32 level, at 19
// instructions where erff() costs ~35 with both of its branches taken by most warps.
__device__ __forceinline__ float gelu_rational(float v) {
  const float x = fminf(fmaxf(v * 0.70710678118654752440f, -4.0f), 4.0f), x2 = x * x;
  float p = -2.72614225801306e-10f;
  p = fmaf(p, x2, 2.77068142495902e-08f);
  p = fmaf(p, x2, -2.10102402082508e-06f);
  p = fmaf(p, x2, -5.69250639462346e-05f);
  p = fmaf(p, x2, -7.34990630326855e-04f);
  p = fmaf(p, x2, -2.95459980854025e-03f);
  p = fmaf(p, x2, -1.60960333262415e-02f);
  float q = -1.45660718464996e-05f;
  q = fmaf(q, x2, -2.13374055278905e-04f);
  q = fmaf(q, x2, -1.68282697438203e-03f);
  q = fmaf(q, x2, -7.37332916720468e-03f);
  q = fmaf(q, x2, -1.42647390514189e-02f);
  const float e = __fdividef(p * x, q), hv = 0.5f * v;
  return fmaf(hv, e, hv);
}

// The same function on two values with the packed fp32 instructions of sm_100 (FFMA2 / FMUL2: two IEEE fp32 operations per
// issue slot, bit-identical to the scalar form): the epilogue is issue-bound, and the two Horner chains are 11 of its ~28
// instructions per value.
#ifndef MSGM_TCP_GELU2
#define MSGM_TCP_GELU2 1
#endif
__device__ __forceinline__ unsigned long long f2pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long f2dup(float c) { return f2pack(c, c); }
__device__ __forceinline__ unsigned long long f2fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f2mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void gelu_rational_x2(float& v0, float& v1) {
#if MSGM_TCP_GELU2
  const float x0 = fminf(fmaxf(v0 * 0.70710678118654752440f, -4.0f), 4.0f);
  const float x1 = fminf(fmaxf(v1 * 0.70710678118654752440f, -4.0f), 4.0f);
  const unsigned long long x = f2pack(x0, x1), x2 = f2mul(x, x);
  unsigned long long p = f2dup(-2.72614225801306e-10f);
  p = f2fma(p, x2, f2dup(2.77068142495902e-08f));
  p = f2fma(p, x2, f2dup(-2.10102402082508e-06f));
  p = f2fma(p, x2, f2dup(-5.69250639462346e-05f));
  p = f2fma(p, x2, f2dup(-7.34990630326855e-04f));
  p = f2fma(p, x2, f2dup(-2.95459980854025e-03f));
  p = f2fma(p, x2, f2dup(-1.60960333262415e-02f));
  unsigned long long q = f2dup(-1.45660718464996e-05f);
  q = f2fma(q, x2, f2dup(-2.13374055278905e-04f));
  q = f2fma(q, x2, f2dup(-1.68282697438203e-03f));
  q = f2fma(q, x2, f2dup(-7.37332916720468e-03f));
  q = f2fma(q, x2, f2dup(-1.42647390514189e-02f));
  p = f2mul(p, x);
  float p0, p1, q0, q1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(p0), "=f"(p1) : "l"(p));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(q));
  const float e0 = __fdividef(p0, q0), e1 = __fdividef(p1, q1), h0 = 0.5f * v0, h1 = 0.5f * v1;
  v0 = fmaf(h0, e0, h0);
  v1 = fmaf(h1, e1, h1);
#else
  v0 = gelu_rational(v0);
  v1 = gelu_rational(v1);
#endif
}

template <int NOUT, int NT, bool CONST_BASE>
__global__ void __launch_bounds__(TCP_THREADS, 1) conv1d_tcp_kernel(const __grid_constant__ TcpParams P) {
  constexpr int WCHUNK = NT * 2 * 2 * NOUT * 16;  // [hi|lo][tap][kc][NOUT][8] fp16 per 16-channel chunk
  const int WSTAGE = P.fast ? WCHUNK / 2 : WCHUNK;
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem);  // [8] stage filled (5 bulk copies landed)
  uint64_t* bar_empty = bar_full + 8;                      // [8] stage consumed (its MMAs completed)
  uint64_t* bar_accf = bar_full + 16;                      // [2] accumulator complete
  uint64_t* bar_acce = bar_full + 18;                      // [2] accumulator drained (every epilogue warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 20);
  const int PS = P.SL * 16;                                // one plane of a stage (bytes)
  const int ASTAGE = (P.fast ? 2 : 4) * PS;                // [hi|lo][8-channel group]
  const int STAGE = ASTAGE + WSTAGE;
  unsigned char* stage0 = smem + 256;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int S = P.S, MB = P.MB;

  if (tid == 32 * TCP_EPI_WARPS) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + s, 1);
      mbar_init(bar_empty + s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_accf + a, 1);
      mbar_init(bar_acce + a, TCP_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TCP_EPI_WARPS + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(P.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == TCP_EPI_WARPS) {
    // ================================================= producer ===================================================
    // Lane 0 waits for the stage and posts the byte count; lanes 0..4 then issue one bulk copy each (a single thread needs
    // ~700 clk to issue the five copies of a chunk, measured with MSGM_TCP_PROF: more than the chunk's MMAs take in the
    // single-product mode).  MSGM_TCP_PROD5=0 keeps everything on lane 0.
#ifndef MSGM_TCP_PROD5
#define MSGM_TCP_PROD5 1
#endif
    {
      int s = 0;
      uint32_t ph = 1;  // a fresh barrier passes a wait on parity 1: the first S stages are free
      bool ok = true;
      const size_t G1 = (size_t)(P.C1 >> 3) * (size_t)P.Rin * 16, G2 = (size_t)(P.C2 >> 3) * (size_t)P.Rin * 16;  // hi -> lo plane
      TcpProf pf{(P.prof && blockIdx.x == 0 && lane == 0) ? P.prof + 16 : nullptr, 0};
      pf.start();
      for (long long tile = blockIdx.x; tile < P.ntiles && ok; tile += gridDim.x) {
        const long long m = tile / P.ntile_n;
        const int n = (int)(tile - m * P.ntile_n);
        const long long row0 = PL_GUARD + m * (128LL * MB) - PL_HALO;
        for (int k = 0; k < P.NC && ok; ++k) {
          if (lane == 0) {
            ok = mbar_wait(bar_empty + s, ph, P.flags);
            if (ok) mbar_expect_tx(bar_full + s, (uint32_t)(ASTAGE + WSTAGE));
          }
          ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;
          if (!ok) break;
          pf.tick(0);
          unsigned char* dst = stage0 + (size_t)s * STAGE;
          const int ch0 = k * 16;
          const bool from1 = ch0 < P.C1;
          const unsigned char* xb = from1 ? P.x1 : P.x2;
          const int kc0 = (from1 ? ch0 : ch0 - P.C1) >> 3;
          const size_t glo = from1 ? G1 : G2;
          const unsigned char* src = xb + ((size_t)kc0 * (size_t)P.Rin + (size_t)row0) * 16;
          // copy i of the chunk: 0, 1 = hi planes of the two 8-channel groups, 2, 3 = their lo planes, 4 = weights
          const int nA = P.fast ? 2 : 4;
#if MSGM_TCP_PROD5
          if (lane < nA) {
            tma_bulk_g2s(dst + lane * PS, src + (size_t)(lane >> 1) * glo + (size_t)(lane & 1) * (size_t)P.Rin * 16, (uint32_t)PS,
                         bar_full + s);
          } else if (lane == 4) {
            tma_bulk_g2s(dst + ASTAGE, P.wimg + ((size_t)n * P.NC + k) * WCHUNK, (uint32_t)WSTAGE, bar_full + s);
          }
#else
          if (lane == 0) {
            for (int i = 0; i < nA; ++i)
              tma_bulk_g2s(dst + i * PS, src + (size_t)(i >> 1) * glo + (size_t)(i & 1) * (size_t)P.Rin * 16, (uint32_t)PS,
                           bar_full + s);
            tma_bulk_g2s(dst + ASTAGE, P.wimg + ((size_t)n * P.NC + k) * WCHUNK, (uint32_t)WSTAGE, bar_full + s);
          }
#endif
          pf.tick(1);
          pf.count(2);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == TCP_EPI_WARPS + 1) {
    // ================================================ MMA issuer ==================================================
    const uint32_t idesc = umma_idesc_f16(128, NOUT);
    const uint32_t sbase = CONST_BASE ? 1024u : smem_u32(smem);
    if (sbase != smem_u32(smem)) {
      if (lane == 0) tc_raise(P.flags, 2);
    } else {
      int s = 0;
      uint32_t ph = 0;
      bool ok = true;
      int tl = 0;
      TcpProf pf{(P.prof && blockIdx.x == 0 && lane == 0) ? P.prof + 8 : nullptr, 0};
      pf.start();
      for (long long tile = blockIdx.x; tile < P.ntiles && ok; tile += gridDim.x, ++tl) {
        const int ab = tl & 1;
        ok = __all_sync(0xffffffffu, mbar_wait(bar_acce + ab, (uint32_t)(((tl >> 1) & 1) ^ 1), P.flags));
        if (!ok) break;
        tc_fence_after();
        pf.tick(0);
        const uint32_t dbase = tbase + (uint32_t)(ab * MB * NOUT);
        for (int k = 0; k < P.NC; ++k) {
          if (!__all_sync(0xffffffffu, mbar_wait(bar_full + s, ph, P.flags))) { ok = false; break; }
          tc_fence_after();
          pf.tick(1);
          // descriptors: only the 14-bit start-address field (16-byte units) moves between the MMAs of a chunk
          const uint32_t a_base = sbase + 256u + (uint32_t)(s * STAGE);
          const uint64_t dA0 = umma_desc(a_base + (uint32_t)((PL_HALO - 1) * 16), PS, 128);  // tap 0 of block 0, hi planes
          const uint64_t dW0 = umma_desc(a_base + (uint32_t)ASTAGE, NOUT * 16, 128);         // tap 0, hi image
          const uint64_t lo_a = (uint64_t)((2 * PS) >> 4), lo_w = (uint64_t)((NT * 2 * NOUT * 16) >> 4);
#pragma unroll 1
          for (int mb = 0; mb < MB; ++mb) {
            const uint32_t dcol = dbase + (uint32_t)(mb * NOUT);
            const uint64_t dAm = dA0 + (uint64_t)(mb * 128);
#pragma unroll
            for (int t = 0; t < NT; ++t) {
              const uint64_t dAh = dAm + (uint64_t)t, dWh = dW0 + (uint64_t)(t * ((2 * NOUT * 16) >> 4));
              umma_ss(dcol, dAh, dWh, idesc, (k > 0 || t > 0) ? 1u : 0u, 0);
              if (!P.fast) {
                umma_ss(dcol, dAh + lo_a, dWh, idesc, 1u, 0);
                umma_ss(dcol, dAh, dWh + lo_w, idesc, 1u, 0);
              }
            }
          }
          umma_commit(bar_empty + s, 0);
          pf.tick(2);
          pf.count(3);
          if (++s == S) { s = 0; ph ^= 1u; }
        }
        if (ok) umma_commit(bar_accf + ab, 0);
      }
    }
    __syncwarp();
  } else {
    // ================================================== epilogue ==================================================
    const int q4 = warp & 3, half = (warp >> 2) & 1, mb0 = warp >> 3;  // lane quarter, column half, blocks mb0, mb0 + 2, ..
    constexpr int NH = NOUT / 2;
    bool ok = true;
    int tl = 0;
    const size_t olo = (size_t)(P.CoutT >> 3) * (size_t)P.Rout * 16;  // hi -> lo plane of the output
    TcpProf pf{(P.prof && blockIdx.x == 0 && tid == 0) ? P.prof : nullptr, 0};
    pf.start();
    for (long long tile = blockIdx.x; tile < P.ntiles && ok; tile += gridDim.x, ++tl) {
      const int ab = tl & 1;
      const long long m = tile / P.ntile_n;
      const int co0 = (int)(tile - m * P.ntile_n) * NOUT;
      ok = mbar_wait(bar_accf + ab, (uint32_t)((tl >> 1) & 1), P.flags);
      tc_fence_after();
      pf.tick(0);
      for (int mb = mb0; mb < MB; mb += 2) {
        const long long p = m * (128LL * MB) + mb * 128 + q4 * 32 + lane;
        bool valid = ok && p < P.total;
        int b = 0, c = 0, ocol = 0;
        if (valid) {
          b = tcp_div((int)p, P.mul_row, P.shr_row);
          c = (int)p - b * P.Wp - 1;
          valid = c >= 0 && c < P.Lin;
          ocol = c;
          if (P.stride == 2) {
            valid = valid && !(c & 1);
            ocol = c >> 1;
            valid = valid && ocol < P.Lout;
          }
        }
        const uint32_t taddr = tbase + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(ab * MB * NOUT + mb * NOUT + half * NH);
#pragma unroll 1
        for (int cc = 0; cc < NH; cc += 16) {
          uint32_t rr[16];
          TMEM_LD16(taddr + cc, rr);
          tc_wait_ld();
          if (!valid) continue;
          const int cbase = co0 + half * NH + cc;  // first of 16 consecutive columns of the product
          int cb = cbase, oc = ocol;
          if (P.convt) {  // columns [0, CoutT) are the even outputs, [CoutT, 2 CoutT) the odd ones
            const int par = cbase >= P.CoutT ? 1 : 0;
            cb = cbase - par * P.CoutT;
            oc = 2 * c + par;
            if (oc >= P.Lout) continue;
          }
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
          if (P.bias) {
            const float* bp = P.bias + cb;
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(bp + j);
          }
          if (P.etab) {  // embedding channels are constant along the signal: a tap contributes where it reads inside it
            const float* et = P.etab + ((size_t)b * P.Cout + cbase) * NT;
            float mk[NT];
#pragma unroll
            for (int t = 0; t < NT; ++t) mk[t] = (c + t - 1 >= 0 && c + t - 1 < P.Lin) ? 1.0f : 0.0f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
#pragma unroll
              for (int t = 0; t < NT; ++t) v[j] = fmaf(mk[t], __ldg(et + j * NT + t), v[j]);
            }
          }
          if (P.gelu) {
#pragma unroll
            for (int j = 0; j < 16; j += 2) gelu_rational_x2(v[j], v[j + 1]);
          }
          if (P.outf) {
            float* op = P.outf + ((size_t)b * P.CoutT + cb) * P.Lout + oc;
#pragma unroll
            for (int j = 0; j < 16; ++j) op[(size_t)j * P.Lout] = v[j];
          }
          if (P.outp) {
            const size_t orow = (size_t)PL_GUARD + (size_t)b * P.Wpo + 1 + oc;
            unsigned char* ob = P.outp + ((size_t)(cb >> 3) * (size_t)P.Rout + orow) * 16;
            uint4 h0, l0, h1, l1;
            split2_f16(v[0], v[1], h0.x, l0.x); split2_f16(v[2], v[3], h0.y, l0.y);
            split2_f16(v[4], v[5], h0.z, l0.z); split2_f16(v[6], v[7], h0.w, l0.w);
            split2_f16(v[8], v[9], h1.x, l1.x); split2_f16(v[10], v[11], h1.y, l1.y);
            split2_f16(v[12], v[13], h1.z, l1.z); split2_f16(v[14], v[15], h1.w, l1.w);
            *reinterpret_cast<uint4*>(ob) = h0;
            *reinterpret_cast<uint4*>(ob + (size_t)P.Rout * 16) = h1;
            *reinterpret_cast<uint4*>(ob + olo) = l0;
            *reinterpret_cast<uint4*>(ob + olo + (size_t)P.Rout * 16) = l1;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acce + ab);
      pf.tick(1);
      pf.count(2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TCP_EPI_WARPS + 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(P.tmem_cols));
}

// ---- plane <-> fp32 NCL conversion (first / last layers, tests, layers the tensor-core path does not cover) ----------------
// thread = one row of one 8-channel group
__global__ void __launch_bounds__(256) planes_pack_kernel(const float* __restrict__ x, unsigned char* __restrict__ pl, int B, int C,
                                                          int L, long long R, long long nitem) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= nitem) return;
  const int l = (int)(e % L);
  const long long r = e / L;
  const int b = (int)(r % B), kc = (int)(r / B);
  const float* src = x + ((size_t)b * C + kc * 8) * L + l;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __ldg(src + (size_t)j * L);
  uint4 hi, lo;
  split2_f16(v[0], v[1], hi.x, lo.x); split2_f16(v[2], v[3], hi.y, lo.y);
  split2_f16(v[4], v[5], hi.z, lo.z); split2_f16(v[6], v[7], hi.w, lo.w);
  unsigned char* dst = pl + ((size_t)kc * (size_t)R + (size_t)PL_GUARD + (size_t)b * (L + PL_PAD) + 1 + l) * 16;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + (size_t)(C >> 3) * (size_t)R * 16) = lo;
}

__global__ void __launch_bounds__(256) planes_unpack_kernel(const unsigned char* __restrict__ pl, float* __restrict__ x, int B, int C,
                                                            int L, long long R, long long nitem) {
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  if (e >= nitem) return;
  const int l = (int)(e % L);
  const long long r = e / L;
  const int b = (int)(r % B), kc = (int)(r / B);
  const unsigned char* src = pl + ((size_t)kc * (size_t)R + (size_t)PL_GUARD + (size_t)b * (L + PL_PAD) + 1 + l) * 16;
  const uint4 hi = *reinterpret_cast<const uint4*>(src);
  const uint4 lo = *reinterpret_cast<const uint4*>(src + (size_t)(C >> 3) * (size_t)R * 16);
  const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
  float* dst = x + ((size_t)b * C + kc * 8) * L + l;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hw[j]));
    const float2 c2 = __half22float2(*reinterpret_cast<const __half2*>(&lw[j]));
    dst[(size_t)(2 * j) * L] = a.x + c2.x;
    dst[(size_t)(2 * j + 1) * L] = a.y + c2.y;
  }
}

// First conv of the net (NNUnet1D.py:81-84): ONE real input channel + the folded embedding table, k3 p1, exact GELU,
// written straight into planes.  HBM-bound on its output; thread = one position, Cout % 8 == 0, Cout <= 128.
__global__ void __launch_bounds__(256) conv1d_first_planes_kernel(const float* __restrict__ x, const float* __restrict__ W, int Cw,
                                                                  const float* __restrict__ bias, const float* __restrict__ E,
                                                                  unsigned char* __restrict__ pl, int B, int Cout, int L, long long R,
                                                                  int gelu) {
  extern __shared__ float swb[];  // [Cout][3] weights of the real channel + [Cout] bias
  for (int e = threadIdx.x; e < Cout * 3; e += 256) swb[e] = W[(size_t)(e / 3) * Cw * 3 + e % 3];
  for (int e = threadIdx.x; e < Cout; e += 256) swb[Cout * 3 + e] = bias ? bias[e] : 0.0f;
  __syncthreads();
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= (long long)B * L) return;
  const int b = (int)(p / L), l = (int)(p % L);
  const float* xr = x + (size_t)b * L;
  const bool hasl = l > 0, hasr = l + 1 < L;
  const float xm = hasl ? __ldg(xr + l - 1) : 0.0f, x0 = __ldg(xr + l), xp = hasr ? __ldg(xr + l + 1) : 0.0f;
  unsigned char* dst = pl + ((size_t)PL_GUARD + (size_t)b * (L + PL_PAD) + 1 + l) * 16;
  const size_t glo = (size_t)(Cout >> 3) * (size_t)R * 16;
  for (int c8 = 0; c8 < Cout; c8 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = c8 + j;
      float a = swb[Cout * 3 + co];
      a = fmaf(swb[co * 3], xm, a);
      a = fmaf(swb[co * 3 + 1], x0, a);
      a = fmaf(swb[co * 3 + 2], xp, a);
      if (E) {
        const float* e = E + ((size_t)b * Cout + co) * 3;
        a += __ldg(e + 1);
        if (hasl) a += __ldg(e);
        if (hasr) a += __ldg(e + 2);
      }
      v[j] = a;
    }
    if (gelu) {
#pragma unroll
      for (int j = 0; j < 8; j += 2) gelu_rational_x2(v[j], v[j + 1]);
    }
    uint4 hi, lo;
    split2_f16(v[0], v[1], hi.x, lo.x); split2_f16(v[2], v[3], hi.y, lo.y);
    split2_f16(v[4], v[5], hi.z, lo.z); split2_f16(v[6], v[7], hi.w, lo.w);
    unsigned char* d8 = dst + (size_t)(c8 >> 3) * (size_t)R * 16;
    *reinterpret_cast<uint4*>(d8) = hi;
    *reinterpret_cast<uint4*>(d8 + glo) = lo;
  }
}

// ---- host dispatch ---------------------------------------------------------------------------------------------------
static void tcp_find_divisor(uint32_t d, uint32_t* mul, uint32_t* shr) {  // n / d == umulhi(n, mul) >> shr, 0 <= n < 2^31, d >= 2
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const uint32_t p = 31 + l;
  *mul = (uint32_t)(((1ull << p) + d - 1) / d);
  *shr = p - 32;
}

int64_t planes_bytes(int64_t B, int C, int L) { return 2 * (int64_t)(C / 8) * planes_rows(B, L) * 16; }

template <int NOUT, int NT>
static int launch_tcp(msgm_ctx* ctx, TcpParams& P, cudaStream_t stream) {
  const int WSTAGE = NT * (P.fast ? 1 : 2) * 2 * NOUT * 16;
  const int nplane = P.fast ? 2 : 4;
  const long long nblk = (P.total + 127) / 128;
  // 128-row blocks per tile: 2 x MB x NOUT accumulator columns must fit TMEM; among those, the MB with the fewest
  // "waves x MB" (whole waves of the persistent grid), larger MB on ties (weights are fetched once per tile)
  int MB = 0;
  long long best = 0;
  for (int mb = 1; mb <= 4; ++mb) {
    if (2 * mb * NOUT > 512) break;
    const long long tiles = ((nblk + mb - 1) / mb) * P.ntile_n;
    const long long cost = ((tiles + ctx->num_sms - 1) / ctx->num_sms) * mb;
    if (MB == 0 || cost <= best) { best = cost; MB = mb; }
  }
  P.MB = MB;
  P.SL = 128 * MB + 2 * PL_HALO;
  const size_t stage = (size_t)nplane * P.SL * 16 + (size_t)WSTAGE;
  P.S = (int)std::min<size_t>(8, ((size_t)227 * 1024 - 256 - 128) / stage);
  if (P.S < 2) {
    set_error("msgm_conv1d_tcp: two pipeline stages do not fit shared memory");
    return MSGM_ERR_UNSUPPORTED;
  }
  P.ntiles = ((nblk + MB - 1) / MB) * P.ntile_n;
  int cols = 32;
  while (cols < 2 * MB * NOUT) cols <<= 1;
  P.tmem_cols = cols;
  const size_t smem = 256 + 128 + (size_t)P.S * stage;
  uint32_t sb = 0;
  int rc = dyn_smem_base(ctx, stream, &sb);
  if (rc) return rc;
  auto kern = sb == 1024u ? conv1d_tcp_kernel<NOUT, NT, true> : conv1d_tcp_kernel<NOUT, NT, false>;
  MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<long long>(P.ntiles, (long long)ctx->num_sms);
  kern<<<grid, TCP_THREADS, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int NT>
static int launch_tcp_n(msgm_ctx* ctx, TcpParams& P, cudaStream_t stream) {
  const int nout = conv2d_tc_nout(P.Cout, NT);
  P.ntile_n = P.Cout / nout;
  if (nout == 192) return launch_tcp<192, NT>(ctx, P, stream);
  if (nout == 128) return launch_tcp<128, NT>(ctx, P, stream);
  if (nout == 64) return launch_tcp<64, NT>(ctx, P, stream);
  return launch_tcp<32, NT>(ctx, P, stream);
}

int conv1d_tcp(msgm_ctx* ctx, const msgm_conv1d_tcp_desc* D, cudaStream_t stream) {
  if (D->B == 0) return MSGM_OK;
  TcpParams P{};
  P.x1 = reinterpret_cast<const unsigned char*>(D->x1); P.C1 = D->C1;
  P.x2 = reinterpret_cast<const unsigned char*>(D->x2); P.C2 = D->x2 ? D->C2 : 0;
  P.wimg = reinterpret_cast<const unsigned char*>(D->wimg);
  P.bias = D->bias; P.etab = D->E;
  P.outp = reinterpret_cast<unsigned char*>(D->out_planes); P.outf = D->out_f32;
  P.gelu = D->gelu; P.fast = D->fast ? 1 : 0;
  P.B = D->B; P.Lin = D->Lin;
  P.Wp = D->Lin + PL_PAD;
  P.total = (long long)D->B * P.Wp;
  if (P.total >= (1LL << 31) - 4096) {
    set_error("msgm_conv1d_tcp: more than 2^31 padded positions in one call; split the batch");
    return MSGM_ERR_UNSUPPORTED;
  }
  tcp_find_divisor((uint32_t)P.Wp, &P.mul_row, &P.shr_row);
  P.NC = (P.C1 + P.C2) / 16;
  P.Rin = planes_rows(D->B, D->Lin);
  if (D->transposed) {  // ConvTranspose1d k4 s2 p1 as a 3-tap conv with 2 Cout columns (convt1d_tc_pack_kernel image)
    P.convt = 1; P.stride = 1; P.CoutT = D->Cout; P.Cout = 2 * D->Cout; P.Lout = D->Lout;
  } else {
    P.convt = 0; P.stride = D->K == 4 ? 2 : 1; P.CoutT = D->Cout; P.Cout = D->Cout;
    P.Lout = D->K == 4 ? (D->Lin + 2 - 4) / 2 + 1 : D->Lin;
  }
  P.Wpo = P.Lout + PL_PAD;
  P.Rout = planes_rows(D->B, P.Lout);
  P.flags = next_tc_flags(ctx);
  P.prof = std::getenv("MSGM_TCP_PROF") ? reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(ctx->ws) + 64) : nullptr;
  if (P.prof) MSGM_CUDA_TRY(cudaMemsetAsync(P.prof, 0, 192, stream));
  return (D->K == 4 && !D->transposed) ? launch_tcp_n<4>(ctx, P, stream) : launch_tcp_n<3>(ctx, P, stream);
}

int planes_pack(msgm_ctx* ctx, const float* x, void* planes, int B, int C, int L, cudaStream_t stream) {
  const long long nitem = (long long)(C / 8) * B * L;
  if (nitem == 0) return MSGM_OK;
  planes_pack_kernel<<<(unsigned)((nitem + 255) / 256), 256, 0, stream>>>(x, reinterpret_cast<unsigned char*>(planes), B, C, L,
                                                                        planes_rows(B, L), nitem);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int planes_unpack(msgm_ctx* ctx, const void* planes, float* x, int B, int C, int L, cudaStream_t stream) {
  const long long nitem = (long long)(C / 8) * B * L;
  if (nitem == 0) return MSGM_OK;
  planes_unpack_kernel<<<(unsigned)((nitem + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const unsigned char*>(planes), x, B, C,
                                                                          L, planes_rows(B, L), nitem);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int conv1d_first_planes(msgm_ctx* ctx, const float* x, const float* W, int Cw, const float* bias, const float* E, void* planes,
                        int B, int Cout, int L, int gelu, cudaStream_t stream) {
  const long long npos = (long long)B * L;
  if (npos == 0) return MSGM_OK;
  conv1d_first_planes_kernel<<<(unsigned)((npos + 255) / 256), 256, sizeof(float) * 4 * Cout, stream>>>(
      x, W, Cw, bias, E, reinterpret_cast<unsigned char*>(planes), B, Cout, L, planes_rows(B, L), gelu);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
