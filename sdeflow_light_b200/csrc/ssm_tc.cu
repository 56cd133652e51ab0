// Sliced-score-matching training step of the MLP score net on tcgen05 tensor cores: loss, activation backward and every
// weight gradient of one batch in ONE launch (msgm_ssm_mlp_fwd_bwd_tc), fp16 operands / fp32 accumulation.
//
// Same mathematics as ssm_fp32.cu (the fp32 parity mode; PluginReverseSDE.ssm_loss, SDEs.py:616-646, evaluated in
// forward mode and differentiated by hand): per sample the primal row u and the tangent row udot run through the MLP
// together, loss = q . adot + |a|^2 / 2 (+ beta |v|^2 / 2 for the additive SDE), and the backward pass carries the
// cotangent pair (zbar, zdotbar) with phi', phi'' of Swish.
//
// Tile = 64 samples = 128 rows (row 2i = primal of sample i, row 2i+1 = its tangent: the two rows that need each other's
// values sit in adjacent lanes of one warp and talk through __shfl_xor(.,1)).  Row = TMEM lane.  Every tile of fp16
// activations is stored ONCE in shared memory as [column/8][row/8][row%8][column%8] and used three ways:
//   * K-major A operand of the next layer's forward product            Z_{l+1} = H_l W_{l+1}^T,
//   * K-major A operand of the data gradient                           Hbar_{l-1} = Cot_l W_l     (W_l: MN-major B),
//   * MN-major A / B operand of the weight gradient (K = the 128 rows) gW_l = Cot_l^T H_{l-1};
// the MN-major reading of that layout (SBO = 2048 B between 8-column groups, LBO = 128 B between 8-row groups, 256 B per
// K = 16 slice) was validated on hardware by tools/mn_probe.cu.  Stacking primal and tangent rows in one tile makes the
// weight gradient zbar^T h + zdotbar^T hdot a single product.  The weight images serve both the forward product (K-major)
// and the data gradient (MN-major), so no transposed copies exist.
//
// TMEM (512 columns): the fp32 pre-activations (z; zdot) of the three hidden layers stay resident in columns 0..383 from
// the forward pass until the backward pass has used them for phi' and phi'' (no activation ever goes to HBM); columns
// 384..511 are the working accumulator (output layer, data gradients), and a layer's own 128 columns receive its weight
// gradient once its pre-activations are dead.  Shared memory: weights 72-84 KB, four activation tiles 100-108 KB (each
// cotangent tile overwrites the activation tile of its own layer after the products that read it have completed).
//
// The phases of a tile are strictly sequential (one tile per SM at a time, 8 warps: warp w owns rows 32 (w & 3) .. and
// columns 64 (w >> 2) ..): at the reference's batch sizes the step is latency-bound either way, and at 16 k samples the
// tensor-pipe work is ~10 us.  Weight-gradient tiles are added into a private slice of a partial buffer per CTA (plain
// read-modify-write, no atomics, deterministic) and summed by ssm_tc_reduce_kernel.
//
// Stated tolerance of this mode (fp16 operands, tanh.approx sigmoid): loss 2e-3 relative to max|loss|, gradients 3e-3
// relative to max|g| per tensor (tests/test_ssm_gpu.py); the fp32 kernels remain the parity mode (1e-7 / 1e-6).
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "msgm_common.cuh"
#include "tc_ptx.cuh"

namespace msgm {

constexpr int STC_THREADS = 256;
constexpr int STC_TS = 64;  // samples per tile

template <int K1, int DP>
struct StcLayout {
  static constexpr int TILE = 128 * 128 * 2;
  static constexpr int oW1 = 0;                        // fp16 [K1/8][128 n][8]   B of layer 1 (bias in column K1ref)
  static constexpr int oW2 = oW1 + K1 * 256;
  static constexpr int oW3 = oW2 + TILE;
  static constexpr int oW4 = oW3 + TILE;               // fp16 [16][16 n][8]      N = 16 output rows
  static constexpr int oIn0 = oW4 + 4096;              // fp16 [K1/8][128 rows][8]
  static constexpr int oIn1 = oIn0 + K1 * 256;
  static constexpr int oIn2 = oIn1 + TILE;
  static constexpr int oIn3 = oIn2 + TILE;
  static constexpr int oC4 = oIn3 + TILE;              // fp16 [2][128 rows][8]   output cotangents (abar; adotbar)
  static constexpr int oBias = oC4 + 4096;             // fp32 b2[128] b3[128] b4[DP]
  static constexpr int oW4f = oBias + 4 * (256 + DP);  // fp32 [DP][128]          W4 for the CUDA-core data gradient
  static constexpr int oX = oW4f + 4 * DP * 128;       // fp32 [128 rows][DP]     output cotangents in fp32
  static constexpr int oDb = oX + 4 * 128 * DP;        // fp32 db2[128] db3[128] db4[DP]  (accumulated over the CTA's tiles)
  static constexpr int oQ = oDb + 4 * (256 + DP);      // fp32 [128 rows][DP]     second half of the q partial sums
  static constexpr int oG = oQ + 4 * 128 * DP;         // fp32 dense G [d][d][d] (d <= 16)
  static constexpr int oBar = oG + 4 * 16 * 16 * 16;
  static constexpr int SMEM = oBar + 64;
};

struct StcParams {
  int d, pre, kind;
  float bmin, bdel;
  const float* G;
  const float* W[4];
  const float* b[4];
  const float* y;     // (B,d)
  const float* v;     // (B,d)
  const float* t;     // (B,)
  const float* gout;  // (B,) upstream gradient of the per-sample loss
  float* loss;        // (B,)
  float* part;        // [gridDim.x][pstride] per-CTA partial gradients (torch parameter order)
  int nparam, pstride;  // pstride = nparam rounded up to 4 floats (16-byte read-modify-writes)
  float cot_scale;      // power of two applied to gout inside the kernel and removed by the reduction: the cotangents are
                        // fp16 tensor-core operands, and gout = 1/B of a batch mean would put them in the subnormal range
  long long B;
  TcFlags flags;
  long long* prof;  // NULL, or 24 cycle counters of CTA 0 / thread 0 (MSGM_TC_PROF; msgm_debug_counters)
};

constexpr uint32_t IDESC_A_MN = 1u << 15, IDESC_B_MN = 1u << 16;  // "transposed" (MN-major) operand bits

// sum over the 32 lanes of a warp of 32 per-lane values: lane L returns sum_lanes v[L]   (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// sum over the 16 lanes of equal parity of 16 per-lane values: returns sum_lanes v[pair_col(lane)]   (15 shuffles)
__device__ __forceinline__ float warp_pair_sum16(float (&v)[16], int lane) {
#pragma unroll
  for (int off = 16; off >= 2; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off / 2; ++i) {
      const float send = up ? v[i] : v[i + off / 2];
      const float keep = up ? v[i + off / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// The two rows of a sample (primal row in the even lane, tangent row in the odd lane) need each other's values in every
// epilogue.  Instead of both lanes evaluating sigmoid / phi' / phi'' of all 32 columns of a chunk, each lane takes 16 columns
// for BOTH rows: the even lane columns 0..15, the odd lane columns 16..31 (16 shuffles instead of 32, half the MUFU work).
// p[j] / t[j] = primal-row / tangent-row value of this lane's column j.
__device__ __forceinline__ void pair_split(const uint32_t (&mine)[32], bool primal, float (&p)[16], float (&t)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float lo = __uint_as_float(mine[j]), hi = __uint_as_float(mine[16 + j]);
    const float recv = __shfl_xor_sync(0xffffffffu, primal ? hi : lo, 1);
    p[j] = primal ? lo : recv;
    t[j] = primal ? recv : hi;
  }
}

// 8 fp32 -> one 16-byte store of 8 fp16 at (row, 8-column chunk) of a [chunk][128 rows][8] tile
__device__ __forceinline__ void store_chunk(unsigned char* tile, int row, int chunk, const float* x) {
  *reinterpret_cast<uint4*>(tile + chunk * 2048 + row * 16) =
      make_uint4(pack_f16x2(x[0], x[1]), pack_f16x2(x[2], x[3]), pack_f16x2(x[4], x[5]), pack_f16x2(x[6], x[7]));
}

template <int K1, int DP, int KIND>
__global__ void __launch_bounds__(STC_THREADS, 1) ssm_tc_kernel(const __grid_constant__ StcParams P) {
  using L = StcLayout<K1, DP>;
  extern __shared__ __align__(128) unsigned char smem[];
  float* sB2 = reinterpret_cast<float*>(smem + L::oBias);  // b2 | b3 | b4
  float* sW4f = reinterpret_cast<float*>(smem + L::oW4f);
  float* sX = reinterpret_cast<float*>(smem + L::oX);
  float* sDb = reinterpret_cast<float*>(smem + L::oDb);
  float* sG = reinterpret_cast<float*>(smem + L::oG);
  float* sQ = reinterpret_cast<float*>(smem + L::oQ);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L::oBar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = 32 * (warp & 3) + lane;  // tile row == TMEM lane
  const int half = warp >> 2;              // column half this thread works on in the epilogues
  const bool primal = (row & 1) == 0;
  const int d = P.d, K1ref = d + 1 + P.pre;
  float* part = P.part + (size_t)blockIdx.x * P.pstride;
  // flat gradient offsets, torch parameter order (main.0.weight, main.0.bias, main.2.*, main.4.*, main.6.*)
  const int gW1 = 0, gb1 = 128 * K1ref, gW2 = gb1 + 128, gb2 = gW2 + 128 * 128, gW3 = gb2 + 128, gb3 = gW3 + 128 * 128,
            gW4 = gb3 + 128, gb4 = gW4 + 128 * d;

  // ---- setup ----------------------------------------------------------------------------------------------------------
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  for (int i = tid; i < P.nparam; i += STC_THREADS) part[i] = 0.0f;
  // fp16 weight images [k/8][n][8] straight from the fp32 parameters (torch layout [n][k]): 32 contiguous bytes in, one
  // 16-byte store out per (n, 8-column chunk)
  for (int l = 1; l <= 2; ++l)
    for (int q = tid; q < 128 * 16; q += STC_THREADS) {
      const int n = q >> 4, kc = q & 15;
      const float4 a = __ldg(reinterpret_cast<const float4*>(P.W[l] + n * 128 + 8 * kc));
      const float4 b = __ldg(reinterpret_cast<const float4*>(P.W[l] + n * 128 + 8 * kc + 4));
      const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      store_chunk(smem + (l == 1 ? L::oW2 : L::oW3), n, kc, x);
    }
  for (int q = tid; q < 128 * (K1 / 8); q += STC_THREADS) {
    const int n = q / (K1 / 8), kc = q % (K1 / 8);
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 8 * kc + j;
      x[j] = k < K1ref ? __ldg(P.W[0] + n * K1ref + k) : (k == K1ref ? __ldg(P.b[0] + n) : 0.0f);
    }
    store_chunk(smem + L::oW1, n, kc, x);
  }
  for (int q = tid; q < 16 * 16; q += STC_THREADS) {  // W4 image: [kc][16 n][8], chunk stride 256 B
    const int n = q & 15, kc = q >> 4;
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = n < d ? __ldg(P.W[3] + n * 128 + 8 * kc + j) : 0.0f;
    *reinterpret_cast<uint4*>(smem + L::oW4 + kc * 256 + n * 16) =
        make_uint4(pack_f16x2(x[0], x[1]), pack_f16x2(x[2], x[3]), pack_f16x2(x[4], x[5]), pack_f16x2(x[6], x[7]));
  }
  for (int e = tid; e < DP * 128; e += STC_THREADS) sW4f[e] = (e >> 7) < d ? __ldg(P.W[3] + e) : 0.0f;
  for (int e = tid; e < 256; e += STC_THREADS) sB2[e] = __ldg(P.b[1 + (e >> 7)] + (e & 127));
  for (int e = tid; e < DP; e += STC_THREADS) sB2[256 + e] = e < d ? __ldg(P.b[3] + e) : 0.0f;
  for (int e = tid; e < 256 + DP; e += STC_THREADS) sDb[e] = 0.0f;
  if (KIND == MSGM_SDE_MSGM_DENSE)
    for (int e = tid; e < d * d * d; e += STC_THREADS) sG[e] = __ldg(P.G + e);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t tlane = tbase + ((uint32_t)(32 * (warp & 3)) << 16);  // this warp's lane quarter
  const uint32_t sbase = smem_u32(smem);
  uint32_t par = 0;
  bool ok = true;
  long long* prof = (P.prof && blockIdx.x == 0 && tid == 0) ? P.prof : nullptr;
  long long tprev = prof ? clock64() : 0;
  auto tick = [&](int slot) {
    if (prof) {
      const long long now = clock64();
      prof[slot] += now - tprev;
      tprev = now;
    }
  };
  tick(17);  // setup

  // K-major [chunk][rows][8] tile with `rows` rows: the two 8-column chunks of a K=16 slice are rows*16 bytes apart
  auto kdesc = [&](int off, int slice, int rows) { return umma_desc(sbase + off + slice * rows * 32, rows * 16, 128); };
  // the same tile read MN-major (K = rows): 8-column groups rows*16 bytes apart, 8-row groups 128 bytes apart
  auto mdesc = [&](int off, int slice, int rows) { return umma_desc(sbase + off + slice * 256, 128, rows * 16); };
  auto publish = [&]() {  // generic-proxy writes of all threads -> visible to the MMAs issued after this point
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  };
  auto wait_mma = [&]() {
    ok = mbar_wait(bar, par, P.flags) && ok;
    par ^= 1u;
    tc_fence_after();
  };
  constexpr uint32_t ID_H = umma_idesc_f16(128, 128), ID_O = umma_idesc_f16(128, 16), ID_1 = umma_idesc_f16(128, K1);
  constexpr uint32_t D1 = 0, D2 = 128, D3 = 256, WK = 384;  // TMEM column bases

  const int rowP = row & ~1, rowT = row | 1;  // the sample's primal / tangent tile rows
  // Forward epilogue of hidden layer l (1..3): (z; zdot) in TMEM -> (h; hdot) = (phi(z); phi'(z) zdot) as fp16 rows of `dst`.
  auto fwd_epilogue = [&](uint32_t dcol, const float* bias, int dst) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = 64 * half + 32 * c, cb = col0 + (primal ? 0 : 16);
      uint32_t r[32];
      TMEM_LD32(tlane + dcol + col0, r);
      tc_wait_ld();
      float z[16], zd[16];
      pair_split(r, primal, z, zd);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float zz = z[j] + (bias != nullptr ? bias[cb + j] : 0.0f);
        const float sg = fmaf(0.5f, tanh_fast(0.5f * zz), 0.5f);  // sigmoid(z), one MUFU op
        z[j] = zz * sg;
        zd[j] = sg * fmaf(zz, 1.0f - sg, 1.0f) * zd[j];
      }
      store_chunk(smem + dst, rowP, cb >> 3, z);
      store_chunk(smem + dst, rowP, (cb >> 3) + 1, z + 8);
      store_chunk(smem + dst, rowT, cb >> 3, zd);
      store_chunk(smem + dst, rowT, (cb >> 3) + 1, zd + 8);
    }
  };
  // Backward epilogue of hidden layer l: (hbar; hdotbar) of this lane's 16 columns (from `hb_pair`) and (z; zdot) ->
  //   zbar = hbar phi'(z) + hdotbar phi''(z) zdot (primal row),  zdotbar = hdotbar phi'(z) (tangent row)
  // written as fp16 rows of `dst`; the bias gradient sum over the primal rows of zbar goes to db (shared-memory accumulator).
  auto bwd_epilogue = [&](uint32_t dcol, const float* bias, int dst, float* db, auto&& hb_pair) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = 64 * half + 32 * c, cb = col0 + (primal ? 0 : 16);
      float z[16], zd[16], hp[16], ht[16];
      {
        uint32_t r[32];
        TMEM_LD32(tlane + dcol + col0, r);
        tc_wait_ld();
        pair_split(r, primal, z, zd);
      }
      hb_pair(col0, cb, hp, ht);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float zz = z[j] + (bias != nullptr ? bias[cb + j] : 0.0f);
        const float sg = fmaf(0.5f, tanh_fast(0.5f * zz), 0.5f);
        const float d1 = sg * fmaf(zz, 1.0f - sg, 1.0f);
        const float d2 = sg * (1.0f - sg) * fmaf(zz, 1.0f - 2.0f * sg, 2.0f);
        z[j] = fmaf(hp[j], d1, ht[j] * d2 * zd[j]);
        zd[j] = ht[j] * d1;
      }
      store_chunk(smem + dst, rowP, cb >> 3, z);
      store_chunk(smem + dst, rowP, (cb >> 3) + 1, z + 8);
      store_chunk(smem + dst, rowT, cb >> 3, zd);
      store_chunk(smem + dst, rowT, (cb >> 3) + 1, zd + 8);
      if (db != nullptr) {
        const float sum = warp_pair_sum16(z, lane);  // over this warp's 16 samples, column cb + bits(4..1) of the lane
        atomicAdd(db + cb + (((lane >> 4) & 1) << 3 | ((lane >> 3) & 1) << 2 | ((lane >> 2) & 1) << 1 | ((lane >> 1) & 1)), sum);
      }
    }
  };
  // add this row's 64 columns of a 128 x 128 weight-gradient tile (TMEM) into the CTA's partial buffer
  auto flush_tile = [&](uint32_t dcol, int goff) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int col0 = 64 * half + 32 * c;
      uint32_t r[32];
      TMEM_LD32(tlane + dcol + col0, r);
      tc_wait_ld();
      // fire-and-forget 16-byte reductions into this CTA's private slice (no read latency, no contention)
      float* dst = part + goff + row * 128 + col0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4 * q), "f"(__uint_as_float(r[4 * q])),
                     "f"(__uint_as_float(r[4 * q + 1])), "f"(__uint_as_float(r[4 * q + 2])),
                     "f"(__uint_as_float(r[4 * q + 3]))
                     : "memory");
    }
  };

  const long long ntiles = (P.B + STC_TS - 1) / STC_TS;
  for (long long tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x) {
    const long long b = tile * STC_TS + (row >> 1);
    const bool live = b < P.B;
    float q[DP];     // cotangent direction of adot (used on the tangent rows)
    float gout = 0.0f, half_bv2 = 0.0f;

    // ---- layer-1 operand: (u; udot) of the premodule (NN.py:64-70) and q ----------------------------------------------------
    // The four threads of a sample (primal / tangent lane x two column halves) all load y and v; the half-0 pair writes the
    // layer-1 operand; the d^3 contraction of q is split four ways (i by column half, j by lane parity).
    {
      float yv[DP], vv[DP];
      float r2 = 0.0f, ydotv = 0.0f, v2 = 0.0f;
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        yv[c] = (live && c < d) ? P.y[b * d + c] : ((c == 0 && !live) ? 1.0f : 0.0f);
        vv[c] = (live && c < d) ? P.v[b * d + c] : 0.0f;
        r2 = fmaf(yv[c], yv[c], r2);
        ydotv = fmaf(yv[c], vv[c], ydotv);
        v2 = fmaf(vv[c], vv[c], v2);
      }
      const float s = live ? P.t[b] : 0.0f;
      gout = live ? P.gout[b] * P.cot_scale : 0.0f;
      const float bt = beta_of(P.bmin, P.bdel, s), sb = sqrtf(bt);
      if (KIND == MSGM_SDE_SGM) half_bv2 = 0.5f * bt * v2;
      if (half == 0) {
        float x[K1];
#pragma unroll
        for (int k = 0; k < K1; ++k) x[k] = 0.0f;
        if (P.pre) {
          const float r = sqrtf(r2), rn = r + 1e-6f, rdot = ydotv / r;
#pragma unroll
          for (int c = 0; c < DP; ++c)
            if (c < d) x[c] = primal ? yv[c] / rn : vv[c] / rn - yv[c] * rdot / (rn * rn);
#pragma unroll
          for (int k = 0; k < K1; ++k)  // runtime column d through a select chain: x[] stays in registers
            if (k == d) x[k] = primal ? logf(rn) : rdot / rn;
        } else {
#pragma unroll
          for (int c = 0; c < DP; ++c)
            if (c < d) x[c] = primal ? yv[c] : vv[c];
        }
#pragma unroll
        for (int k = 0; k < K1; ++k) {
          if (k == d + P.pre) x[k] = primal ? s : 0.0f;           // time input
          if (k == d + P.pre + 1) x[k] = primal ? 1.0f : 0.0f;    // ones column: the bias b1 rides on the tensor pipe
        }
#pragma unroll
        for (int ch = 0; ch < K1 / 8; ++ch) store_chunk(smem + L::oIn0, row, ch, x + 8 * ch);
      }
      // q_k: MSGM dense sqrt(beta) sum_ij v_i G_ijk y_j; sparse c sqrt(beta) (v_k y_{k+1} - v_{k+1} y_k); SGM sqrt(beta) v_k
#pragma unroll
      for (int k = 0; k < DP; ++k) q[k] = 0.0f;
      if (KIND == MSGM_SDE_SGM) {
#pragma unroll
        for (int k = 0; k < DP; ++k) q[k] = sb * vv[k];
      } else if (KIND == MSGM_SDE_MSGM_SPARSE) {
#pragma unroll
        for (int k = 0; k < DP; ++k) {
          float yn = 0.0f, vn = 0.0f;
#pragma unroll
          for (int e = 0; e < DP; ++e) {
            const int kn = (k + 1 == d) ? 0 : k + 1;
            yn = (e == kn) ? yv[e] : yn;
            vn = (e == kn) ? vv[e] : vn;
          }
          q[k] = k < d ? SQRT_HALF * sb * (vv[k] * yn - vn * yv[k]) : 0.0f;
        }
      } else {
        const int jpar = primal ? 0 : 1;
#pragma unroll
        for (int i2 = 0; i2 < DP / 2; ++i2) {
#pragma unroll
          for (int j2 = 0; j2 < DP / 2; ++j2) {
            // i = 2 i2 + half, j = 2 j2 + jpar: compile-time register picks, two selects instead of a DP-way chain
            const float vi = half ? vv[2 * i2 + 1] : vv[2 * i2];
            const float yj = jpar ? yv[2 * j2 + 1] : yv[2 * j2];
            const int i = 2 * i2 + half, j = 2 * j2 + jpar;
            if (i < d && j < d) {
              const float w = sb * vi * yj;
              const float* g = sG + (i * d + j) * d;
#pragma unroll
              for (int k = 0; k < DP; ++k)
                if (k < d) q[k] = fmaf(w, g[k], q[k]);
            }
          }
        }
#pragma unroll
        for (int k = 0; k < DP; ++k) q[k] += __shfl_xor_sync(0xffffffffu, q[k], 1);  // the two j halves
        if (half == 1) {
#pragma unroll
          for (int k = 0; k < DP; ++k) sQ[row * DP + k] = q[k];  // the other i half: added by the half-0 thread in the loss phase
        }
      }
    }
    publish();
    tick(0);

    // ---- forward: three hidden layers and the output layer ---------------------------------------------------------------
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < K1 / 16; ++s) umma_ss(tbase + D1, kdesc(L::oIn0, s, 128), kdesc(L::oW1, s, 128), ID_H, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(1);
    fwd_epilogue(D1, nullptr, L::oIn1);
    publish();
    tick(2);
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s) umma_ss(tbase + D2, kdesc(L::oIn1, s, 128), kdesc(L::oW2, s, 128), ID_H, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(3);
    fwd_epilogue(D2, sB2, L::oIn2);
    publish();
    tick(4);
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s) umma_ss(tbase + D3, kdesc(L::oIn2, s, 128), kdesc(L::oW3, s, 128), ID_H, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(5);
    fwd_epilogue(D3, sB2 + 128, L::oIn3);
    publish();
    tick(6);
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s) umma_ss(tbase + WK, kdesc(L::oIn3, s, 128), kdesc(L::oW4, s, 16), ID_O, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(7);

    // ---- loss and output cotangents: abar = gout a (primal rows), adotbar = gout q (tangent rows) ------------------------------
    if (half == 0) {
      uint32_t r[16];
      TMEM_LD16(tlane + WK, r);
      tc_wait_ld();
      if (KIND == MSGM_SDE_MSGM_DENSE) {
#pragma unroll
        for (int k = 0; k < DP; ++k) q[k] += sQ[row * DP + k];
      }
      float lp = 0.0f, x[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        float a = 0.0f;
        if (c < DP) {
          a = c < d ? __uint_as_float(r[c]) + (primal ? sB2[256 + c] : 0.0f) : 0.0f;   // a (primal) / adot (tangent)
          lp = primal ? fmaf(0.5f * a, a, lp) : fmaf(q[c], a, lp);
          x[c] = gout * (primal ? a : q[c]);
          sX[row * DP + c] = x[c];
          // bias gradient of the output layer: sum of abar over the warp's primal rows, one shared-memory atomic per warp
          float sabar = primal ? x[c] : 0.0f;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) sabar += __shfl_xor_sync(0xffffffffu, sabar, off);
          if (lane == 0 && c < d) atomicAdd(sDb + 256 + c, sabar);
        } else {
          x[c] = 0.0f;
        }
      }
      const float lt = __shfl_xor_sync(0xffffffffu, lp, 1);
      if (primal && live) P.loss[b] = lp + lt + half_bv2;
      store_chunk(smem + L::oC4, row, 0, x);
      store_chunk(smem + L::oC4, row, 1, x + 8);
    }
    publish();
    tick(8);

    // ---- backward -----------------------------------------------------------------------------------------------------------
    // layer 4 weight gradient, transposed: gW4^T [in 128][out 16] = H3^T Cot4 (both MN-major, K = rows)
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_ss(tbase + WK, mdesc(L::oIn3, s, 128), mdesc(L::oC4, s, 128), ID_O | IDESC_A_MN | IDESC_B_MN, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(9);
    if (half == 0) {  // row = input feature of W4
      uint32_t r[16];
      TMEM_LD16(tlane + WK, r);
      tc_wait_ld();
#pragma unroll
      for (int o = 0; o < 16; ++o)
        if (o < d) atomicAdd(part + gW4 + o * 128 + row, __uint_as_float(r[o]));  // result unused: a fire-and-forget RED
    }
    {  // layer 3: (hbar; hdotbar) = Cot4 W4 on the CUDA cores (K = d is tiny), cotangents into the tile of H3
      float cp[DP], ct[DP];
#pragma unroll
      for (int c = 0; c < DP; ++c) {
        cp[c] = sX[rowP * DP + c];
        ct[c] = sX[rowT * DP + c];
      }
      bwd_epilogue(D3, sB2 + 128, L::oIn3, sDb + 128, [&](int, int cb, float (&hp)[16], float (&ht)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j) hp[j] = ht[j] = 0.0f;
        for (int o = 0; o < d; ++o) {
          float a = 0.0f, bq = 0.0f;
#pragma unroll
          for (int e = 0; e < DP; ++e) {
            a = (e == o) ? cp[e] : a;
            bq = (e == o) ? ct[e] : bq;
          }
          const float4* w = reinterpret_cast<const float4*>(sW4f + o * 128 + cb);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const float4 ww = w[j4];
            hp[4 * j4] = fmaf(a, ww.x, hp[4 * j4]);         ht[4 * j4] = fmaf(bq, ww.x, ht[4 * j4]);
            hp[4 * j4 + 1] = fmaf(a, ww.y, hp[4 * j4 + 1]); ht[4 * j4 + 1] = fmaf(bq, ww.y, ht[4 * j4 + 1]);
            hp[4 * j4 + 2] = fmaf(a, ww.z, hp[4 * j4 + 2]); ht[4 * j4 + 2] = fmaf(bq, ww.z, ht[4 * j4 + 2]);
            hp[4 * j4 + 3] = fmaf(a, ww.w, hp[4 * j4 + 3]); ht[4 * j4 + 3] = fmaf(bq, ww.w, ht[4 * j4 + 3]);
          }
        }
      });
    }
    publish();
    tick(10);
    for (int l = 3; l >= 2; --l) {
      // gW_l = Cot_l^T H_{l-1} into layer l's own (dead) TMEM columns; Hbar_{l-1} = Cot_l W_l into the working columns
      const int cotT = l == 3 ? L::oIn3 : L::oIn2, inT = l == 3 ? L::oIn2 : L::oIn1, wT = l == 3 ? L::oW3 : L::oW2;
      const uint32_t dl = l == 3 ? D3 : D2;
      if (warp == 0) {
#pragma unroll
        for (int s = 0; s < 8; ++s)
          umma_ss(tbase + dl, mdesc(cotT, s, 128), mdesc(inT, s, 128), ID_H | IDESC_A_MN | IDESC_B_MN, s > 0, 0);
#pragma unroll
        for (int s = 0; s < 8; ++s) umma_ss(tbase + WK, kdesc(cotT, s, 128), mdesc(wT, s, 128), ID_H | IDESC_B_MN, s > 0, 0);
        umma_commit(bar, 0);
      }
      wait_mma();
      tick(l == 3 ? 11 : 13);
      flush_tile(dl, l == 3 ? gW3 : gW2);
      tick(l == 3 ? 18 : 19);
      auto hb_tmem = [&](int col0, int, float (&hp)[16], float (&ht)[16]) {
        uint32_t r[32];
        TMEM_LD32(tlane + WK + col0, r);
        tc_wait_ld();
        pair_split(r, primal, hp, ht);
      };
      if (l == 3) bwd_epilogue(D2, sB2, L::oIn2, sDb, hb_tmem);
      else bwd_epilogue(D1, nullptr, L::oIn1, nullptr, hb_tmem);  // b1 is inside D1; its gradient comes out of gW1's ones column
      publish();
      tick(l == 3 ? 12 : 14);
    }
    // layer 1: gW1 [out 128][K1] = Cot1^T (u; udot); column K1ref is the bias gradient
    if (warp == 0) {
#pragma unroll
      for (int s = 0; s < 8; ++s)
        umma_ss(tbase + D1, mdesc(L::oIn1, s, 128), mdesc(L::oIn0, s, 128), ID_1 | IDESC_A_MN | IDESC_B_MN, s > 0, 0);
      umma_commit(bar, 0);
    }
    wait_mma();
    tick(15);
    if (half == 0) {
#pragma unroll
      for (int c = 0; c < K1 / 16; ++c) {
        uint32_t r[16];
        TMEM_LD16(tlane + D1 + 16 * c, r);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int k = 16 * c + j;
          if (k < K1ref) atomicAdd(part + gW1 + row * K1ref + k, __uint_as_float(r[j]));
          else if (k == K1ref) atomicAdd(part + gb1 + row, __uint_as_float(r[j]));
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // every TMEM read of this tile is done before the next tile's products overwrite the columns
    tc_fence_after();
    tick(16);
  }

  // ---- bias gradients of layers 2-4 accumulated in shared memory over the CTA's tiles ---------------------------------------------
  __syncthreads();
  for (int e = tid; e < 256; e += STC_THREADS) part[(e < 128 ? gb2 : gb3) + (e & 127)] = sDb[e];
  for (int e = tid; e < d; e += STC_THREADS) part[gb4 + e] = sDb[256 + e];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

__global__ void __launch_bounds__(256) ssm_tc_reduce_kernel(const float* __restrict__ part, int nparts, int nparam,
                                                            int pstride, float inv_scale, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nparam) return;
  float s = 0.0f;
  for (int p = 0; p < nparts; ++p) s += part[(size_t)p * pstride + i];
  grad[i] = s * inv_scale;
}

// ---- host side ---------------------------------------------------------------------------------------------------------
int ssm_tc_grid(const msgm_ctx* ctx, long long B) {
  return (int)std::min<long long>((B + STC_TS - 1) / STC_TS, ctx->num_sms);
}

template <int K1, int DP, int KIND>
static int launch_stc(msgm_ctx* ctx, StcParams& P, float* grad_flat, cudaStream_t stream) {
  using L = StcLayout<K1, DP>;
  auto kern = ssm_tc_kernel<K1, DP, KIND>;
  MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
  const int grid = ssm_tc_grid(ctx, P.B);
  P.flags = next_tc_flags(ctx);
  P.prof = std::getenv("MSGM_TC_PROF") ? reinterpret_cast<long long*>(reinterpret_cast<unsigned char*>(ctx->ws) + 64) : nullptr;
  if (P.prof) MSGM_CUDA_TRY(cudaMemsetAsync(P.prof, 0, 192, stream));
  kern<<<grid, STC_THREADS, L::SMEM, stream>>>(P);
  ssm_tc_reduce_kernel<<<(P.nparam + 255) / 256, 256, 0, stream>>>(P.part, grid, P.nparam, P.pstride, 1.0f / P.cot_scale, grad_flat);
  ctx->launches += 2;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int K1, int DP>
static int launch_stc_kind(msgm_ctx* ctx, int kind, StcParams& P, float* grad_flat, cudaStream_t stream) {
  switch (kind) {
    case MSGM_SDE_SGM: return launch_stc<K1, DP, MSGM_SDE_SGM>(ctx, P, grad_flat, stream);
    case MSGM_SDE_MSGM_DENSE: return launch_stc<K1, DP, MSGM_SDE_MSGM_DENSE>(ctx, P, grad_flat, stream);
    case MSGM_SDE_MSGM_SPARSE: return launch_stc<K1, DP, MSGM_SDE_MSGM_SPARSE>(ctx, P, grad_flat, stream);
  }
  set_error("unknown sde kind");
  return MSGM_ERR_INVALID;
}

int ssm_fwd_bwd_tc(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y, const float* v,
                   const float* t, const float* gout, float* loss, float* grad_flat, float* partials, float cot_scale,
                   int64_t B, cudaStream_t stream) {
  const int d = sde->dim;
  if (d > 16) {
    set_error("f16tc SSM step is built for d <= 16; use the fp32 kernels");
    return MSGM_ERR_UNSUPPORTED;
  }
  StcParams P{};
  P.d = d;
  P.pre = mlp->premodule;
  P.kind = sde->kind;
  P.bmin = sde->beta_min;
  P.bdel = sde->beta_delta;
  P.G = sde->G;
  for (int l = 0; l < 4; ++l) { P.W[l] = mlp->W[l]; P.b[l] = mlp->b[l]; }
  P.y = y; P.v = v; P.t = t; P.gout = gout;
  P.loss = loss;
  P.part = partials;
  const int K1ref = d + 1 + mlp->premodule;
  P.nparam = 128 * K1ref + 128 + 2 * (128 * 128 + 128) + 128 * d + d;
  P.pstride = (P.nparam + 3) & ~3;
  P.cot_scale = cot_scale;
  P.B = B;
  // layer-1 operand width: inputs + the ones column, rounded up to the MMA's K = 16
  if (d <= 8) return launch_stc_kind<16, 8>(ctx, sde->kind, P, grad_flat, stream);   // K1ref + 1 <= 11
  return launch_stc_kind<32, 16>(ctx, sde->kind, P, grad_flat, stream);              // K1ref + 1 <= 19
}

}  // namespace msgm
