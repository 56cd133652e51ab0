// fp32 "parity mode" sampler: the whole EM / Heun / RK4-Stratonovich loop of sde_scheme.py:43-269 for an
// MLP score net (NN.py:73-120) in ONE persistent launch.  Everything is CUDA-core fp32 (the reference is fp32
// end to end); the tensor-core path lives in sampler_tc.cu.
//
// Work decomposition: one CTA (256 threads) owns a tile of 64 particles for all N steps.  The two 128x128
// hidden layers stay resident in shared memory (128 KB) for the CTA's lifetime, the particle state
// (x, RK sum, dW, r0) stays in registers, so a particle touches HBM once at the start and once at the end
// (plus 4d bytes/step when a trajectory is captured, plus 4d bytes/step when noise is injected).
//
// Per stage the CTA runs: premodule -> 3 register-tiled (4 particles x 8 features per thread) dense layers
// with Swish -> output layer -> w = c_a a + c_w dW -> K = g(s,y) . w (+ Ito/lambda term) in registers.
#include <algorithm>
#include <cmath>

#include "msgm_common.cuh"

namespace msgm {

constexpr int TP = 64;         // particles per CTA tile
constexpr int NTHREADS = 256;  // 16 feature groups x 16 particle groups; also 64 particles x 4 component lanes

struct SampleParams {
  // sde
  int d;
  float bmin, bdel, Tsde;
  const float* G;   // dense: (d,d,d) unpadded, or padded (32,32,32) workspace copy when DP == 32
  const float* LG;  // dense: (d,d)
  // net
  int pre;
  const float* W[4];
  const float* b[4];
  // run
  int scheme, N, nc, inc_t0, fwd;
  float lmbd;
  float delta, delta_half, sqrt_delta;
  const float* ts;
  const float* noise;
  unsigned long long seed, poff;
  float* traj;
  const int* keep_step;
  float* keep_out;
  const float* T_rows;
  const float* t_noise;       // noising mode: per-row noise time t_k (SDEs.py:78-122), see msgm_noise_forward
  const float* noise_single;  // noising mode: (B,d) normals of the one-step rows, or NULL
  float* x;
  long long B;
};

// XOR-swizzled [row][64] activation layout: 16-byte chunks of a row are permuted by the row index so that the
// transposing epilogue store (lanes = different rows, same chunk) is bank-conflict free.
__device__ __forceinline__ int act_idx(int row, int p) {
  return row * TP + ((((p >> 2) ^ (row >> 2)) & 15) << 2) + (p & 3);
}

__device__ __forceinline__ float swishf(float z) { return z / (1.0f + expf(-z)); }  // NN.py:52-53

// out[f][p] = act( bias[f] + sum_k in[k][p] * Wt[k][f] ), f < 128, p < 64.  `in` and `out` may alias.
// Thread (ng, pg) owns features {4ng..4ng+3, 64+4ng..64+4ng+3} x particles {4pg..4pg+3}.
__device__ __forceinline__ void dense128(const float* __restrict__ Wt, const float* in, int K,
                                         const float* __restrict__ bias, float* out, int ng, int pg) {
  float acc[4][8];
  {
    float4 b0 = *reinterpret_cast<const float4*>(bias + ng * 4);
    float4 b1 = *reinterpret_cast<const float4*>(bias + 64 + ng * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[i][0] = b0.x; acc[i][1] = b0.y; acc[i][2] = b0.z; acc[i][3] = b0.w;
      acc[i][4] = b1.x; acc[i][5] = b1.y; acc[i][6] = b1.z; acc[i][7] = b1.w;
    }
  }
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(in + k * TP + (((pg ^ (k >> 2)) & 15) << 2));
    const float4 w0 = *reinterpret_cast<const float4*>(Wt + k * HID + ng * 4);
    const float4 w1 = *reinterpret_cast<const float4*>(Wt + k * HID + 64 + ng * 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
  __syncthreads();  // every thread is done reading `in` (it may alias `out`)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int f = ng * 4 + (j & 3) + 64 * (j >> 2);
    float4 o = make_float4(swishf(acc[0][j]), swishf(acc[1][j]), swishf(acc[2][j]), swishf(acc[3][j]));
    *reinterpret_cast<float4*>(out + f * TP + (((pg ^ (f >> 2)) & 15) << 2)) = o;
  }
  __syncthreads();
}

template <int DP>
struct SmemLayout {
  static constexpr int K1MAX = DP + 2;
  static constexpr int oWh = 0;                          // [2][128][128]
  static constexpr int oAct = oWh + 2 * HID * HID;       // [128][64] swizzled
  static constexpr int oW1 = oAct + HID * TP;            // [K1MAX][128]
  static constexpr int oW4 = oW1 + K1MAX * HID;          // [128][DP]
  static constexpr int oB = oW4 + HID * DP;              // b1,b2,b3 [3][128], b4 [32]
  static constexpr int oIn = oB + 3 * HID + 32;          // [K1MAX rounded to 4][64] swizzled
  static constexpr int oY = oIn + ((K1MAX + 3) & ~3) * TP;  // [DP][64]
  static constexpr int oWv = oY + DP * TP;               // [DP][64]
  static constexpr int oLG = oWv + DP * TP;              // [DP][DP]
  static constexpr int oG = oLG + DP * DP;               // [DP][DP][DP] when DP <= 16
  static constexpr int total = oG + (DP <= 16 ? DP * DP * DP : 0);
  static constexpr size_t bytes = sizeof(float) * (size_t)total;
};

template <int DP, int KIND>
__global__ void __launch_bounds__(NTHREADS, 1) sample_fp32_kernel(const __grid_constant__ SampleParams P) {
  using L = SmemLayout<DP>;
  constexpr int NC = (DP + 3) / 4;  // components per state thread
  extern __shared__ __align__(16) float smem[];
  float* sWh = smem + L::oWh;
  float* sAct = smem + L::oAct;
  float* sW1 = smem + L::oW1;
  float* sW4 = smem + L::oW4;
  float* sB = smem + L::oB;
  float* sIn = smem + L::oIn;
  float* sY = smem + L::oY;
  float* sWv = smem + L::oWv;
  float* sLG = smem + L::oLG;
  float* sG = smem + L::oG;

  const int tid = threadIdx.x;
  const int ng = tid & 15, pg = tid >> 4;
  const int sp = tid & 63, c0 = tid >> 6;
  const int d = P.d;
  const int K1 = d + 1 + P.pre;
  const bool fwd = P.fwd != 0;

  // ---- one-time: stage weights into shared memory (torch Linear layout (out,in) -> [in][out]) -------------
  if (!fwd) {
    for (int l = 0; l < 2; ++l) {
      const float* W = P.W[1 + l];
      for (int e = tid; e < HID * HID; e += NTHREADS) {  // e = k*128 + n ; lanes along n
        int k = e >> 7, n = e & 127;
        sWh[l * HID * HID + e] = __ldg(W + n * HID + k);
      }
    }
    for (int e = tid; e < K1 * HID; e += NTHREADS) {
      int k = e >> 7, n = e & 127;
      sW1[e] = __ldg(P.W[0] + n * K1 + k);
    }
    for (int e = tid; e < HID * DP; e += NTHREADS) {
      int k = e / DP, c = e % DP;
      sW4[e] = c < d ? __ldg(P.W[3] + c * HID + k) : 0.0f;
    }
    for (int e = tid; e < 3 * HID; e += NTHREADS) sB[e] = __ldg(P.b[e >> 7] + (e & 127));
    if (tid < 32) sB[3 * HID + tid] = tid < d ? __ldg(P.b[3] + tid) : 0.0f;
  }
  if (KIND == MSGM_SDE_MSGM_DENSE) {
    for (int e = tid; e < DP * DP; e += NTHREADS) {
      int i = e / DP, j = e % DP;
      sLG[e] = (i < d && j < d) ? __ldg(P.LG + i * d + j) : 0.0f;
    }
    if (DP <= 16) {
      for (int e = tid; e < DP * DP * DP; e += NTHREADS) {
        int i = e / (DP * DP), j = (e / DP) % DP, k = e % DP;
        sG[e] = (i < d && j < d && k < d) ? __ldg(P.G + (i * d + j) * d + k) : 0.0f;
      }
    }
  }
  __syncthreads();

  // ---- constants of the lambda family (SDEs.py:561,584,588) ------------------------------------------------
  const float lm = P.lmbd;
  const float c_w = fwd ? 1.0f : sqrtf(1.0f - lm);      // weight of dW in w
  const bool ito = (P.scheme == MSGM_SCHEME_EM);
  // coefficient of f(s,y) in the stage drift: reverse Strato -lambda, reverse Ito (1-2 lambda);
  // forward adapter Strato 0 (f_strato = 0 for MSGM), forward Ito +1 (SDEs.py:38-43).
  const float c_f = fwd ? (ito ? 1.0f : 0.0f) : (ito ? (1.0f - 2.0f * lm) : -lm);
  const int nstage = P.scheme == MSGM_SCHEME_RK4 ? 4 : (P.scheme == MSGM_SCHEME_HEUN ? 2 : 1);

  const long long ntiles = (P.B + TP - 1) / TP;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long gp = tile * TP + sp;  // global particle row
    const bool live = gp < P.B;
    float x[NC], y[NC], ks[NC], dw[NC], K[NC];
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + 4 * m;
      x[m] = (live && c < d) ? P.x[gp * d + c] : ((c == 0 && !live) ? 1.0f : 0.0f);
      y[m] = x[m];
      ks[m] = 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NC; ++m)
      if (c0 + 4 * m < DP) sY[(c0 + 4 * m) * TP + sp] = x[m];
    __syncthreads();
    float r0 = 0.0f;
    if (P.nc) {
      for (int j = 0; j < d; ++j) r0 = fmaf(sY[j * TP + sp], sY[j * TP + sp], r0);
      r0 = sqrtf(r0);  // torch.norm(x_t, dim=1), sde_scheme.py:66,124,205
    }
    if (P.traj && P.inc_t0 && live) {
#pragma unroll
      for (int m = 0; m < NC; ++m)
        if (c0 + 4 * m < d) P.traj[gp * d + c0 + 4 * m] = x[m];
    }
    const int keep = (P.keep_step && live) ? P.keep_step[gp] : -1;
    // step size: a Python double T_/N in the reference, rounded to fp32 where it meets tensors
    float delta = P.delta, delta_half = P.delta_half, sqrt_delta = P.sqrt_delta, Trow = 1.0f;
    if (P.T_rows) {
      Trow = live ? P.T_rows[gp] : 1.0f;
      const double dd = (double)Trow / (double)P.N;
      delta = (float)dd;
      delta_half = (float)(dd * 0.5);
      sqrt_delta = (float)sqrt(dd);
    }
    // noising mode (SDE.sample_scheme, SDEs.py:86-118): row k stops after n_k = trunc(N t_k / T) steps of the common
    // grid; a row with n_k == 0 takes ONE step of size t_k instead (the reference's per-row sampler call with T_ = t[k]).
    int my_steps = P.N;
    bool single = false;
    if (P.t_noise) {
      const float tk = live ? P.t_noise[gp] : 0.0f;
      const int nk = tk >= P.Tsde ? P.N : (int)truncf(__fdiv_rn(__fmul_rn((float)P.N, tk), P.Tsde));
      single = nk == 0;
      my_steps = single ? 1 : nk;
      if (single) {
        delta = tk;                       // T_/1 with T_ = t_k
        delta_half = (float)((double)tk * 0.5);
        sqrt_delta = (float)sqrt((double)tk);
      }
    }
    const float c_a = delta * (1.0f - 0.5f * lm);  // weight of a in w

    for (int step = 0; step < P.N; ++step) {
      const bool active = step < my_steps;
      const float tcur = single ? 0.0f
                                : (P.T_rows ? __fmul_rn(__ldg(P.ts + step), Trow)
                                            : (P.ts ? __ldg(P.ts + step) : __fmul_rn((float)step, delta)));
      // ---- Wiener increment, shared by all stages of the step (sde_scheme.py:227) --------------------------
#pragma unroll
      for (int m = 0; m < NC; ++m) {
        const int c = c0 + 4 * m;
        float xi = 0.0f;
        if (c < d) {
          if (single && P.noise_single) {
            xi = (live && step == 0) ? __ldg(P.noise_single + gp * d + c) : 0.0f;
          } else if (P.noise && !(single && P.t_noise)) {
            xi = live ? __ldg(P.noise + ((long long)step * P.B + gp) * d + c) : 0.0f;
          } else {
            // one-step rows draw from their own stream id so that they do not reuse step 0 of the common grid
            float4 z = philox_normal4(P.seed, P.poff + (unsigned long long)gp, single ? 0xFFFF0002u : (uint32_t)step, (uint32_t)m);
            xi = c0 == 0 ? z.x : (c0 == 1 ? z.y : (c0 == 2 ? z.z : z.w));
          }
        }
        dw[m] = sqrt_delta * xi;
      }

      for (int st = 0; st < nstage; ++st) {
        // stage time: t, t+delta/2, t+delta/2, t+delta (RK4) | t, t+delta (Heun) | t (EM)
        float tst = tcur;
        if (st > 0) tst = (nstage == 4 && st < 3) ? __fadd_rn(tcur, delta_half) : __fadd_rn(tcur, delta);
        const float sv = fwd ? tst : __fsub_rn(P.Tsde, tst);  // reverse SDE runs in s = T - t (SDEs.py:557)
        const float bt = beta_of(P.bmin, P.bdel, sv);
        const float sb = sqrtf(bt);

        if (st > 0) {
          __syncthreads();  // contraction of the previous stage is done reading sY / sWv
#pragma unroll
          for (int m = 0; m < NC; ++m)
            if (c0 + 4 * m < DP) sY[(c0 + 4 * m) * TP + sp] = y[m];
          __syncthreads();
        }

        float a[NC] = {};
        if (!fwd) {
          // ---- premodule + layer-1 operand (NN.py:64-70,115-118) ------------------------------------------
          float rn = 1.0f, lognorm = 0.0f;
          if (P.pre) {
            float r = 0.0f;
            for (int j = 0; j < d; ++j) r = fmaf(sY[j * TP + sp], sY[j * TP + sp], r);
            rn = sqrtf(r) + 1e-6f;
            lognorm = logf(rn);
          }
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            const int c = c0 + 4 * m;
            if (c < d) sIn[act_idx(c, sp)] = P.pre ? y[m] / rn : y[m];
          }
          if (c0 == 0) {
            if (P.pre) sIn[act_idx(d, sp)] = lognorm;
            sIn[act_idx(d + P.pre, sp)] = sv;
          }
          __syncthreads();
          dense128(sW1, sIn, K1, sB, sAct, ng, pg);
          dense128(sWh, sAct, HID, sB + HID, sAct, ng, pg);
          dense128(sWh + HID * HID, sAct, HID, sB + 2 * HID, sAct, ng, pg);
          // ---- output layer 128 -> d: thread (sp, c0) produces its own components -------------------------
#pragma unroll
          for (int m = 0; m < NC; ++m) a[m] = sB[3 * HID + ((c0 + 4 * m) & 31)];
#pragma unroll 4
          for (int k = 0; k < HID; ++k) {
            const float h = sAct[act_idx(k, sp)];
#pragma unroll
            for (int m = 0; m < NC; ++m)
              if (c0 + 4 * m < DP) a[m] = fmaf(h, sW4[k * DP + c0 + 4 * m], a[m]);
          }
        }

        // ---- stage increment K = delta * drift + sigma . dW -------------------------------------------------
        if (KIND == MSGM_SDE_SGM) {
#pragma unroll
          for (int m = 0; m < NC; ++m) {
            if (fwd)  // f_strato = f = -beta y / 2, g = sqrt(beta) (SDEs.py:183-194)
              K[m] = delta * (-0.5f * bt * y[m]) + sb * dw[m];
            else      // mu = (1 - lambda/2) sqrt(beta) a + beta y / 2 ; sigma = sqrt(1-lambda) sqrt(beta)
              K[m] = delta * ((1.0f - 0.5f * lm) * (sb * a[m]) + 0.5f * bt * y[m]) + (c_w * sb) * dw[m];
          }
        } else {
          // w = delta (1 - lambda/2) a + sqrt(1-lambda) dW, so that K = g(s,y) . w + delta c_f f(s,y)
#pragma unroll
          for (int m = 0; m < NC; ++m)
            if (c0 + 4 * m < DP) sWv[(c0 + 4 * m) * TP + sp] = fwd ? dw[m] : fmaf(c_a, a[m], c_w * dw[m]);
          __syncthreads();
          if (KIND == MSGM_SDE_MSGM_SPARSE) {
            // cyclic stencil of SDEs.py:369-399 / 427-430 / sde_scheme.py:27-32
#pragma unroll
            for (int m = 0; m < NC; ++m) {
              const int c = c0 + 4 * m;
              float acc = 0.0f;
              if (c < d) {
                const int cn = (c + 1 == d) ? 0 : c + 1, cp = (c == 0) ? d - 1 : c - 1;
                const float t1 = (SQRT_HALF * (sb * sY[cn * TP + sp])) * sWv[c * TP + sp];
                const float t2 = (-SQRT_HALF * (sb * sY[cp * TP + sp])) * sWv[cp * TP + sp];
                acc = t1 + t2;
                acc = fmaf(delta * c_f, 0.5f * bt * y[m], acc);  // sparse f = +beta y / 2 (SDEs.py:412-413)
              }
              K[m] = acc;
            }
          } else {
            float yr[DP];
#pragma unroll
            for (int j = 0; j < DP; ++j) yr[j] = sY[j * TP + sp];
#pragma unroll
            for (int m = 0; m < NC; ++m) {
              const int c = c0 + 4 * m;
              float acc = 0.0f, fc = 0.0f;
              if (c < DP) {
                if constexpr (DP >= 4) {
                  const float* Gc = (DP <= 16 ? sG : P.G) + c * DP * DP;
#pragma unroll 1
                  for (int k4 = 0; k4 < DP; k4 += 4) {
                    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < DP; ++j) {
                      const float4 g4 = DP <= 16 ? *reinterpret_cast<const float4*>(Gc + j * DP + k4)
                                                 : __ldg(reinterpret_cast<const float4*>(Gc + j * DP + k4));
                      u.x = fmaf(g4.x, yr[j], u.x);
                      u.y = fmaf(g4.y, yr[j], u.y);
                      u.z = fmaf(g4.z, yr[j], u.z);
                      u.w = fmaf(g4.w, yr[j], u.w);
                    }
                    acc = fmaf(u.x, sWv[(k4 + 0) * TP + sp], acc);
                    acc = fmaf(u.y, sWv[(k4 + 1) * TP + sp], acc);
                    acc = fmaf(u.z, sWv[(k4 + 2) * TP + sp], acc);
                    acc = fmaf(u.w, sWv[(k4 + 3) * TP + sp], acc);
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < DP; ++k) {
                    float u = 0.0f;
#pragma unroll
                    for (int j = 0; j < DP; ++j) u = fmaf(sG[(c * DP + j) * DP + k], yr[j], u);
                    acc = fmaf(u, sWv[k * TP + sp], acc);
                  }
                }
                if (c_f != 0.0f) {
#pragma unroll
                  for (int j = 0; j < DP; ++j) fc = fmaf(sLG[c * DP + j], yr[j], fc);
                }
              }
              K[m] = fmaf(delta * c_f, bt * fc, sb * acc);
            }
          }
        }

        // ---- Runge-Kutta bookkeeping (sde_scheme.py:86 | 147,156 | 232-253) ----------------------------------
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          if (!active) K[m] = 0.0f;  // noising mode: this row has already reached its noise time
          if (nstage == 1) {
            x[m] = x[m] + K[m];
          } else if (nstage == 2) {
            if (st == 0) { ks[m] = K[m]; y[m] = x[m] + K[m]; }
            else { x[m] = x[m] + (ks[m] + K[m]) / 2.0f; }
          } else {
            if (st == 0) { ks[m] = K[m]; y[m] = x[m] + K[m] / 2.0f; }
            else if (st == 1) { ks[m] = ks[m] + 2.0f * K[m]; y[m] = x[m] + K[m] / 2.0f; }
            else if (st == 2) { ks[m] = ks[m] + 2.0f * K[m]; y[m] = x[m] + K[m]; }
            else { x[m] = x[m] + (ks[m] + K[m]) / 6.0f; }
          }
        }
      }  // stages

      // ---- end of step: radius re-pin, capture ------------------------------------------------------------
      __syncthreads();
#pragma unroll
      for (int m = 0; m < NC; ++m)
        if (c0 + 4 * m < DP) sY[(c0 + 4 * m) * TP + sp] = x[m];
      __syncthreads();
      if (P.nc) {
        float r = 0.0f;
        for (int j = 0; j < d; ++j) r = fmaf(sY[j * TP + sp], sY[j * TP + sp], r);
        const float sc = r0 / sqrtf(r);  // sde_scheme.py:86,158,255
        __syncthreads();
#pragma unroll
        for (int m = 0; m < NC; ++m) {
          x[m] = x[m] * sc;
          if (c0 + 4 * m < DP) sY[(c0 + 4 * m) * TP + sp] = x[m];
        }
        __syncthreads();
      }
#pragma unroll
      for (int m = 0; m < NC; ++m) y[m] = x[m];
      if (P.traj) {
        // coalesced copy of the tile's (64 x d) block: sY is [c][p], the global block is [p][c]
        float* dst = P.traj + ((long long)(step + P.inc_t0) * P.B + tile * TP) * d;
        const long long nvalid = min((long long)TP, P.B - tile * TP) * d;
        for (int e = tid; e < nvalid; e += NTHREADS) dst[e] = sY[(e % d) * TP + e / d];
      }
      if (keep >= 0 && keep == step + P.inc_t0) {
#pragma unroll
        for (int m = 0; m < NC; ++m)
          if (c0 + 4 * m < d) P.keep_out[gp * d + c0 + 4 * m] = x[m];
      }
    }  // steps

    if (live) {
#pragma unroll
      for (int m = 0; m < NC; ++m)
        if (c0 + 4 * m < d) P.x[gp * d + c0 + 4 * m] = x[m];
    }
  }  // tiles
}

// Zero-padded (32,32,32) copy of a dense G with 16 < d <= 32 (read through L1/L2 by the DP == 32 instantiation).
__global__ void pad_G32_kernel(const float* __restrict__ G, int d, float* __restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 32 * 32 * 32) return;
  int i = e >> 10, j = (e >> 5) & 31, k = e & 31;
  out[e] = (i < d && j < d && k < d) ? G[(i * d + j) * d + k] : 0.0f;
}

// ---- stand-alone score net forward: a(y, s), NN.py:108-120 ------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_forward_fp32_kernel(int d, int pre, const float* W0, const float* W1, const float* W2, const float* W3,
                        const float* b0, const float* b1, const float* b2, const float* b3,
                        const float* __restrict__ yin, const float* __restrict__ sin_, float* __restrict__ out,
                        long long B) {
  using L = SmemLayout<DP>;
  extern __shared__ __align__(16) float smem[];
  float* sWh = smem + L::oWh;
  float* sAct = smem + L::oAct;
  float* sW1 = smem + L::oW1;
  float* sW4 = smem + L::oW4;
  float* sB = smem + L::oB;
  float* sIn = smem + L::oIn;
  float* sY = smem + L::oY;
  const int tid = threadIdx.x, ng = tid & 15, pg = tid >> 4, sp = tid & 63, c0 = tid >> 6;
  const int K1 = d + 1 + pre;
  const float* Wh[2] = {W1, W2};
  for (int l = 0; l < 2; ++l)
    for (int e = tid; e < HID * HID; e += NTHREADS) sWh[l * HID * HID + e] = __ldg(Wh[l] + (e & 127) * HID + (e >> 7));
  for (int e = tid; e < K1 * HID; e += NTHREADS) sW1[e] = __ldg(W0 + (e & 127) * K1 + (e >> 7));
  for (int e = tid; e < HID * DP; e += NTHREADS) sW4[e] = (e % DP) < d ? __ldg(W3 + (e % DP) * HID + e / DP) : 0.0f;
  for (int e = tid; e < HID; e += NTHREADS) { sB[e] = b0[e]; sB[HID + e] = b1[e]; sB[2 * HID + e] = b2[e]; }
  if (tid < 32) sB[3 * HID + tid] = tid < d ? b3[tid] : 0.0f;
  __syncthreads();
  constexpr int NC = (DP + 3) / 4;
  const long long ntiles = (B + TP - 1) / TP;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long gp = tile * TP + sp;
    const bool live = gp < B;
    float y[NC];
    __syncthreads();
#pragma unroll
    for (int m = 0; m < NC; ++m) {
      const int c = c0 + 4 * m;
      y[m] = (live && c < d) ? yin[gp * d + c] : ((c == 0 && !live) ? 1.0f : 0.0f);
      if (c < DP) sY[c * TP + sp] = y[m];
    }
    __syncthreads();
    float rn = 1.0f;
    if (pre) {
      float r = 0.0f;
      for (int j = 0; j < d; ++j) r = fmaf(sY[j * TP + sp], sY[j * TP + sp], r);
      rn = sqrtf(r) + 1e-6f;
    }
#pragma unroll
    for (int m = 0; m < NC; ++m)
      if (c0 + 4 * m < d) sIn[act_idx(c0 + 4 * m, sp)] = pre ? y[m] / rn : y[m];
    if (c0 == 0) {
      if (pre) sIn[act_idx(d, sp)] = logf(rn);
      sIn[act_idx(d + pre, sp)] = live ? sin_[gp] : 0.0f;
    }
    __syncthreads();
    dense128(sW1, sIn, K1, sB, sAct, ng, pg);
    dense128(sWh, sAct, HID, sB + HID, sAct, ng, pg);
    dense128(sWh + HID * HID, sAct, HID, sB + 2 * HID, sAct, ng, pg);
    float a[NC];
#pragma unroll
    for (int m = 0; m < NC; ++m) a[m] = sB[3 * HID + ((c0 + 4 * m) & 31)];
    for (int k = 0; k < HID; ++k) {
      const float h = sAct[act_idx(k, sp)];
#pragma unroll
      for (int m = 0; m < NC; ++m)
        if (c0 + 4 * m < DP) a[m] = fmaf(h, sW4[k * DP + c0 + 4 * m], a[m]);
    }
    if (live) {
#pragma unroll
      for (int m = 0; m < NC; ++m)
        if (c0 + 4 * m < d) out[gp * d + c0 + 4 * m] = a[m];
    }
  }
}

// ---- host-side dispatch --------------------------------------------------------------------------------------

template <int DP, int KIND>
static int launch_sample(msgm_ctx* ctx, const SampleParams& P, cudaStream_t stream) {
  using L = SmemLayout<DP>;
  auto kern = sample_fp32_kernel<DP, KIND>;
  MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  const long long ntiles = (P.B + TP - 1) / TP;
  const int grid = (int)std::min<long long>(ntiles, ctx->num_sms);
  kern<<<grid, NTHREADS, L::bytes, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

template <int DP>
static int launch_sample_kind(msgm_ctx* ctx, int kind, const SampleParams& P, cudaStream_t stream) {
  switch (kind) {
    case MSGM_SDE_SGM: return launch_sample<DP, MSGM_SDE_SGM>(ctx, P, stream);
    case MSGM_SDE_MSGM_DENSE: return launch_sample<DP, MSGM_SDE_MSGM_DENSE>(ctx, P, stream);
    case MSGM_SDE_MSGM_SPARSE: return launch_sample<DP, MSGM_SDE_MSGM_SPARSE>(ctx, P, stream);
  }
  set_error("unknown sde kind");
  return MSGM_ERR_INVALID;
}

int sample_mlp_fp32(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const msgm_sample_args* a,
                    float* x, int64_t B, cudaStream_t stream, const float* t_noise, const float* noise_single) {
  SampleParams P{};
  const int d = sde->dim;
  P.d = d;
  P.bmin = sde->beta_min;
  P.bdel = sde->beta_delta;
  P.Tsde = sde->T;
  P.G = sde->G;
  P.LG = sde->L_G;
  P.fwd = a->forward_only;
  if (!P.fwd) {
    P.pre = mlp->premodule;
    for (int l = 0; l < 4; ++l) { P.W[l] = mlp->W[l]; P.b[l] = mlp->b[l]; }
  }
  P.scheme = a->scheme;
  P.N = a->num_steps;
  P.nc = a->norm_correction;
  P.inc_t0 = a->include_t0 ? 1 : 0;
  P.lmbd = a->lmbd;
  // delta = T_/N is a Python double in the reference (sde_scheme.py:58,116,200); it meets fp32 tensors as
  // a scalar and is rounded to fp32 there.
  const double Trun = a->T_ >= 0.0f ? (double)a->T_ : (double)sde->T;
  const double delta = Trun / (double)a->num_steps;
  P.delta = (float)delta;
  P.delta_half = (float)(delta / 2.0);
  P.sqrt_delta = (float)std::sqrt(delta);
  P.ts = a->ts;
  P.noise = a->noise;
  P.seed = a->seed;
  P.poff = a->particle_offset;
  P.traj = a->traj;
  P.keep_step = a->keep_step;
  P.keep_out = a->keep_out;
  P.T_rows = a->T_rows;
  P.t_noise = t_noise;
  P.noise_single = noise_single;
  P.x = x;
  P.B = B;
  const int DP = d <= 2 ? 2 : d <= 4 ? 4 : d <= 8 ? 8 : d <= 16 ? 16 : 32;
  if (sde->kind == MSGM_SDE_MSGM_DENSE && DP == 32) {
    const size_t off = 1 << 18;  // second half of the context workspace (first half: tensor-core weight image)
    if (ctx->ws_bytes < off + sizeof(float) * 32 * 32 * 32) {
      set_error("internal: context workspace too small");
      return MSGM_ERR_INVALID;
    }
    float* gpad = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(ctx->ws) + off);
    pad_G32_kernel<<<128, 256, 0, stream>>>(sde->G, d, gpad);
    ctx->launches += 1;
    MSGM_CUDA_TRY(cudaGetLastError());
    P.G = gpad;
  }
  switch (DP) {
    case 2: return launch_sample_kind<2>(ctx, sde->kind, P, stream);
    case 4: return launch_sample_kind<4>(ctx, sde->kind, P, stream);
    case 8: return launch_sample_kind<8>(ctx, sde->kind, P, stream);
    case 16: return launch_sample_kind<16>(ctx, sde->kind, P, stream);
    default: return launch_sample_kind<32>(ctx, sde->kind, P, stream);
  }
}

template <int DP>
static int launch_fwd(msgm_ctx* ctx, const msgm_mlp_desc* m, const float* y, const float* s, float* out, int64_t B,
                      cudaStream_t stream) {
  using L = SmemLayout<DP>;
  auto kern = mlp_forward_fp32_kernel<DP>;
  MSGM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  const long long ntiles = (B + TP - 1) / TP;
  const int grid = (int)std::min<long long>(ntiles, ctx->num_sms);
  kern<<<grid, NTHREADS, L::bytes, stream>>>(m->input_dim, m->premodule, m->W[0], m->W[1], m->W[2], m->W[3],
                                              m->b[0], m->b[1], m->b[2], m->b[3], y, s, out, B);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int mlp_forward_fp32(msgm_ctx* ctx, const msgm_mlp_desc* m, const float* y, const float* s, float* out, int64_t B,
                     cudaStream_t stream) {
  const int d = m->input_dim;
  if (d <= 2) return launch_fwd<2>(ctx, m, y, s, out, B, stream);
  if (d <= 4) return launch_fwd<4>(ctx, m, y, s, out, B, stream);
  if (d <= 8) return launch_fwd<8>(ctx, m, y, s, out, B, stream);
  if (d <= 16) return launch_fwd<16>(ctx, m, y, s, out, B, stream);
  return launch_fwd<32>(ctx, m, y, s, out, B, stream);
}

}  // namespace msgm
