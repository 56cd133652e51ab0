// Sliced-score-matching training step for the MLP score net, fp32 CUDA cores.
//
// Replaces PluginReverseSDE.ssm_loss (SDEs.py:616-646): the reference evaluates
//     loss_b = v^T d/dy[ mu_to_div(y) ] v + |a|^2 / 2
// with a VJP through the net (autograd.grad(..., create_graph=True)) and then back-propagates through that VJP
// (double backward).  Here the same quantity is computed in FORWARD mode and differentiated by hand:
//     MSGM:  mu_to_div = g(s,y) a(y,s)   =>  loss = q . adot + |a|^2/2,  q_k = sqrt(beta) sum_ij v_i G_ijk y_j
//            (the term v^T g(s,v) a vanishes identically because every slice G[:,:,k] is skew-symmetric)
//     SGM :  mu_to_div = sqrt(beta) a + beta y / 2  =>  loss = sqrt(beta) v . adot + beta |v|^2 / 2 + |a|^2/2
// where adot = (da/dy) v is the tangent of the net output.  The primal and the tangent run through the MLP as two
// rows of the same tile (they share the weights):  z = W u + b, zdot = W udot, h = phi(z), hdot = phi'(z) zdot with
// phi = Swish.  The backward pass carries cotangents (zbar, zdotbar) and needs phi' and phi''.
//
// Three kernels:
//   ssm_forward_kernel   tile of 32 samples = 64 rows (primal/tangent interleaved); writes loss, and the per-layer
//                        pre-activations Z_l and layer inputs U_l to a global scratch (needed by the backward);
//   ssm_backward_kernel  activation backward: Cot_l = (zbar_l, zdotbar_l) for l = 3,2,1 and the output cotangent;
//   ssm_wgrad_kernel     gW_l = Cot_l^T U_{l-1} (reduction over all 2B rows) and gb_l, into one flat gradient buffer
//                        laid out like torch's parameter order [W0,b0,W1,b1,W2,b2,W3,b3].
#include <algorithm>

#include "msgm_common.cuh"

namespace msgm {

constexpr int TS = 32;           // samples per CTA tile
constexpr int TR = 2 * TS;       // rows per tile: row 2i = primal of sample i, row 2i+1 = its tangent
constexpr int SSM_THREADS = 256;

__device__ __forceinline__ int tile_idx(int feat, int r) {  // XOR-swizzled [feature][64 rows] tile (see sampler_fp32)
  return feat * TR + ((((r >> 2) ^ (feat >> 2)) & 15) << 2) + (r & 3);
}

// acc[i][j] = sum_k in[k][4pg+i] * Wt[k*ldw + f_j],  f_j = 4ng + (j&3) + 64 (j>>2);  Wt in shared or global memory
template <bool W_GLOBAL>
__device__ __forceinline__ void gemm_tile(const float* __restrict__ Wt, int ldw, const float* in, int K, int ng, int pg,
                                          float (&acc)[4][8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(in + k * TR + (((pg ^ (k >> 2)) & 15) << 2));
    float4 w0, w1;
    if (W_GLOBAL) {
      w0 = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)k * ldw + ng * 4));
      w1 = __ldg(reinterpret_cast<const float4*>(Wt + (size_t)k * ldw + 64 + ng * 4));
    } else {
      w0 = *reinterpret_cast<const float4*>(Wt + k * ldw + ng * 4);
      w1 = *reinterpret_cast<const float4*>(Wt + k * ldw + 64 + ng * 4);
    }
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}

struct SwishD {
  float h, d1, d2;  // phi(z), phi'(z), phi''(z)
};
__device__ __forceinline__ SwishD swish_derivs(float z) {
  const float sg = 1.0f / (1.0f + expf(-z));
  SwishD r;
  r.h = z * sg;
  r.d1 = sg * (1.0f + z * (1.0f - sg));
  r.d2 = sg * (1.0f - sg) * (2.0f + z * (1.0f - 2.0f * sg));
  return r;
}

struct SsmParams {
  int d, pre, kind;
  float bmin, bdel;
  const float* G;      // dense (d,d,d)
  const float* W[4];
  const float* b[4];
  const float* y;      // (B,d) noised samples
  const float* v;      // (B,d) Hutchinson probes
  const float* t;      // (B,) noise times s
  const float* gout;   // (B,) upstream gradient of the per-sample loss (backward only)
  float* loss;         // (B,)
  // scratch, row-major with 2B rows (row 2b primal, 2b+1 tangent)
  float* U0;           // [2B][K1P]   layer-1 input and its tangent (K1P = 36)
  float* Z[3];         // [2B][128]   pre-activations z_l, zdot_l
  float* U[3];         // [2B][128]   h_l, hdot_l
  float* A;            // [2B][DP32]  a, adot           (DP32 = 32)
  float* Q;            // [B][DP32]   q (MSGM) or sqrt(beta) v (SGM): cotangent direction of adot
  float* C[3];         // [2B][128]   cotangents zbar_l, zdotbar_l
  float* C4;           // [2B][DP32]  output cotangents abar, adotbar
  long long B;
};
constexpr int K1P = 36;   // padded row length of U0 (d + 2 <= 34)
constexpr int DP32 = 32;

// ------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSM_THREADS, 1) ssm_forward_kernel(const __grid_constant__ SsmParams P) {
  extern __shared__ __align__(16) float smem[];
  float* sWh = smem;                       // [2][128][128]  W2^T, W3^T  ([k][n])
  float* sAct = sWh + 2 * HID * HID;       // [128][64] swizzled
  float* sW1 = sAct + HID * TR;            // [34][128]
  float* sW4 = sW1 + 34 * HID;             // [128][32]
  float* sB = sW4 + HID * DP32;            // b1,b2,b3 [3][128], b4[32]
  float* sIn = sB + 3 * HID + 32;          // [36][64] swizzled
  float* sOut = sIn + 36 * TR;             // [32][64]  a / adot per row
  const int tid = threadIdx.x, ng = tid & 15, pg = tid >> 4;
  const int d = P.d, K1 = d + 1 + P.pre;

  // W (torch layout [n][k]) -> [k][n].  A warp reads 8 rows x 64 contiguous bytes per request (every sector fully used;
  // the element-wise transposing gather it replaces fetched each 32-byte sector eight times) and takes a 4-way bank
  // conflict on the four scalar stores instead -- this prologue is what a 256-row batch mostly consists of.
  for (int l = 0; l < 2; ++l)
    for (int q = tid; q < HID * HID / 4; q += SSM_THREADS) {
      const int kq = q & 3, ns = (q >> 2) & 7, r = q >> 5;
      const int k = (r & 7) * 16 + 4 * kq, n = (r >> 3) * 8 + ns;
      const float4 w = __ldg(reinterpret_cast<const float4*>(P.W[1 + l] + n * HID + k));
      float* dst = sWh + l * HID * HID + k * HID + n;
      dst[0] = w.x; dst[HID] = w.y; dst[2 * HID] = w.z; dst[3 * HID] = w.w;
    }
  for (int e = tid; e < K1 * HID; e += SSM_THREADS) sW1[e] = __ldg(P.W[0] + (e & 127) * K1 + (e >> 7));
  for (int e = tid; e < HID * DP32; e += SSM_THREADS) sW4[e] = (e & 31) < d ? __ldg(P.W[3] + (e & 31) * HID + (e >> 5)) : 0.0f;
  for (int e = tid; e < 3 * HID; e += SSM_THREADS) sB[e] = __ldg(P.b[e >> 7] + (e & 127));
  if (tid < 32) sB[3 * HID + tid] = tid < d ? __ldg(P.b[3] + tid) : 0.0f;
  __syncthreads();

  const long long ntiles = (P.B + TS - 1) / TS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * TS;
    __syncthreads();
    // ---- layer-1 input and its tangent (premodule NN.py:64-70 and its Jacobian-vector product) -------------------
    if (tid < TS) {
      const long long b = b0 + tid;
      const bool live = b < P.B;
      float r2 = 0.0f, yv = 0.0f;
      for (int c = 0; c < d; ++c) {
        const float yc = live ? P.y[b * d + c] : (c == 0 ? 1.0f : 0.0f);
        const float vc = live ? P.v[b * d + c] : 0.0f;
        r2 = fmaf(yc, yc, r2);
        yv = fmaf(yc, vc, yv);
      }
      const float s = live ? P.t[b] : 0.0f;
      const int rp = 2 * tid, rt = 2 * tid + 1;
      if (P.pre) {
        const float r = sqrtf(r2), rn = r + 1e-6f;
        const float rdot = yv / r;  // d|y| along v
        for (int c = 0; c < d; ++c) {
          const float yc = live ? P.y[b * d + c] : (c == 0 ? 1.0f : 0.0f);
          const float vc = live ? P.v[b * d + c] : 0.0f;
          sIn[tile_idx(c, rp)] = yc / rn;
          sIn[tile_idx(c, rt)] = vc / rn - yc * rdot / (rn * rn);
        }
        sIn[tile_idx(d, rp)] = logf(rn);
        sIn[tile_idx(d, rt)] = rdot / rn;
      } else {
        for (int c = 0; c < d; ++c) {
          sIn[tile_idx(c, rp)] = live ? P.y[b * d + c] : 0.0f;
          sIn[tile_idx(c, rt)] = live ? P.v[b * d + c] : 0.0f;
        }
      }
      sIn[tile_idx(d + P.pre, rp)] = s;
      sIn[tile_idx(d + P.pre, rt)] = 0.0f;
    }
    __syncthreads();
    for (int e = tid; e < TR * K1; e += SSM_THREADS) {  // U0 scratch
      const int r = e / K1, k = e % K1;
      const long long gr = 2 * b0 + r;
      if (gr < 2 * P.B) P.U0[gr * K1P + k] = sIn[tile_idx(k, r)];
    }

    // ---- three hidden layers: z = W u + b, zdot = W udot ; h = phi(z), hdot = phi'(z) zdot ---------------------------
    for (int l = 0; l < 3; ++l) {
      float acc[4][8];
      if (l == 0) gemm_tile<false>(sW1, HID, sIn, K1, ng, pg, acc);
      else gemm_tile<false>(sWh + (l - 1) * HID * HID, HID, sAct, HID, ng, pg, acc);
      __syncthreads();  // all reads of the input tile done (sAct is overwritten below)
      const float* bias = sB + l * HID;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {  // the thread's two quads of 4 contiguous features: 16-byte global stores
        const int f0 = ng * 4 + 64 * jh;
        float o[4][4];                  // [feature in quad][row in the thread's 4-row group]
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          float4 z4, zd4, h4, hd4;
          float* zp = &z4.x; float* zdp = &zd4.x; float* hp = &h4.x; float* hdp = &hd4.x;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float z = acc[2 * si][4 * jh + q] + bias[f0 + q], zd = acc[2 * si + 1][4 * jh + q];
            const SwishD sw = swish_derivs(z);
            zp[q] = z; zdp[q] = zd; hp[q] = sw.h; hdp[q] = sw.d1 * zd;
            o[q][2 * si] = sw.h;
            o[q][2 * si + 1] = sw.d1 * zd;
          }
          const long long gr = 2 * b0 + 4 * pg + 2 * si;
          if (gr < 2 * P.B) {
            *reinterpret_cast<float4*>(P.Z[l] + gr * HID + f0) = z4;
            *reinterpret_cast<float4*>(P.Z[l] + (gr + 1) * HID + f0) = zd4;
            *reinterpret_cast<float4*>(P.U[l] + gr * HID + f0) = h4;
            *reinterpret_cast<float4*>(P.U[l] + (gr + 1) * HID + f0) = hd4;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(sAct + (f0 + q) * TR + (((pg ^ ((f0 + q) >> 2)) & 15) << 2)) =
              make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
      }
      __syncthreads();
    }

    // ---- output layer: a = W4 h3 + b4 (primal rows), adot = W4 hdot3 (tangent rows) -----------------------------------------
    {
      const int r = tid & 63, c0 = tid >> 6;  // thread: row r, components 8 c0 .. 8 c0 + 7 (a warp shares c0)
      float o[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) o[m] = (r & 1) ? 0.0f : sB[3 * HID + 8 * c0 + m];
      if (8 * c0 < d) {
#pragma unroll 4
        for (int k = 0; k < HID; ++k) {
          const float h = sAct[tile_idx(k, r)];
          const float4 w0 = *reinterpret_cast<const float4*>(sW4 + k * DP32 + 8 * c0);
          const float4 w1 = *reinterpret_cast<const float4*>(sW4 + k * DP32 + 8 * c0 + 4);
          o[0] = fmaf(h, w0.x, o[0]); o[1] = fmaf(h, w0.y, o[1]); o[2] = fmaf(h, w0.z, o[2]); o[3] = fmaf(h, w0.w, o[3]);
          o[4] = fmaf(h, w1.x, o[4]); o[5] = fmaf(h, w1.y, o[5]); o[6] = fmaf(h, w1.z, o[6]); o[7] = fmaf(h, w1.w, o[7]);
        }
      }
      const long long gr = 2 * b0 + r;
#pragma unroll
      for (int m = 0; m < 8; ++m) sOut[(8 * c0 + m) * TR + r] = o[m];
      if (gr < 2 * P.B) {
        *reinterpret_cast<float4*>(P.A + gr * DP32 + 8 * c0) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(P.A + gr * DP32 + 8 * c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    __syncthreads();

    // ---- cotangent direction q of adot: one (sample, k) pair per thread pass (sIn is free: reused as q[k][sample]) ----
    float* sQ = sIn;
    for (int e = tid; e < TS * d; e += SSM_THREADS) {
      const int si = e / d, k = e - si * d;
      const long long b = b0 + si;
      float q = 0.0f;
      if (b < P.B) {
        const float sb = sqrtf(beta_of(P.bmin, P.bdel, P.t[b]));
        const float* yb = P.y + b * d;
        const float* vb = P.v + b * d;
        if (P.kind == MSGM_SDE_SGM) {
          q = sb * vb[k];
        } else if (P.kind == MSGM_SDE_MSGM_SPARSE) {
          // q_k = c sqrt(beta) (v_k y_{k+1} - v_{k+1} y_k): entries (i=k,j=k+1,k):+c and (i=k+1,j=k,k):-c (SDEs.py:375-380)
          const int kn = (k + 1 == d) ? 0 : k + 1;
          q = SQRT_HALF * sb * (vb[k] * yb[kn] - vb[kn] * yb[k]);
        } else {
          float acc = 0.0f;
          for (int i = 0; i < d; ++i) {
            float u = 0.0f;
            for (int j = 0; j < d; ++j) u = fmaf(__ldg(P.G + (i * d + j) * d + k), yb[j], u);
            acc = fmaf(vb[i], u, acc);
          }
          q = sb * acc;
        }
        P.Q[b * DP32 + k] = q;
      }
      sQ[k * TS + si] = q;
    }
    __syncthreads();
    // ---- per-sample loss ------------------------------------------------------------------------------------------------
    if (tid < TS) {
      const long long b = b0 + tid;
      if (b < P.B) {
        const float bt = beta_of(P.bmin, P.bdel, P.t[b]);
        float loss = 0.0f;
        for (int k = 0; k < d; ++k) {
          const float a = sOut[k * TR + 2 * tid], ad = sOut[k * TR + 2 * tid + 1];
          if (P.kind == MSGM_SDE_SGM) {
            const float vk = P.v[b * d + k];
            loss = fmaf(0.5f * bt, vk * vk, loss);
          }
          loss = fmaf(sQ[k * TS + tid], ad, loss);
          loss = fmaf(0.5f * a, a, loss);
        }
        P.loss[b] = loss;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// activation backward: cotangents (zbar, zdotbar) of every hidden layer
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SSM_THREADS, 1) ssm_backward_kernel(const __grid_constant__ SsmParams P) {
  extern __shared__ __align__(16) float smem[];
  float* sW = smem;                      // [2][128][128]  W3, W2 in torch layout ([n][k]): H̄_{l-1} = Cot_l W_l
  float* sCot = sW + 2 * HID * HID;      // [128][64] swizzled
  float* sW4 = sCot + HID * TR;          // [32][128]  W4 (d,128)
  float* sC4 = sW4 + DP32 * HID;         // [32][64]   abar / adotbar per row
  const int tid = threadIdx.x, ng = tid & 15, pg = tid >> 4;
  const int d = P.d;
  for (int e = tid; e < HID * HID; e += SSM_THREADS) {
    sW[e] = __ldg(P.W[2] + e);
    sW[HID * HID + e] = __ldg(P.W[1] + e);
  }
  for (int e = tid; e < DP32 * HID; e += SSM_THREADS) sW4[e] = (e >> 7) < d ? __ldg(P.W[3] + e) : 0.0f;
  __syncthreads();

  const long long ntiles = (P.B + TS - 1) / TS;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * TS;
    __syncthreads();
    // ---- output cotangents: abar = w a (from |a|^2/2), adotbar = w q (from q . adot) ----------------------------------
    for (int e = tid; e < DP32 * TR; e += SSM_THREADS) {
      const int c = e / TR, r = e % TR;
      const long long b = b0 + (r >> 1);
      float val = 0.0f;
      if (b < P.B && c < d) {
        const float w = P.gout[b];
        val = (r & 1) ? w * P.Q[b * DP32 + c] : w * P.A[(2 * b) * DP32 + c];
        P.C4[(2 * b + (r & 1)) * DP32 + c] = val;
      }
      sC4[e] = val;
    }
    __syncthreads();

    for (int l = 2; l >= 0; --l) {
      // this layer's pre-activations: issued before the product below so that their latency hides behind it
      float4 zq[2][2], zdq[2][2];
#pragma unroll
      for (int jh = 0; jh < 2; ++jh)
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          const long long gr = 2 * b0 + 4 * pg + 2 * si;
          const int f0 = ng * 4 + 64 * jh;
          if (gr < 2 * P.B) {
            zq[jh][si] = __ldg(reinterpret_cast<const float4*>(P.Z[l] + gr * HID + f0));
            zdq[jh][si] = __ldg(reinterpret_cast<const float4*>(P.Z[l] + (gr + 1) * HID + f0));
          } else {
            zq[jh][si] = zdq[jh][si] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      // hbar tile (rows: hbar, hdotbar) = Cot_{l+1} W_{l+1}
      float acc[4][8];
      if (l == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
        for (int c = 0; c < d; ++c) {
          const float4 c4 = *reinterpret_cast<const float4*>(sC4 + c * TR + 4 * pg);
          const float4 w0 = *reinterpret_cast<const float4*>(sW4 + c * HID + ng * 4);
          const float4 w1 = *reinterpret_cast<const float4*>(sW4 + c * HID + 64 + ng * 4);
          const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(cv[i], wv[j], acc[i][j]);
        }
      } else {
        gemm_tile<false>(sW + (l == 1 ? 0 : HID * HID), HID, sCot, HID, ng, pg, acc);
        __syncthreads();  // sCot is overwritten below
      }
      // zbar = hbar phi'(z) + hdotbar phi''(z) zdot ;  zdotbar = hdotbar phi'(z)
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {  // two quads of 4 contiguous features: 16-byte global loads and stores
        const int f0 = ng * 4 + 64 * jh;
        float o[4][4] = {};
#pragma unroll
        for (int si = 0; si < 2; ++si) {
          const long long gr = 2 * b0 + 4 * pg + 2 * si;
          if (gr < 2 * P.B) {
            const float4 z4 = zq[jh][si], zd4 = zdq[jh][si];
            const float* zp = &z4.x; const float* zdp = &zd4.x;
            float4 c4, cd4;
            float* cp = &c4.x; float* cdp = &cd4.x;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const SwishD sw = swish_derivs(zp[q]);
              const float hb = acc[2 * si][4 * jh + q], hdb = acc[2 * si + 1][4 * jh + q];
              cp[q] = fmaf(hb, sw.d1, hdb * sw.d2 * zdp[q]);
              cdp[q] = hdb * sw.d1;
              o[q][2 * si] = cp[q];
              o[q][2 * si + 1] = cdp[q];
            }
            *reinterpret_cast<float4*>(P.C[l] + gr * HID + f0) = c4;
            *reinterpret_cast<float4*>(P.C[l] + (gr + 1) * HID + f0) = cd4;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(sCot + (f0 + q) * TR + (((pg ^ ((f0 + q) >> 2)) & 15) << 2)) =
              make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradients: out[m][n] (+)= sum_r Cot[r][m] * U[r][n] over a slice of rows; bias: sum over PRIMAL rows of Cot
// ------------------------------------------------------------------------------------------------------------------
struct WgradJob {
  const float* cot;  // [R][ldc]
  const float* u;    // [R][ldu]
  int ldc, ldu, M, N;
  float* gW;         // [M][N]
  float* gb;         // [M]
};
struct WgradParams {
  WgradJob job[4];
  long long R;       // rows (= 2B)
  int rows_per_slice;
  int atomic;        // more than one row slice -> accumulate with atomics into a zeroed buffer
};

__global__ void __launch_bounds__(256) ssm_wgrad_kernel(const __grid_constant__ WgradParams P) {
  // grid: x = 64x64 output tile (up to 4 per job), y = job, z = row slice
  const WgradJob J = P.job[blockIdx.y];
  const int tm = (blockIdx.x >> 1) * 64, tn = (blockIdx.x & 1) * 64;
  if (tm >= J.M || tn >= J.N) return;
  __shared__ float sC[32][65];
  __shared__ float sU[32][65];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // thread computes rows m = ty*4.., cols n = tx*4..
  float acc[4][4] = {};
  float bacc = 0.0f;  // bias partial: thread tid < 64 sums column tm+tid over primal rows
  const long long r_begin = (long long)blockIdx.z * P.rows_per_slice;
  const long long r_end = min(P.R, r_begin + P.rows_per_slice);
  // 32-row chunks, software-pipelined through registers: the global loads of chunk c+1 are in flight while chunk c
  // is multiplied out of shared memory (at the reference's batch of 256 a slice is only a few chunks long and the
  // load latency would otherwise be the whole run time).
  float pc[8], pu[8];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 256 * q, rr = e >> 6, cc = e & 63;
      const long long r = r0 + rr;
      pc[q] = (r < r_end && tm + cc < J.M) ? __ldg(J.cot + r * J.ldc + tm + cc) : 0.0f;
      pu[q] = (r < r_end && tn + cc < J.N) ? __ldg(J.u + r * J.ldu + tn + cc) : 0.0f;
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (long long r0 = r_begin; r0 < r_end; r0 += 32) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 256 * q;
      sC[e >> 6][e & 63] = pc[q];
      sU[e >> 6][e & 63] = pu[q];
    }
    __syncthreads();
    if (r0 + 32 < r_end) fetch(r0 + 32);
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      float c4[4], u4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { c4[i] = sC[rr][ty * 4 + i]; u4[i] = sU[rr][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(c4[i], u4[j], acc[i][j]);
    }
    if (tn == 0 && tid < 64) {
#pragma unroll
      for (int rr = 0; rr < 32; rr += 2) bacc += sC[rr][tid];  // r0 is even, so even rr = primal rows
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = tm + ty * 4 + i, n = tn + tx * 4 + j;
      if (m < J.M && n < J.N) {
        if (P.atomic) atomicAdd(J.gW + (size_t)m * J.N + n, acc[i][j]);
        else J.gW[(size_t)m * J.N + n] = acc[i][j];
      }
    }
  if (tn == 0 && tid < 64 && tm + tid < J.M) {
    if (P.atomic) atomicAdd(J.gb + tm + tid, bacc);
    else J.gb[tm + tid] = bacc;
  }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
size_t ssm_scratch_floats(long long B) {
  const size_t R = 2 * (size_t)B;
  return R * K1P + 9 * R * HID + 2 * R * DP32 + (size_t)B * DP32;
}

static void carve(SsmParams& P, float* s, long long B) {
  const size_t R = 2 * (size_t)B;
  P.U0 = s; s += R * K1P;
  for (int l = 0; l < 3; ++l) { P.Z[l] = s; s += R * HID; }
  for (int l = 0; l < 3; ++l) { P.U[l] = s; s += R * HID; }
  for (int l = 0; l < 3; ++l) { P.C[l] = s; s += R * HID; }
  P.A = s; s += R * DP32;
  P.C4 = s; s += R * DP32;
  P.Q = s;
}

static int fill(SsmParams& P, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y, const float* v,
                const float* t, float* scratch, long long B) {
  P.d = sde->dim;
  P.pre = mlp->premodule;
  P.kind = sde->kind;
  P.bmin = sde->beta_min;
  P.bdel = sde->beta_delta;
  P.G = sde->G;
  for (int l = 0; l < 4; ++l) { P.W[l] = mlp->W[l]; P.b[l] = mlp->b[l]; }
  P.y = y; P.v = v; P.t = t;
  P.B = B;
  carve(P, scratch, B);
  return MSGM_OK;
}

int ssm_forward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y, const float* v,
                const float* t, float* loss, float* scratch, int64_t B, cudaStream_t stream) {
  SsmParams P{};
  fill(P, sde, mlp, y, v, t, scratch, B);
  P.loss = loss;
  const size_t smem = sizeof(float) * (2 * HID * HID + HID * TR + 34 * HID + HID * DP32 + 3 * HID + 32 + 36 * TR + 32 * TR);
  MSGM_CUDA_TRY(cudaFuncSetAttribute(ssm_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (B + TS - 1) / TS;
  ssm_forward_kernel<<<(int)std::min<long long>(ntiles, ctx->num_sms), SSM_THREADS, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

int ssm_backward(msgm_ctx* ctx, const msgm_sde_desc* sde, const msgm_mlp_desc* mlp, const float* y, const float* v,
                 const float* t, const float* gout, float* scratch, float* grad_flat, int64_t B, cudaStream_t stream) {
  SsmParams P{};
  fill(P, sde, mlp, y, v, t, scratch, B);
  P.gout = gout;
  const size_t smem = sizeof(float) * (2 * HID * HID + HID * TR + DP32 * HID + DP32 * TR);
  MSGM_CUDA_TRY(cudaFuncSetAttribute(ssm_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = (B + TS - 1) / TS;
  ssm_backward_kernel<<<(int)std::min<long long>(ntiles, ctx->num_sms), SSM_THREADS, smem, stream>>>(P);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());

  const int d = sde->dim, K1 = d + 1 + mlp->premodule;
  WgradParams W{};
  float* g = grad_flat;
  // torch parameter order: main.0.weight (128,K1), main.0.bias, main.2.*, main.4.*, main.6.weight (d,128), main.6.bias
  W.job[0] = {P.C[0], P.U0, HID, K1P, HID, K1, g, g + HID * K1};                g += HID * K1 + HID;
  W.job[1] = {P.C[1], P.U[0], HID, HID, HID, HID, g, g + HID * HID};            g += HID * HID + HID;
  W.job[2] = {P.C[2], P.U[1], HID, HID, HID, HID, g, g + HID * HID};            g += HID * HID + HID;
  W.job[3] = {P.C4, P.U[2], DP32, HID, d, HID, g, g + d * HID};
  W.R = 2 * B;
  const int max_slices = 64;
  long long rps = (W.R + max_slices - 1) / max_slices;
  rps = std::max<long long>(64, (rps + 31) / 32 * 32);  // >= 2 chunks per slice; small batches still fill the GPU
  W.rows_per_slice = (int)rps;
  const int nslice = (int)((W.R + rps - 1) / rps);
  W.atomic = nslice > 1;
  if (W.atomic) {
    const size_t nparam = (size_t)HID * K1 + HID + 2 * ((size_t)HID * HID + HID) + (size_t)d * HID + d;
    MSGM_CUDA_TRY(cudaMemsetAsync(grad_flat, 0, sizeof(float) * nparam, stream));
  }
  ssm_wgrad_kernel<<<dim3(4, 4, nslice), 256, 0, stream>>>(W);
  ctx->launches += 1;
  MSGM_CUDA_TRY(cudaGetLastError());
  return MSGM_OK;
}

}  // namespace msgm
