// Readout tail of the SHMP neighborhood-counting model, one launch each (sm_100a, fp32 FFMA):
//
//   readout_chain_kernel      post_mp  576 -> 64 -> LeakyReLU(.1) -> 64 -> ReLU -> 256 -> ReLU -> 64
//                             (subgraph_counting/gnn_model.py:44-53,108; Dropout between is identity in eval / p = 0)
//   count_head_fused_kernel   query-conditioned count head, factorised (lightning_model.py:127-131,176-193,212-221):
//                             pred[g,q] = w2 . leaky_.01( t_g . W1a + (q_q . W1b + b1) ) + b2 ;  count = 2^pred - 1
//
// The reference runs these as 4 + 29 x 2 cuBLAS launches on [G, .] activations that round-trip through HBM; here a CTA
// owns 16 neighborhoods and keeps every intermediate in shared memory (two CTAs per SM: the weights stream from L2 with
// no explicit prefetch, so a second CTA is what hides their latency).  These GEMMs are tiny per row (90 k MAC) and the
// sums cancel heavily, so they run as exact fp32 FFMA on all SMs (G = 4096 -> 256 CTAs) instead of a 6-pass bf16 split
// on a 128-row tensor-core tile (32 CTAs).  The query half of the head's first Linear (q_q . W1b + b1, the same for every
// neighborhood) is computed once per call by count_head_query_kernel.  Two warp-level GEMM shapes:
//   * N = 64  : split-K - each of the 8 warps multiplies its K-slice into a full 16 x 64 partial (4 x 8 outputs per
//               lane, weights straight from L1/L2 as float4), partials are summed through shared memory;
//   * N = 256 : split-N - each warp owns 32 output columns over the whole K = 64 (4 x 4 outputs per lane).
#include "common.cuh"
#include "shmp_internal.h"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;
constexpr int ROWS = 16;       // neighborhoods per CTA (16: ~95 KB of shared memory -> two CTAs per SM hide the weight loads of each other)
constexpr int RI = ROWS / 4;   // rows per lane: ry + 4 i
constexpr int THREADS = 256;
constexpr int NW = THREADS / 32;
constexpr int H4 = 4 * F;      // 256
constexpr int LD64 = F + 4;    // 68  = 4 mod 32: rows r, r+1, r+2, r+3 hit disjoint bank quads
constexpr int LD256 = H4 + 4;  // 260 = 4 mod 32
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

__device__ __forceinline__ float act_fn(float v, int act, float slope) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LEAKY) return v > 0.f ? v : v * slope;
  return v;
}
__device__ __forceinline__ float comp(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// sPart[warp][r][0..64) = sum_{k in this warp's slice} sX[r][k] . W[k][0..64)     (K % 32 == 0, W row-major [K][64])
// lane = (ry, cx): rows ry + 4 i (i < 8), columns 8 cx .. 8 cx + 7
template <int LDX>
__device__ __forceinline__ void splitk_gemm64(const float* sX, int K, const float* __restrict__ W, float* sPart, int warp,
                                              int lane) {
  const int ry = lane >> 3, cx = lane & 7;
  float acc[RI][8];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int kper = K / NW;
  const int k0 = warp * kper;
#pragma unroll 2
  for (int k4 = k0; k4 < k0 + kper; k4 += 4) {
    float4 a[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4*>(sX + (ry + 4 * i) * LDX + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k4 + kk) * F + 8 * cx));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k4 + kk) * F + 8 * cx + 4));
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float av = comp(a[i], kk);
        acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
        acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
        acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
        acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    float* dst = sPart + ((size_t)warp * ROWS + ry + 4 * i) * F + 8 * cx;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
}

// sY[r][n] = act(sum_w sPart[w][r][n] + bias[n])  (to shared memory, and/or to global rows g0 + r < G)
template <int LDY>
__device__ __forceinline__ void reduce_partials(const float* sPart, const float* __restrict__ bias, int act, float slope,
                                                float* sY, float* __restrict__ gY, int g0, int G, int tid) {
#pragma unroll
  for (int j = 0; j < ROWS * F / THREADS; ++j) {
    const int idx = tid + THREADS * j;
    const int r = idx >> 6, n = idx & 63;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += sPart[((size_t)w * ROWS + r) * F + n];
    s = act_fn(s + __ldg(bias + n), act, slope);
    if (sY) sY[r * LDY + n] = s;
    if (gY && g0 + r < G) gY[(size_t)(g0 + r) * F + n] = s;
  }
}

// sY[r][32 warp + 4 cx ..+3] = act(sX[r][0..64) . W[0..64)[..] + bias)   (W row-major [64][256]); rows ry + 4 i
template <int LDX, int LDY>
__device__ __forceinline__ void splitn_gemm256(const float* sX, const float* __restrict__ W, const float* __restrict__ bias,
                                               int act, float slope, float* sY, int warp, int lane) {
  const int ry = lane >> 3, cx = lane & 7;
  const int c0 = 32 * warp + 4 * cx;
  float acc[RI][4];
#pragma unroll
  for (int i = 0; i < RI; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
  for (int k4 = 0; k4 < F; k4 += 4) {
    float4 a[RI];
#pragma unroll
    for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4*>(sX + (ry + 4 * i) * LDX + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)(k4 + kk) * H4 + c0));
#pragma unroll
      for (int i = 0; i < RI; ++i) {
        const float av = comp(a[i], kk);
        acc[i][0] = fmaf(av, w.x, acc[i][0]); acc[i][1] = fmaf(av, w.y, acc[i][1]);
        acc[i][2] = fmaf(av, w.z, acc[i][2]); acc[i][3] = fmaf(av, w.w, acc[i][3]);
      }
    }
  }
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + c0));
#pragma unroll
  for (int i = 0; i < RI; ++i) {
    float4 o;
    o.x = act_fn(acc[i][0] + b.x, act, slope); o.y = act_fn(acc[i][1] + b.y, act, slope);
    o.z = act_fn(acc[i][2] + b.z, act, slope); o.w = act_fn(acc[i][3] + b.w, act, slope);
    *reinterpret_cast<float4*>(sY + (ry + 4 * i) * LDY + c0) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// post_mp chain
// ------------------------------------------------------------------------------------------------------------------
struct ChainArgs {
  const float* Z; int ldz; int K0; int G;        // [G][K0] pooled embedding (K0 = (layers + 1) * 64)
  const float *P0, *b0, *P1, *b1, *P2, *b2, *P3, *b3;  // row-major [in][out]
  float* out;                                    // [G][64]
  const int32_t* g_dev;                          // device-resident G (stream-ordered form) or NULL
};

__global__ void __launch_bounds__(THREADS, 2) readout_chain_kernel(const ChainArgs p) {
  extern __shared__ __align__(16) float sm[];
  const int ldx = p.K0 + 4;                      // K0 % 32 == 0 -> ldx = 4 mod 32
  float* sX = sm;                                // [ROWS][ldx]
  float* sPart = sX + ROWS * ldx;                // [NW][ROWS][64]
  float* sT1 = sPart + NW * ROWS * F;            // [ROWS][LD64]
  float* sT2 = sT1 + ROWS * LD64;                // [ROWS][LD64]
  float* sT3 = sT2 + ROWS * LD64;                // [ROWS][LD256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * ROWS;
  const int G = p.g_dev ? min(p.G, *p.g_dev) : p.G;
  if (g0 >= G) return;

  const int k4n = p.K0 / 4;
  for (int i = tid; i < ROWS * k4n; i += THREADS) {
    const int r = i / k4n, c4 = (i - r * k4n) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g0 + r < G) v = __ldg(reinterpret_cast<const float4*>(p.Z + (size_t)(g0 + r) * p.ldz + c4));
    *reinterpret_cast<float4*>(sX + r * ldx + c4) = v;
  }
  __syncthreads();
  // Linear(K0,64): split-K like splitk_gemm64, written out because the row pitch of sX is a run-time value
  {
    const int ry = lane >> 3, cx = lane & 7;
    float acc[RI][8];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    const int kper = p.K0 / NW;  // K0 % 32 == 0
    const int k0 = warp * kper;
#pragma unroll 2
    for (int k4 = k0; k4 < k0 + kper; k4 += 4) {
      float4 a[RI];
#pragma unroll
      for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4*>(sX + (ry + 4 * i) * ldx + k4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.P0 + (size_t)(k4 + kk) * F + 8 * cx));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.P0 + (size_t)(k4 + kk) * F + 8 * cx + 4));
#pragma unroll
        for (int i = 0; i < RI; ++i) {
          const float av = comp(a[i], kk);
          acc[i][0] = fmaf(av, w0.x, acc[i][0]); acc[i][1] = fmaf(av, w0.y, acc[i][1]);
          acc[i][2] = fmaf(av, w0.z, acc[i][2]); acc[i][3] = fmaf(av, w0.w, acc[i][3]);
          acc[i][4] = fmaf(av, w1.x, acc[i][4]); acc[i][5] = fmaf(av, w1.y, acc[i][5]);
          acc[i][6] = fmaf(av, w1.z, acc[i][6]); acc[i][7] = fmaf(av, w1.w, acc[i][7]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RI; ++i) {
      float* dst = sPart + ((size_t)warp * ROWS + ry + 4 * i) * F + 8 * cx;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
  }
  __syncthreads();
  reduce_partials<LD64>(sPart, p.b0, ACT_LEAKY, 0.1f, sT1, nullptr, g0, G, tid);   // Linear(576,64) + LeakyReLU(.1)
  __syncthreads();
  splitk_gemm64<LD64>(sT1, F, p.P1, sPart, warp, lane);
  __syncthreads();
  reduce_partials<LD64>(sPart, p.b1, ACT_RELU, 0.f, sT2, nullptr, g0, G, tid);      // Linear(64,64) + ReLU
  __syncthreads();
  splitn_gemm256<LD64, LD256>(sT2, p.P2, p.b2, ACT_RELU, 0.f, sT3, warp, lane);       // Linear(64,256) + ReLU
  __syncthreads();
  splitk_gemm64<LD256>(sT3, H4, p.P3, sPart, warp, lane);
  __syncthreads();
  reduce_partials<LD64>(sPart, p.b3, ACT_NONE, 0.f, nullptr, p.out, g0, G, tid);    // Linear(256,64)
}

// ------------------------------------------------------------------------------------------------------------------
// count head
// ------------------------------------------------------------------------------------------------------------------
struct HeadArgs {
  const float* emb_t; int G;    // [G][64]
  const float* Bq; int Q;       // [Q][256] = q_q . W1b + b1 (count_head_query_kernel), Q <= 32
  const float *W1a, *w2, *b2;   // [64][256], [256], [1]
  float* pred;                  // [G][Q] or NULL
  float* count;                 // [G][Q] or NULL
  const int32_t* g_dev;         // device-resident G (stream-ordered form) or NULL
};
constexpr int QROWS = 32;       // query rows held in shared memory

// Bq[q][0..256) = emb_q[q] . W1b + b1: the query half of the head's first Linear, once per call instead of once per CTA
__global__ void __launch_bounds__(H4) count_head_query_kernel(const float* __restrict__ emb_q, const float* __restrict__ W1b,
                                                              const float* __restrict__ b1, float* __restrict__ Bq) {
  __shared__ float sq[F];
  const int q = blockIdx.x, j = threadIdx.x;
  if (j < F) sq[j] = emb_q[(size_t)q * F + j];
  __syncthreads();
  float w[F];  // all 64 loads in flight before the first FMA (the kernel is one L2 round trip long)
#pragma unroll
  for (int k = 0; k < F; ++k) w[k] = __ldg(W1b + (size_t)k * H4 + j);
  float acc = __ldg(b1 + j);
#pragma unroll
  for (int k = 0; k < F; ++k) acc = fmaf(sq[k], w[k], acc);
  Bq[(size_t)q * H4 + j] = acc;
}

__global__ void __launch_bounds__(THREADS, 2) count_head_fused_kernel(const HeadArgs p) {
  extern __shared__ __align__(16) float sm[];
  float* sE = sm;                      // [ROWS][LD64]   target embeddings
  float* sT = sE + ROWS * LD64;        // [ROWS][LD256]  t_g . W1a
  float* sB = sT + ROWS * LD256;       // [QROWS][LD256] q_q . W1b + b1
  float* sW2 = sB + QROWS * LD256;     // [256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g0 = blockIdx.x * ROWS;
  const int G = p.g_dev ? min(p.G, *p.g_dev) : p.G;
  if (g0 >= G) return;
  for (int i = tid; i < ROWS * (F / 4); i += THREADS) {
    const int r = i >> 4, c4 = (i & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g0 + r < G) v = __ldg(reinterpret_cast<const float4*>(p.emb_t + (size_t)(g0 + r) * F + c4));
    *reinterpret_cast<float4*>(sE + r * LD64 + c4) = v;
  }
  for (int i = tid; i < p.Q * (H4 / 4); i += THREADS) {
    const int r = i >> 6, c4 = (i & 63) * 4;
    *reinterpret_cast<float4*>(sB + r * LD256 + c4) = __ldg(reinterpret_cast<const float4*>(p.Bq + (size_t)r * H4 + c4));
  }
  for (int i = tid; i < H4; i += THREADS) sW2[i] = __ldg(p.w2 + i);
  __syncthreads();
  splitn_gemm256<LD64, LD256>(sE, p.W1a, nullptr, ACT_NONE, 0.f, sT, warp, lane);
  __syncthreads();
  const float bias2 = __ldg(p.b2);
  const int pairs = ROWS * p.Q;
  for (int i = tid; i < pairs; i += THREADS) {
    const int r = i / p.Q, q = i - r * p.Q;   // consecutive lanes: consecutive queries (bank-disjoint sB rows)
    const float* t = sT + r * LD256;
    const float* bq = sB + q * LD256;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
    for (int j = 0; j < H4; j += 4) {
      const float4 tv = *reinterpret_cast<const float4*>(t + j);
      const float4 bv = *reinterpret_cast<const float4*>(bq + j);
      const float4 wv = *reinterpret_cast<const float4*>(sW2 + j);
      float v0 = tv.x + bv.x, v1 = tv.y + bv.y, v2 = tv.z + bv.z, v3 = tv.w + bv.w;
      v0 = v0 > 0.f ? v0 : 0.01f * v0;  // nn.LeakyReLU() default slope
      v1 = v1 > 0.f ? v1 : 0.01f * v1;
      v2 = v2 > 0.f ? v2 : 0.01f * v2;
      v3 = v3 > 0.f ? v3 : 0.01f * v3;
      a0 = fmaf(v0, wv.x, a0); a1 = fmaf(v1, wv.y, a1); a2 = fmaf(v2, wv.z, a2); a3 = fmaf(v3, wv.w, a3);
    }
    const float acc = ((a0 + a1) + (a2 + a3)) + bias2;
    if (g0 + r < G) {
      if (p.pred) p.pred[(size_t)(g0 + r) * p.Q + q] = acc;
      if (p.count) p.count[(size_t)(g0 + r) * p.Q + q] = exp2f(acc) - 1.f;  // 2**pred - 1 (lightning_model.py:221)
    }
  }
}

}  // namespace

int desco_internal_readout_chain(const float* Z, int ldz, int K0, int G, const float* P0, const float* b0, const float* P1,
                                 const float* b1, const float* P2, const float* b2, const float* P3, const float* b3,
                                 float* out, const int32_t* g_dev, cudaStream_t s) {
  if (G == 0) return DESCO_OK;
  if (K0 % 32 || K0 <= 0) return DESCO_EINVAL;
  const size_t smem = ((size_t)ROWS * (K0 + 4) + NW * ROWS * F + 2 * ROWS * LD64 + ROWS * LD256) * sizeof(float);
  if (smem > 227 * 1024) return DESCO_ERANGE;
  static size_t attr = 0;
  if (smem > attr) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(readout_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  ChainArgs a;
  a.Z = Z; a.ldz = ldz; a.K0 = K0; a.G = G;
  a.P0 = P0; a.b0 = b0; a.P1 = P1; a.b1 = b1; a.P2 = P2; a.b2 = b2; a.P3 = P3; a.b3 = b3; a.out = out; a.g_dev = g_dev;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  readout_chain_kernel<<<(G + ROWS - 1) / ROWS, THREADS, smem, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_internal_count_head_fused(const float* emb_t, int G, const float* emb_q, int Q, const float* W1a, const float* W1b,
                                    const float* b1, const float* w2, const float* b2, float* pred, float* count,
                                    float* Bq, const int32_t* g_dev, cudaStream_t s) {
  if (G == 0 || Q == 0) return DESCO_OK;
  if (Q > QROWS) return DESCO_ERANGE;
  if (!Bq) return DESCO_EINVAL;
  const size_t smem = ((size_t)ROWS * LD64 + ROWS * LD256 + QROWS * LD256 + H4) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(count_head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  HeadArgs a;
  a.emb_t = emb_t; a.G = G; a.Bq = Bq; a.Q = Q;
  a.W1a = W1a; a.w2 = w2; a.b2 = b2; a.pred = pred; a.count = count; a.g_dev = g_dev;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s, 2);
  count_head_query_kernel<<<Q, H4, 0, s>>>(emb_q, W1b, b1, Bq);
  count_head_fused_kernel<<<(G + ROWS - 1) / ROWS, THREADS, smem, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}
