// Gossip propagation forward, sm_100a (fp32 parity path).
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/gnn_model.py:280-359   GossipConv (gate MLP, per-EDGE lin_com, gated sum, lin_update)
//   subgraph_counting/gnn_model.py:231-260   GOSSIP branch of BaseGNNCore.forward (query-embedding concat, 2 layers)
//   subgraph_counting/gnn_model.py:102-103   post_mp on every node
//   subgraph_counting/lightning_model.py:613-628  GossipCountingModel.graph_to_count (Python loop over queries)
//
// Formulation (exact in real arithmetic; DESIGN.md "Gossip kernels"):
//   * all Q queries are processed in one pass (the reference loops over them);
//   * lin_com is applied per NODE, not per edge; the j<i / j>i gating is a split of the sorted adjacency list;
//   * layer 0 is rank-structured: x0_i = [qe ; w_pre c_i + b_pre], so its gated aggregate needs only the scalars
//       deg_<, deg_>, s_< = sum_{j<i} c_j, s_> = sum_{j>i} c_j  and  x1_i = relu(dmix*alpha_q + smix*beta + gamma_q + c_i*delta);
//   * layer 1 never gathers 64-wide rows: a neighbour's x1_j is RECOMPUTED from its 3 scalars (16 B per edge per query
//     instead of 256 B), then x2, post_mp run as a register-tiled FFMA GEMM chain on 64-node tiles held in shared memory.
#include "common.cuh"
#include "tc05.cuh"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;
constexpr int TM = 64;
constexpr int LDX = 3 * F + 4;
constexpr int THREADS = 256;
constexpr int QV = 4 * F;  // per-query vector block: alpha | gamma | eta | {g0, g1, pad...}

// blob offsets (floats) -------------------------------------------------------------------------------------------
// w_gossip_query: what turns a query embedding into its per-query vectors
constexpr int WQ_MALPHA = 0;                    // [64][64] K-major
constexpr int WQ_CALPHA = WQ_MALPHA + F * F;    // [64]
constexpr int WQ_MGAMMA = WQ_CALPHA + F;
constexpr int WQ_CGAMMA = WQ_MGAMMA + F * F;
constexpr int WQ_META = WQ_CGAMMA + F;
constexpr int WQ_CETA = WQ_META + F * F;
constexpr int WQ_GATE = WQ_CETA + F;            // per layer: W1[64][64] K-major, b1[64], w2[64], b2[1] (+3 pad)
constexpr int WQ_GATE_STRIDE = F * F + F + F + 4;
constexpr int WQ_TOTAL = WQ_GATE + 2 * WQ_GATE_STRIDE;
// w_gossip: query independent
constexpr int WG_BETA = 0, WG_DELTA = F, WG_THETA = 2 * F, WG_V1 = 3 * F, WG_BUP1 = 4 * F;
constexpr int WG_WX2 = 5 * F;                   // [128][64]
constexpr int WG_WY1 = WG_WX2 + 2 * F * F;      // [128][64]
constexpr int WG_P1 = WG_WY1 + 2 * F * F;       // [64][64]
constexpr int WG_B1 = WG_P1 + F * F;            // [64]
constexpr int WG_P2 = WG_B1 + F;                // [64][256]
constexpr int WG_B2 = WG_P2 + F * 4 * F;        // [256]
constexpr int WG_P3 = WG_B2 + 4 * F;            // [256]
constexpr int WG_B3 = WG_P3 + 4 * F;            // [1] (+3 pad)
constexpr int WG_FP32 = WG_B3 + 4;              // end of the fp32 part
// tensor-core operand images of the four GEMMs of the layer-1 / post_mp chain (tcpack.pack_b_operand per 64-wide K block:
// bf16 hi rows then lo rows, 128 B per row, SWIZZLE_128B), appended to the same blob as raw bytes
constexpr int IMG_W1 = 0;                       // Wx2  N = 64,  K = 128: [kb0 hi | kb0 lo | kb1 hi | kb1 lo], 8 KB each
constexpr int IMG_W2 = IMG_W1 + 4 * F * 128;    // Wy1  N = 64,  K = 128
constexpr int IMG_P1 = IMG_W2 + 4 * F * 128;    // P1   N = 64,  K = 64:  [hi | lo]
constexpr int IMG_P2 = IMG_P1 + 2 * F * 128;    // P2   N = 256, K = 64:  [hi | lo], 32 KB each
constexpr int IMG_BYTES = IMG_P2 + 2 * 4 * F * 128;
constexpr int WG_TC = (WG_FP32 + 3) / 4 * 4;    // float offset of the images (16-byte aligned)
constexpr int WG_TOTAL = WG_TC + IMG_BYTES / 4;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// one CTA (64 threads) per query: gates (gnn_model.py:294-301,340) and the query-dependent layer-0 / post_mp vectors
__global__ void gossip_query_kernel(const float* __restrict__ qemb, int Q, const float* __restrict__ wq,
                                    float* __restrict__ qvec, float* __restrict__ out_gates) {
  __shared__ float s_q[F];
  __shared__ float s_h[F];
  const int q = blockIdx.x, t = threadIdx.x;
  s_q[t] = qemb[(size_t)q * F + t];
  __syncthreads();
  auto matvec = [&](const float* M, const float* c) {
    float acc = c[t];
    for (int k = 0; k < F; ++k) acc = fmaf(s_q[k], M[k * F + t], acc);
    return acc;
  };
  float* o = qvec + (size_t)q * QV;
  o[t] = matvec(wq + WQ_MALPHA, wq + WQ_CALPHA);
  o[F + t] = matvec(wq + WQ_MGAMMA, wq + WQ_CGAMMA);
  o[2 * F + t] = matvec(wq + WQ_META, wq + WQ_CETA);
  for (int l = 0; l < 2; ++l) {
    const float* g = wq + WQ_GATE + l * WQ_GATE_STRIDE;
    __syncthreads();
    s_h[t] = sigmoidf_(matvec(g, g + F * F)) * g[F * F + F + t];
    __syncthreads();
    if (t == 0) {
      float acc = g[F * F + 2 * F];
      for (int k = 0; k < F; ++k) acc += s_h[k];
      float gate = sigmoidf_(acc);
      gate = gate > 0.f ? gate : 0.01f * gate;  // the trailing nn.LeakyReLU() (identity on (0,1))
      o[3 * F + l] = gate;
      if (out_gates) out_gates[(size_t)l * Q + q] = gate;
    }
  }
}

// s4 layout: the queries are cut into groups of QG consecutive queries; group g is a dense [n_rows][qc_g] block of float4
// (qc_g = min(QG, Q - g*QG)) at float4 offset g*QG*n_rows.  QG >= Q is the plain [N][Q] layout.  A sharded forward
// all-gathers and consumes one group at a time (desco_b200/distributed.py), so the blocks must be contiguous.
__device__ __forceinline__ size_t s4_index(int i, int q, int Q, int QG, long long n_rows) {
  const int q0 = (q / QG) * QG;
  const int qc = min(QG, Q - q0);
  return (size_t)q0 * (size_t)n_rows + (size_t)i * qc + (q - q0);
}

// hub rows of the layer-0 sweep: a power-law target has rows of 10^4 neighbours, which one warp walks in milliseconds -
// invisible on one GPU, the whole tail of the kernel once the rows are sharded 8 ways.  The sweep parks such rows here
// and a second kernel gives each of them a whole CTA.  (One list per device: launches of this library on different
// streams of one device must not overlap in the gossip forward.)
constexpr int L0_HUB_DEG = 2048;
constexpr int L0_HUB_CAP = 8192;
__device__ int g_l0_hub_count;
__device__ int g_l0_hub_rows[L0_HUB_CAP];

__device__ __forceinline__ void layer0_store(int i, int q, int Q, int QG, long long n_rows, const float* __restrict__ x,
                                             const float* __restrict__ qvec, float4* __restrict__ S4, int d_lt, int deg,
                                             float s_lt, float s_gt) {
  const float g0 = qvec[(size_t)q * QV + 3 * F], g1 = qvec[(size_t)q * QV + 3 * F + 1];
  const float dl = (float)d_lt, dg = (float)(deg - d_lt);
  float4 o;
  o.x = g0 * dl + (1.f - g0) * dg;      // dmix (layer 0)
  o.y = g0 * s_lt + (1.f - g0) * s_gt;  // smix (layer 0)
  o.z = x[(size_t)i * Q + q];           // c_i
  o.w = g1 * dl + (1.f - g1) * dg;      // dmix (layer 1)
  S4[s4_index(i, q, Q, QG, n_rows)] = o;
}

// layer 0: scalar gated SpMV for all queries; one warp per node, lane = query
__global__ void gossip_layer0_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int node_begin,
                                     int node_end, const float* __restrict__ x, int Q, const float* __restrict__ qvec,
                                     float4* __restrict__ S4, int QG, long long n_rows) {
  const int i = node_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= node_end) return;
  const int lane = lane_id();
  const int eb = rowptr[i], ee = rowptr[i + 1];
  if (ee - eb > L0_HUB_DEG) {  // park the row for gossip_layer0_hub_kernel (if the list is full it is walked here after all)
    int slot = 0;
    if (lane == 0) slot = atomicAdd(&g_l0_hub_count, 1);
    slot = __shfl_sync(FULL_MASK, slot, 0);
    if (slot < L0_HUB_CAP) {
      if (lane == 0) g_l0_hub_rows[slot] = i;
      return;
    }
  }
  for (int q0 = 0; q0 < Q; q0 += 32) {
    const int q = q0 + lane;
    float s_lt = 0.f, s_gt = 0.f;
    int d_lt = 0;
    // 32 adjacency entries per step (one coalesced load), then the 32 neighbours' count rows back to back: the loads are
    // independent, so a hub row (10^4 neighbours, one warp) keeps eight 116-byte row reads in flight instead of one
    for (int base = eb; base < ee; base += 32) {
      const int myj = (base + lane < ee) ? col[base + lane] : 0x7fffffff;
      const int n = min(32, ee - base);
      d_lt += __popc(__ballot_sync(FULL_MASK, myj < i));
#pragma unroll 8
      for (int k = 0; k < n; ++k) {
        const int j = __shfl_sync(FULL_MASK, myj, k);
        const float c = (q < Q) ? __ldg(x + (size_t)j * Q + q) : 0.f;
        if (j < i) s_lt += c; else s_gt += c;
      }
    }
    if (q < Q) layer0_store(i, q, Q, QG, n_rows, x, qvec, S4, d_lt, ee - eb, s_lt, s_gt);
  }
}

// one CTA per parked hub row: the warps take 32-neighbour chunks round robin, partial sums are added in warp order
constexpr int L0H_THREADS = 512;
__global__ void __launch_bounds__(L0H_THREADS) gossip_layer0_hub_kernel(const int32_t* __restrict__ rowptr,
                                                                       const int32_t* __restrict__ col,
                                                                       const float* __restrict__ x, int Q,
                                                                       const float* __restrict__ qvec,
                                                                       float4* __restrict__ S4, int QG, long long n_rows) {
  constexpr int NWH = L0H_THREADS / 32;
  __shared__ float s_part[NWH][2][32];
  __shared__ int s_dlt[NWH];
  const int nh = min(g_l0_hub_count, L0_HUB_CAP);
  const int lane = lane_id(), warp = warp_id();
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int i = g_l0_hub_rows[h];
    const int eb = rowptr[i], ee = rowptr[i + 1];
    for (int q0 = 0; q0 < Q; q0 += 32) {
      const int q = q0 + lane;
      float s_lt = 0.f, s_gt = 0.f;
      int d_lt = 0;
      for (int base = eb + 32 * warp; base < ee; base += 32 * NWH) {
        const int myj = (base + lane < ee) ? col[base + lane] : 0x7fffffff;
        const int n = min(32, ee - base);
        d_lt += __popc(__ballot_sync(FULL_MASK, myj < i));
#pragma unroll 8
        for (int k = 0; k < n; ++k) {
          const int j = __shfl_sync(FULL_MASK, myj, k);
          const float c = (q < Q) ? __ldg(x + (size_t)j * Q + q) : 0.f;
          if (j < i) s_lt += c; else s_gt += c;
        }
      }
      s_part[warp][0][lane] = s_lt;
      s_part[warp][1][lane] = s_gt;
      if (lane == 0) s_dlt[warp] = d_lt;
      __syncthreads();
      if (warp == 0 && q < Q) {
        float a = 0.f, b = 0.f;
        int d = 0;
        for (int w = 0; w < NWH; ++w) { a += s_part[w][0][lane]; b += s_part[w][1][lane]; d += s_dlt[w]; }
        layer0_store(i, q, Q, QG, n_rows, x, qvec, S4, d, ee - eb, a, b);
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ void stage(float* sW, const float* __restrict__ W, int rows, int ldw) {
  // copy a [rows][64] slice (row stride ldw floats) into sW[rows][64]
  for (int i = threadIdx.x; i < rows * (F / 4); i += THREADS) {
    const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
    *reinterpret_cast<float4*>(sW + r * F + c4) = *reinterpret_cast<const float4*>(W + (size_t)r * ldw + c4);
  }
}

__device__ __forceinline__ void zero_acc(float acc[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// layer 1 + post_mp for a tile of 64 nodes x 1 query
__global__ void __launch_bounds__(THREADS, 2) gossip_layer1_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int node_begin, int node_end,
    const float4* __restrict__ S4, int Q, int q_begin, const float* __restrict__ qvec, const float* __restrict__ wg,
    float* __restrict__ out, int ldo) {
  // S4: one query group [.][Q] (Q = queries of the group, q_begin = its first query); out[i * ldo + local query]
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                 // [TM][LDX]: cols 0-63 u / y1, 64-127 x1 / y2, 128-191 x2
  float* sW = X + TM * LDX;        // [128][64]
  float* s_c = sW + 2 * F * F;     // [TM]
  float* s_d1 = s_c + TM;          // [TM]

  const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
  const int ty = tid >> 4, tx = tid & 15;
  const int q = blockIdx.x % Q;
  const int i0 = node_begin + (blockIdx.x / Q) * TM;
  const float* qv = qvec + (size_t)(q_begin + q) * QV;
  const float g1 = qv[3 * F + 1];

  // per-lane constants of x1 = relu(dmix*alpha + smix*beta + gamma + c*delta)
  const float2 alpha = *reinterpret_cast<const float2*>(qv + 2 * lane);
  const float2 gamma = *reinterpret_cast<const float2*>(qv + F + 2 * lane);
  const float2 beta = *reinterpret_cast<const float2*>(wg + WG_BETA + 2 * lane);
  const float2 delta = *reinterpret_cast<const float2*>(wg + WG_DELTA + 2 * lane);
  auto x1_of = [&](const float4 s) {
    float2 r;
    r.x = fmaxf(fmaf(s.x, alpha.x, fmaf(s.y, beta.x, fmaf(s.z, delta.x, gamma.x))), 0.f);
    r.y = fmaxf(fmaf(s.x, alpha.y, fmaf(s.y, beta.y, fmaf(s.z, delta.y, gamma.y))), 0.f);
    return r;
  };

  stage(sW, wg + WG_WX2, 2 * F, F);  // overlaps with the gather below (different smem region)

  // ---- phase A: x1_i and u_i = g1 * sum_{j<i} x1_j + (1-g1) * sum_{j>i} x1_j ----
  for (int r = w; r < TM; r += THREADS / 32) {
    const int i = i0 + r;
    float2 u = make_float2(0.f, 0.f), x1 = u;
    float c = 0.f, d1 = 0.f;
    if (i < node_end) {
      const float4 own = S4[(size_t)i * Q + q];
      x1 = x1_of(own);
      c = own.z;
      d1 = own.w;
      float2 lt = make_float2(0.f, 0.f), gt = lt;
      const int eb = rowptr[i], ee = rowptr[i + 1];
      for (int base = eb; base < ee; base += 32) {
        const int e = base + lane;
        int my_j = -1;
        float4 my_s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < ee) {
          my_j = col[e];
          my_s = S4[(size_t)my_j * Q + q];
        }
        const int n = min(32, ee - base);
        for (int k = 0; k < n; ++k) {
          const int j = __shfl_sync(FULL_MASK, my_j, k);
          float4 s;
          s.x = __shfl_sync(FULL_MASK, my_s.x, k);
          s.y = __shfl_sync(FULL_MASK, my_s.y, k);
          s.z = __shfl_sync(FULL_MASK, my_s.z, k);
          const float2 v = x1_of(s);
          if (j < i) { lt.x += v.x; lt.y += v.y; } else { gt.x += v.x; gt.y += v.y; }
        }
      }
      u.x = g1 * lt.x + (1.f - g1) * gt.x;
      u.y = g1 * lt.y + (1.f - g1) * gt.y;
    }
    *reinterpret_cast<float2*>(X + r * LDX + 2 * lane) = u;
    *reinterpret_cast<float2*>(X + r * LDX + F + 2 * lane) = x1;
    if (lane == 0) {
      s_c[r] = c;
      s_d1[r] = d1;
    }
  }
  __syncthreads();

  float acc[4][4];
  // ---- x2 = relu([u | x1] . Wx2 + d1mix * v1 + b_up1)  (gnn_model.py:341-348 for layer 1) ----
  zero_acc(acc);
  tile_gemm<2 * F, LDX>(X, sW, ty, tx, acc);
  {
    const float4 v1 = *reinterpret_cast<const float4*>(wg + WG_V1 + tx * 4);
    const float4 b = *reinterpret_cast<const float4*>(wg + WG_BUP1 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float d1 = s_d1[r];
      float4 o;
      o.x = fmaxf(acc[i][0] + d1 * v1.x + b.x, 0.f);
      o.y = fmaxf(acc[i][1] + d1 * v1.y + b.y, 0.f);
      o.z = fmaxf(acc[i][2] + d1 * v1.z + b.z, 0.f);
      o.w = fmaxf(acc[i][3] + d1 * v1.w + b.w, 0.f);
      *reinterpret_cast<float4*>(X + r * LDX + 2 * F + tx * 4) = o;
    }
  }
  __syncthreads();
  stage(sW, wg + WG_WY1, 2 * F, F);
  __syncthreads();
  // ---- y1 = leaky_0.1([x1 | x2] . Wy1 + eta_q + c_i * theta)   (post_mp[0..2], x0 part folded into eta / theta) ----
  zero_acc(acc);
  tile_gemm<2 * F, LDX>(X + F, sW, ty, tx, acc);
  {
    const float4 eta = *reinterpret_cast<const float4*>(qv + 2 * F + tx * 4);
    const float4 th = *reinterpret_cast<const float4*>(wg + WG_THETA + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const float c = s_c[r];
      float o[4] = {acc[i][0] + eta.x + c * th.x, acc[i][1] + eta.y + c * th.y, acc[i][2] + eta.z + c * th.z,
                    acc[i][3] + eta.w + c * th.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = o[j] > 0.f ? o[j] : 0.1f * o[j];
      *reinterpret_cast<float4*>(X + r * LDX + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
  __syncthreads();
  stage(sW, wg + WG_P1, F, F);
  __syncthreads();
  // ---- y2 = relu(y1 . P1 + b1) ----
  zero_acc(acc);
  tile_gemm<F, LDX>(X, sW, ty, tx, acc);
  {
    const float4 b = *reinterpret_cast<const float4*>(wg + WG_B1 + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      *reinterpret_cast<float4*>(X + r * LDX + F + tx * 4) =
          make_float4(fmaxf(acc[i][0] + b.x, 0.f), fmaxf(acc[i][1] + b.y, 0.f), fmaxf(acc[i][2] + b.z, 0.f),
                      fmaxf(acc[i][3] + b.w, 0.f));
    }
  }
  // ---- y4 = relu(y2 . P2 + b2) . p3 + b3, 64 output columns at a time, never materialising the 256-wide y3 ----
  float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int nb = 0; nb < 4; ++nb) {
    __syncthreads();
    stage(sW, wg + WG_P2 + nb * F, F, 4 * F);
    __syncthreads();
    zero_acc(acc);
    tile_gemm<F, LDX>(X + F, sW, ty, tx, acc);
    const float4 b = *reinterpret_cast<const float4*>(wg + WG_B2 + nb * F + tx * 4);
    const float4 p = *reinterpret_cast<const float4*>(wg + WG_P3 + nb * F + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      rowsum[i] += fmaxf(acc[i][0] + b.x, 0.f) * p.x + fmaxf(acc[i][1] + b.y, 0.f) * p.y +
                   fmaxf(acc[i][2] + b.z, 0.f) * p.z + fmaxf(acc[i][3] + b.w, 0.f) * p.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float v = rowsum[i];
    v += __shfl_xor_sync(FULL_MASK, v, 8);
    v += __shfl_xor_sync(FULL_MASK, v, 4);
    v += __shfl_xor_sync(FULL_MASK, v, 2);
    v += __shfl_xor_sync(FULL_MASK, v, 1);
    const int r = ty * 4 + i;
    if (tx == 0 && i0 + r < node_end)
      out[(size_t)(i0 + r) * ldo + q] = s_c[r] + v + wg[WG_B3];  // neigh_pred + gossip_pred (lightning_model.py:625)
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core variant of layer 1 + post_mp (DESCO_PRECISION_BF16X3).  Per (node, query) the chain is 36.9 k MACs of
// dense GEMM (x2: 128->64, y1: 128->64, y2: 64->64, y3: 64->256) against a few hundred flops of gather, so the FFMA
// kernel above is compute bound on the CUDA cores.  Here the work is split by what bounds it:
//   * gossip_gather_kernel - the gated sweep (dependent col[] -> S4[] loads, ~1 us each): small CTAs at high occupancy.
//     A CTA owns a tile = 128 consecutive nodes x 1 query, recomputes x1_j per neighbour (warp-level tf32 MMA blocks,
//     see the kernel) and writes the tile's u and x1 rows to a staging buffer AS the bf16 hi/lo SWIZZLE_128B operand
//     images the tensor core wants (64 KB per tile);
//   * gossip_chain_kernel - persistent, one CTA per SM: the four weight matrices stay resident in shared memory as
//     operand images (144 KB, bulk-copied once), a tile's two operand slots arrive by bulk async copy (the next tile's
//     are issued as soon as a slot is dead, under the running tile's epilogues), the four GEMMs run on tcgen05 (M = 128,
//     accumulators in TMEM, three hi/lo passes = fp32-grade products) and every epilogue reads its accumulator with
//     tcgen05.ld, applies bias / rank-1 terms / activation and writes the next GEMM's A operand in place of a dead one.
// The staging buffer is walked in chunks of tiles, so its size is bounded whatever the graph.
// ------------------------------------------------------------------------------------------------------------------
namespace gtc {
// packed fp32x2 add (one issue slot for two lanes; IEEE rounding as the scalar form)
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
constexpr int TR = 128;
constexpr int SLOT = 2 * TR * 128;               // one operand slot: hi (16 KB) | lo (16 KB)
constexpr int TILE_CD = 2 * TR * 4;              // c[TR], d1[TR]
constexpr int TILE_BYTES = 2 * SLOT + TILE_CD;   // staging per tile: u slot | x1 slot | c | d1
constexpr int HUB_DEG = 512;                     // rows with more neighbours are gathered by the whole CTA

// ---------------------------------------------------------------- gather ----------------------------------------
// u_i = g1 * sum_{j<i} x1_j + (1-g1) * sum_{j>i} x1_j with x1_j = relu(a_j . W), a_j = (dmix_j, smix_j, c_j, 1) and W the
// 4 x 64 matrix [alpha_q; beta; delta; gamma_q].  Both gates are positive, so the gate goes inside the relu:
// u_i = sum_j relu((w_j a_j) . W), w_j = g1 or 1-g1 - ONE sum, no j<i / j>i split.  The 4-deep dot products run on the
// warp-level tensor path (mma.sync m16n8k8, tf32; M = 16 features, N = 8 neighbour records): the K = 8 slots hold the
// tf32 hi and lo parts of the weights, the scalars are applied as their hi part and then their lo part, so all four
// partial products are formed and accumulated in fp32 (operands carry 20+ bits: 1e-6 relative, inside the 1e-4 budget).
// What is left on the CUDA cores is relu + add.
//
// The eight columns of an MMA block belong to EIGHT DIFFERENT ROWS (an "octet"; one neighbour of each per block): lane
// (g, t) of the fragment layout brings scalar t of the next neighbour of row g - no shuffles to build the operand - and
// the accumulator columns 2t, 2t+1 it receives are the running sums of rows 2t, 2t+1 - no reduction across lanes, and
// every per-row cost (own record, operand-image stores) is paid once per eight rows.  The rows of the tile are
// counting-sorted by degree so that the rows of an octet are equally long; a row beyond OCT_DEG neighbours is walked by
// a warp alone as eight interleaved sub-rows (sums folded at the end), a row beyond HUB_DEG by the whole CTA.
//
// The sweep is bound by the latency of the dependent col[] -> S4[] loads, so a warp brings a whole SLAB (8 sub-rows x
// up to OCT_DEG neighbours) at once: the adjacency entries with one load per lane and step, then every neighbour's
// 16-byte record with cp.async straight into the warp's shared-memory slab - up to 264 records in flight per warp with
// no register held - and only then runs the tensor-core blocks out of shared memory.
// A row's sum depends only on its own degree class, never on the tile it sits in (sharded == unsharded, bit for bit).
constexpr int G_THREADS = 128;
constexpr int G_NW = G_THREADS / 32;
constexpr int OCT_DEG = 32;             // neighbours per sub-row of a slab
constexpr int G_NBUCKET = OCT_DEG + 3;  // hub | wide | degree OCT_DEG ... 0
constexpr int VB_LD = 140;              // floats per sub-row: OCT_DEG records + the row's own + pad (140 mod 32 = 12: the
                                        // 32 lanes of a block read 32 different banks)
constexpr int JB_LD = OCT_DEG + 1;

__device__ __forceinline__ uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// d = A(16x8, row) . B(8x8, col) + c.  bb holds the lane's B fragment as a register pair: both K halves carry the same
// value (K slots 0-3 meet the hi parts of the weights, slots 4-7 their lo parts); passing the pair as one 64-bit operand
// lets the four MMAs of a block share it (two 32-bit operands made ptxas rebuild the pair before every MMA).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint64_t bb, const float (&c)[4]) {
  asm volatile(
      "{\n\t.reg .b32 b0, b1;\n\tmov.b64 {b0, b1}, %8;\n\t"
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {b0,b1}, {%9,%10,%11,%12};\n\t}"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "l"(bb), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ uint64_t dup64(uint32_t x) {
  uint64_t r;
  asm volatile("mov.b64 %0, {%1, %1};" : "=l"(r) : "r"(x));  // volatile: one pair per block, never rematerialised per MMA
  return r;
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc05::smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// 8 consecutive fp32 -> one 16-byte chunk of bf16 hi and one of lo
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(G_THREADS, 5) gossip_gather_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int node_begin, int node_end,
    const float4* __restrict__ S4, int Q, int q_begin, const float* __restrict__ qvec, const float* __restrict__ wg,
    long long tile0, uint8_t* __restrict__ stage) {
  __shared__ __align__(16) float s_vb[G_NW][8][VB_LD];  // per warp: the slab's records [sub-row][neighbour][4 scalars]
  __shared__ int s_jb[G_NW][8][JB_LD];                  // ... and their node ids
  __shared__ float4 s_part[G_NW][8][2];                 // hub rows: per-warp partial sums, [g][8 features]
  __shared__ int s_eb[TR], s_dg[TR];                    // first adjacency entry and degree of the tile's rows
  __shared__ int s_hist[G_NBUCKET + 1];
  __shared__ uint8_t s_order[TR];                       // tile rows in descending degree order
  __shared__ int s_next, s_nhub, s_nwide;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;  // mma fragment coordinates = (sub-row, scalar / entry phase) of the loader
  const long long tile = tile0 + blockIdx.x;
  const int q = (int)(tile % Q);
  const int i0 = node_begin + (int)(tile / Q) * TR;
  uint8_t* gA0 = stage + (size_t)blockIdx.x * TILE_BYTES;  // u
  uint8_t* gA1 = gA0 + SLOT;                               // x1
  float* g_c = reinterpret_cast<float*>(gA1 + SLOT);
  float* g_d1 = g_c + TR;
  const float* qv = qvec + (size_t)(q_begin + q) * QV;
  const float g1 = qv[3 * F + 1], g1c = 1.f - g1;
  const char* Srec = reinterpret_cast<const char*>(S4) + q * 16;  // record of query q of node 0
  const uint32_t rec_bytes = (uint32_t)Q * 16u;                   // one node's records of the group
  float* vrow = s_vb[warp][g];
  int* jrow = s_jb[warp][g];

  // ---- tile set-up: row extents into shared memory, counting sort of the rows by degree class ----
  if (tid <= G_NBUCKET) s_hist[tid] = 0;
  if (tid == 0) s_next = G_NW;
  __syncthreads();
  int bucket = 0, slot = 0;
  if (tid < TR) {
    const int i = i0 + tid;
    int eb = 0, deg = 0;
    if (i < node_end) { eb = __ldg(rowptr + i); deg = __ldg(rowptr + i + 1) - eb; }
    s_eb[tid] = eb;
    s_dg[tid] = deg;
    bucket = deg > HUB_DEG ? 0 : deg > OCT_DEG ? 1 : 2 + OCT_DEG - deg;
    slot = atomicAdd(&s_hist[bucket], 1);
  }
  __syncthreads();
  if (warp == 0) {  // exclusive scan of the bucket counts, two buckets per lane
    static_assert(G_NBUCKET <= 64, "two buckets per lane");
    int cnt[2], sum = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = 2 * lane + k;
      cnt[k] = b < G_NBUCKET ? s_hist[b] : 0;
      sum += cnt[k];
    }
    if (lane == 0) { s_nhub = cnt[0]; s_nwide = cnt[1]; }
    int run = warp_incl_scan(sum) - sum;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int b = 2 * lane + k;
      if (b < G_NBUCKET) s_hist[b] = run;
      run += cnt[k];
    }
  }
  __syncthreads();
  if (tid < TR) s_order[s_hist[bucket] + slot] = (uint8_t)tid;
  __syncthreads();
  const int n_hub = s_nhub, n_wide = s_nwide;
  const int n_items = n_wide + ((TR - n_hub - n_wide + 7) >> 3);

  // A fragments.  Row g of feature block mb = feature 8 g + 2 mb, row g + 8 = feature 8 g + 2 mb + 1: a lane ends up
  // with the eight consecutive features 8 g ... 8 g + 7 of its rows = one 16-byte chunk of the operand image.
  const float* wrow = t == 0 ? qv : t == 1 ? wg + WG_BETA : t == 2 ? wg + WG_DELTA : qv + F;  // row t of W
  uint32_t aw[4][4];  // (hi f, hi f+1, lo f, lo f+1)
#pragma unroll
  for (int mb = 0; mb < 4; ++mb)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float w = wrow[8 * g + 2 * mb + h];
      aw[mb][h] = tf32_rna(w);
      aw[mb][2 + h] = __float_as_uint(w - __uint_as_float(aw[mb][h]));  // (the tensor core reads the top 19 bits)
    }

  // Slab load.  The lane brings neighbours k = t, t + 4, ... of sub-row g: adjacency entry e0 + k * bs (while < e_end),
  // then the neighbour's record; lane t = 0 also the record of node i_self (>= 0) into slot OCT_DEG.
  auto load_slab = [&](int e0, int bs, int e_end, int i_self) {
    int j[OCT_DEG / 4];
#pragma unroll
    for (int s = 0; s < OCT_DEG / 4; ++s) {
      const int e = e0 + (t + 4 * s) * bs;
      j[s] = e < e_end ? __ldg(col + e) : -1;
    }
    if (t == 0 && i_self >= 0) cp_async16(vrow + 4 * OCT_DEG, Srec + (uint64_t)(uint32_t)i_self * rec_bytes);
#pragma unroll
    for (int s = 0; s < OCT_DEG / 4; ++s)
      if (j[s] >= 0) {
        jrow[t + 4 * s] = j[s];
        cp_async16(vrow + 4 * (t + 4 * s), Srec + (uint64_t)(uint32_t)j[s] * rec_bytes);
      }
    cp_async_wait_all();
    __syncwarp();
  };
  float2 acc[4][2];  // [feature block][feature f / f + 1]: (column 2t, column 2t + 1)
  auto block = [&](float val, bool add) {  // one MMA block: the lane's B entry is val (hi + lo parts)
    const uint32_t bh = __float_as_uint(val) & 0xffffe000u;  // tf32 hi part by truncation; val - hi is exact
    const uint64_t bbh = dup64(bh), bbl = dup64(__float_as_uint(val - __uint_as_float(bh)));
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      const float zero[4] = {0.f, 0.f, 0.f, 0.f};
      float d1[4], d[4];
      mma_tf32(d1, aw[mb], bbh, zero);
      mma_tf32(d, aw[mb], bbl, d1);
      const float2 r0 = make_float2(fmaxf(d[0], 0.f), fmaxf(d[1], 0.f)), r1 = make_float2(fmaxf(d[2], 0.f), fmaxf(d[3], 0.f));
      acc[mb][0] = add ? add2(acc[mb][0], r0) : r0;
      acc[mb][1] = add ? add2(acc[mb][1], r1) : r1;
    }
  };
  // nblk blocks (warp-uniform) out of the slab; the lane's sub-row has n_own neighbours, gated against node i_own
  auto consume_slab = [&](int nblk, int n_own, int i_own) {
#pragma unroll 2
    for (int k = 0; k < nblk; ++k) {
      const int j = jrow[k];
      const float v = t == 3 ? 1.f : vrow[4 * k + t];
      block(k < n_own ? v * (j < i_own ? g1 : g1c) : 0.f, true);
    }
    __syncwarp();
  };
  auto clear_acc = [&]() {
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) acc[mb][0] = acc[mb][1] = make_float2(0.f, 0.f);
  };
  // a row walked as eight interleaved sub-rows: slabs first, first + step, ... of its adjacency
  auto walk_wide = [&](int eb, int deg, int first, int step, int i, bool with_self) {
    clear_acc();
    for (int e0 = eb + first * 8 * OCT_DEG; e0 < eb + deg; e0 += step * 8 * OCT_DEG) {
      const int left = eb + deg - e0;  // entries from this slab on
      load_slab(e0 + g, 8, eb + deg, with_self && g == 0 && e0 == eb ? i : -1);
      consume_slab(min(OCT_DEG, (left + 7) >> 3), min(OCT_DEG, max(0, (left - g + 7) >> 3)), i);
    }
#pragma unroll
    for (int mb = 0; mb < 4; ++mb)  // the sixteen column sums of a feature -> acc[..][..].x of every lane
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = acc[mb][h].x + acc[mb][h].y;
        v += __shfl_xor_sync(FULL_MASK, v, 1);
        v += __shfl_xor_sync(FULL_MASK, v, 2);
        acc[mb][h].x = v;
      }
  };
  auto store8 = [&](uint8_t* img, int r, bool second) {  // features 8 g ... 8 g + 7 of tile row r from acc (.x or .y)
    float v[8];
#pragma unroll
    for (int mb = 0; mb < 4; ++mb) {
      v[2 * mb] = second ? acc[mb][0].y : acc[mb][0].x;
      v[2 * mb + 1] = second ? acc[mb][1].y : acc[mb][1].x;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    const uint32_t off = (uint32_t)r * 128u + ((uint32_t)(g ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(img + off) = hi;
    *reinterpret_cast<uint4*>(img + TR * 128 + off) = lo;
  };
  // the rows' own records as one more block (gate 1, nothing summed): x1_i, c_i, d1_i.  r_col: the tile row of the lane's
  // column (< 0: none), s_own: scalar t of its own record; r_a / r_b: tile rows of accumulator columns 2t, 2t + 1
  auto finish = [&](int r_col, float s_own, int r_a, int r_b) {
    if (r_a >= 0) store8(gA0, r_a, false);
    if (r_b >= 0) store8(gA0, r_b, true);
    if (r_col >= 0) {
      if (t == 2) g_c[r_col] = s_own;
      if (t == 3) g_d1[r_col] = s_own;
    }
    block(r_col < 0 ? 0.f : t == 3 ? 1.f : s_own, false);
    if (r_a >= 0) store8(gA1, r_a, false);
    if (r_b >= 0) store8(gA1, r_b, true);
  };

  // ---- work items dealt by a ticket: the wide rows first (longest work), then the octets in descending degree ----
#pragma unroll 1
  for (int item = warp; item < n_items;) {
    if (item < n_wide) {
      const int r = s_order[n_hub + item];
      walk_wide(s_eb[r], s_dg[r], 0, 1, i0 + r, true);
      const float s_own = vrow[4 * OCT_DEG + t];  // (sub-row 0 holds it; other lanes read a dead slot)
      finish(g == 0 ? r : -1, s_own, t == 0 ? r : -1, -1);
    } else {
      const int base = n_hub + n_wide + 8 * (item - n_wide);
      const int r = base + g < TR ? s_order[base + g] : -1;
      const int i = i0 + r;
      const int eb = r >= 0 ? s_eb[r] : 0, deg = r >= 0 ? s_dg[r] : 0;
      const bool real = r >= 0 && i < node_end;
      load_slab(eb, 1, eb + deg, real ? i : -1);
      clear_acc();
      consume_slab(__reduce_max_sync(FULL_MASK, deg), deg, i);
      const float s_own = real ? vrow[4 * OCT_DEG + t] : 0.f;
      finish(r, s_own, base + 2 * t < TR ? s_order[base + 2 * t] : -1, base + 2 * t + 1 < TR ? s_order[base + 2 * t + 1] : -1);
    }
    __syncwarp();
    if (lane == 0) item = atomicAdd(&s_next, 1);
    item = __shfl_sync(FULL_MASK, item, 0);
  }
  // ---- hub rows: slabs dealt over all warps of the CTA, partial sums added in warp order ----
  for (int h = 0; h < n_hub; ++h) {
    const int r = s_order[h], i = i0 + r;
    walk_wide(s_eb[r], s_dg[r], warp, G_NW, i, false);
    if (t == 0) {
      s_part[warp][g][0] = make_float4(acc[0][0].x, acc[0][1].x, acc[1][0].x, acc[1][1].x);
      s_part[warp][g][1] = make_float4(acc[2][0].x, acc[2][1].x, acc[3][0].x, acc[3][1].x);
    }
    __syncthreads();
    if (warp == 0) {
      if (t == 0) {
        float4 a = s_part[0][g][0], b = s_part[0][g][1];
        for (int w = 1; w < G_NW; ++w) {
          const float4 c = s_part[w][g][0], d = s_part[w][g][1];
          a.x += c.x; a.y += c.y; a.z += c.z; a.w += c.w;
          b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
        }
        acc[0][0].x = a.x; acc[0][1].x = a.y; acc[1][0].x = a.z; acc[1][1].x = a.w;
        acc[2][0].x = b.x; acc[2][1].x = b.y; acc[3][0].x = b.z; acc[3][1].x = b.w;
      }
      const float s_own = g == 0 ? __ldg(reinterpret_cast<const float*>(Srec + (uint64_t)(uint32_t)i * rec_bytes) + t) : 0.f;
      finish(g == 0 ? r : -1, s_own, t == 0 ? r : -1, -1);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- GEMM chain ------------------------------------
constexpr int THREADS_TC = 512;
constexpr int ISSUER = THREADS_TC - 32;
constexpr int SM_B = 0;
constexpr int SM_A0 = SM_B + IMG_BYTES;          // u -> x2 -> y1
constexpr int SM_A1 = SM_A0 + SLOT;              // x1 -> y2
constexpr int SM_CD = SM_A1 + SLOT;              // c[TR], d1[TR] of the tile (arrive with the u slot)
constexpr int SM_VEC = SM_CD + TILE_CD;          // v1, b_up1, theta, b1 [64 each], b2, p3 [256 each], eta [64]
constexpr int SM_PART = SM_VEC + (4 * F + 2 * 4 * F + F) * 4;  // [4][TR]
constexpr int SM_BARS = SM_PART + 4 * TR * 4;
constexpr int SM_TOTAL = SM_BARS + 64;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;
static_assert(SMEM_BYTES <= 232448, "gossip tensor-core kernel exceeds the 227 KB shared-memory limit");
static_assert(SM_A0 % 1024 == 0 && SM_A1 % 1024 == 0 && IMG_W2 % 1024 == 0 && IMG_P1 % 1024 == 0 && IMG_P2 % 1024 == 0,
              "UMMA tiles must be 1024-B aligned");
static_assert(SM_CD % 16 == 0 && TILE_BYTES % 16 == 0, "bulk copies need 16-byte alignment");

// phase timing (thread 0 of every CTA adds its clock64 deltas; read with desco_gossip_tc_phase_cycles)
enum { GPH_LOAD = 0, GPH_X2, GPH_Y1, GPH_Y2, GPH_Y4, GPH_SPARE, GPH_COUNT };
__device__ unsigned long long g_phase_cycles[GPH_COUNT];

__global__ void __launch_bounds__(THREADS_TC, 1) gossip_chain_kernel(
    int node_begin, int node_end, int Q, int q_begin, const float* __restrict__ qvec, const float* __restrict__ wg,
    long long tile0, int num_tiles, const uint8_t* __restrict__ stage, float* __restrict__ out, int ldo,
    int* __restrict__ ticket) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem + SM_B;
  uint8_t* sA0 = smem + SM_A0;
  uint8_t* sA1 = smem + SM_A1;
  float* s_c = reinterpret_cast<float*>(smem + SM_CD);
  float* s_d1 = s_c + TR;
  float* sV1 = reinterpret_cast<float*>(smem + SM_VEC);
  float* sBup1 = sV1 + F;
  float* sTheta = sBup1 + F;
  float* sB1 = sTheta + F;
  float* sB2 = sB1 + F;
  float* sP3 = sB2 + 4 * F;
  float* sEta = sP3 + 4 * F;
  float* s_part = reinterpret_cast<float*>(smem + SM_PART);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);  // [0] weights, [1] MMA done, [2] u slot + c/d1, [3] x1 slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  int* s_next = reinterpret_cast<int*>(tmem_slot + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if ((int)blockIdx.x >= num_tiles) return;

  if (tid == 0) {
    for (int b = 0; b < 4; ++b) tc05::mbar_init(&bars[b], 1);
    tc05::fence_mbar_init();
  }
  if (warp == 0) tc05::tmem_alloc(tmem_slot, 512);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  auto load_u = [&](int t) {  // ISSUER only: u slot + c/d1 of tile t (contiguous in the staging buffer apart from the x1 slot)
    const uint8_t* src = stage + (size_t)t * TILE_BYTES;
    tc05::mbar_arrive_expect_tx(&bars[2], SLOT + TILE_CD);
    tc05::bulk_g2s(sA0, src, 16384, &bars[2]);
    tc05::bulk_g2s(sA0 + 16384, src + 16384, 16384, &bars[2]);
    tc05::bulk_g2s(s_c, src + 2 * SLOT, TILE_CD, &bars[2]);
  };
  auto load_x1 = [&](int t) {
    const uint8_t* src = stage + (size_t)t * TILE_BYTES + SLOT;
    tc05::mbar_arrive_expect_tx(&bars[3], SLOT);
    tc05::bulk_g2s(sA1, src, 16384, &bars[3]);
    tc05::bulk_g2s(sA1 + 16384, src + 16384, 16384, &bars[3]);
  };
  if (tid == ISSUER) {  // all four weight images once per CTA, then the first tile's operands
    const uint8_t* img = reinterpret_cast<const uint8_t*>(wg + WG_TC);
    tc05::mbar_arrive_expect_tx(&bars[0], IMG_BYTES);
    for (int off = 0; off < IMG_BYTES; off += 16384) tc05::bulk_g2s(sB + off, img + off, 16384, &bars[0]);
    load_u(blockIdx.x);
    load_x1(blockIdx.x);
  }
  for (int i = tid; i < 4 * F; i += THREADS_TC) {
    sB2[i] = wg[WG_B2 + i];
    sP3[i] = wg[WG_P3 + i];
    if (i < F) {
      sV1[i] = wg[WG_V1 + i];
      sBup1[i] = wg[WG_BUP1 + i];
      sTheta[i] = wg[WG_THETA + i];
      sB1[i] = wg[WG_B1 + i];
    }
  }
  const float b3 = wg[WG_B3];
  const uint32_t idesc64 = tc05::make_idesc_bf16(TR, F), idesc256 = tc05::make_idesc_bf16(TR, 4 * F);
  uint32_t mphase = 0, lphase = 0;
  bool weights_ready = false;  // meaningful in the ISSUER thread only
  const int qd = warp & 3, cg = warp >> 2;  // TMEM lane quarter / column group of this warp
  const int row = 32 * qd + lane;
  const uint32_t tlane = static_cast<uint32_t>(32 * qd) << 16;

  // one GEMM: D[tmem + dcol] = sum over K blocks (A slot, B image) and the three hi/lo passes
  auto issue = [&](uint32_t dcol, uint32_t idesc, const uint8_t* a0, const uint8_t* b0, int bpart0, const uint8_t* a1,
                   const uint8_t* b1, int bpart1) {
    // a*: A slot (hi at +0, lo at +TR*128); b*: B image of the K block (hi at +0, lo at +bpart bytes)
    bool acc = false;
    for (int pass = 0; pass < 3; ++pass) {
      for (int kb = 0; kb < (a1 ? 2 : 1); ++kb) {
        const uint8_t* a = kb ? a1 : a0;
        const uint8_t* b = kb ? b1 : b0;
        const int bpart = kb ? bpart1 : bpart0;
        const uint64_t da = tc05::make_smem_desc(a + (pass == 1 ? TR * 128 : 0));
        const uint64_t db = tc05::make_smem_desc(b + (pass == 2 ? bpart : 0));
#pragma unroll
        for (int k = 0; k < F / 16; ++k) {
          tc05::mma_bf16(tmem + dcol, da + 2 * k, db + 2 * k, idesc, acc);
          acc = true;
        }
      }
    }
    tc05::mma_commit(&bars[1]);
  };
  auto wait_mma = [&]() {
    if (!tc05::mbar_wait(&bars[1], mphase)) __trap();  // never hang the box: a lost MMA becomes a CUDA error
    mphase ^= 1;
    tc05::fence_after_sync();
  };
  // epilogue of a 64-column GEMM: f(value, column) -> next A operand (slot dst), this warp's 16 columns of its 32 rows
  auto epilogue64 = [&](uint32_t dcol, uint8_t* dst, auto f) {
    float v[16];
    tc05::tmem_ld16(tmem + tlane + dcol + 16 * cg, v);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = f(v[8 * h + i], 16 * cg + 8 * h + i);
      uint4 hi, lo;
      split8(o, hi, lo);
      const uint32_t off = tc05::sw128_offset(row, 16 * cg + 8 * h);
      *reinterpret_cast<uint4*>(dst + off) = hi;
      *reinterpret_cast<uint4*>(dst + TR * 128 + off) = lo;
    }
    tc05::fence_before_sync();
    tc05::fence_proxy_async_smem();
    __syncthreads();
  };

  // tiles are dealt by an atomic ticket (the first one is the CTA's own index): a CTA that starts late - its SM was held
  // by a collective's kernel running beside this one - simply takes fewer tiles
  for (int t = blockIdx.x; t < num_tiles;) {
    const long long tile = tile0 + t;
    const int q = (int)(tile % Q);
    const int i0 = node_begin + (int)(tile / Q) * TR;
    int tn = num_tiles;  // this CTA's next tile (ISSUER only)
    const float* qv = qvec + (size_t)(q_begin + q) * QV;
    __syncthreads();  // previous tile fully retired (s_part, sEta, s_next)
    if (tid == ISSUER) {
      tn = (int)gridDim.x + atomicAdd(ticket, 1);
      *s_next = tn;
    }
    long long tick = clock64();
    auto lap = [&](int phase) {
      if (tid == 0) {
        const long long now = clock64();
        atomicAdd(&g_phase_cycles[phase], (unsigned long long)(now - tick));
        tick = now;
      }
    };
    if (tid < F) sEta[tid] = qv[2 * F + tid];
    // the tile's operands (bulk copies issued one tile ahead); every thread observes the completion itself, so the
    // copied c / d1 words are visible to it
    if (!tc05::mbar_wait(&bars[2], lphase)) __trap();
    const float c = s_c[row], d1 = s_d1[row];
    __syncthreads();  // c / d1 are in registers everywhere: their block may be overwritten by the next tile's copy
    lap(GPH_LOAD);

    // ---- x2 = relu([u | x1] . Wx2 + d1mix * v1 + b_up1)  (gnn_model.py:341-348 for layer 1) ----
    if (tid == ISSUER) {
      if (!weights_ready) {
        if (!tc05::mbar_wait(&bars[0], 0)) __trap();
        weights_ready = true;
      }
      if (!tc05::mbar_wait(&bars[3], lphase)) __trap();
      tc05::fence_after_sync();
      issue(0, idesc64, sA0, sB + IMG_W1, F * 128, sA1, sB + IMG_W1 + 2 * F * 128, F * 128);
    }
    lphase ^= 1;
    wait_mma();
    epilogue64(0, sA0, [&](float a, int n) { return fmaxf(a + d1 * sV1[n] + sBup1[n], 0.f); });  // x2 replaces u
    lap(GPH_X2);
    // ---- y1 = leaky_0.1([x1 | x2] . Wy1 + eta_q + c_i * theta)   (post_mp[0..2], x0 part folded into eta / theta) ----
    if (tid == ISSUER) {
      tc05::fence_after_sync();
      issue(64, idesc64, sA1, sB + IMG_W2, F * 128, sA0, sB + IMG_W2 + 2 * F * 128, F * 128);
    }
    wait_mma();
    epilogue64(64, sA0, [&](float a, int n) {
      const float o = a + sEta[n] + c * sTheta[n];
      return o > 0.f ? o : 0.1f * o;
    });  // y1 replaces x2
    lap(GPH_Y1);
    // ---- y2 = relu(y1 . P1 + b1) ----
    if (tid == ISSUER) {
      tc05::fence_after_sync();
      issue(128, idesc64, sA0, sB + IMG_P1, F * 128, nullptr, nullptr, 0);
    }
    wait_mma();
    if (tid == ISSUER && tn < num_tiles) load_u(tn);  // the u slot is dead: fetch the next tile's under the epilogues below
    epilogue64(128, sA1, [&](float a, int n) { return fmaxf(a + sB1[n], 0.f); });  // y2 replaces x1
    lap(GPH_Y2);
    // ---- y4 = relu(y2 . P2 + b2) . p3 + b3: one N = 256 GEMM, the 256-wide y3 only ever exists in TMEM ----
    if (tid == ISSUER) {
      tc05::fence_after_sync();
      issue(256, idesc256, sA1, sB + IMG_P2, 4 * F * 128, nullptr, nullptr, 0);
    }
    wait_mma();
    if (tid == ISSUER && tn < num_tiles) load_x1(tn);
    {
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[16];
        tc05::tmem_ld16(tmem + tlane + 256 + 64 * cg + 16 * j, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int n = 64 * cg + 16 * j + i;
          part = fmaf(fmaxf(v[i] + sB2[n], 0.f), sP3[n], part);
        }
      }
      s_part[cg * TR + row] = part;
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (cg == 0 && i0 + row < node_end)  // neigh_pred + gossip_pred (lightning_model.py:625)
      out[(size_t)(i0 + row) * ldo + q] =
          c + ((s_part[row] + s_part[TR + row]) + (s_part[2 * TR + row] + s_part[3 * TR + row])) + b3;
    lap(GPH_Y4);
    t = *s_next;  // written before the first barrier of this iteration, overwritten after the next one
  }
  if (tid == ISSUER && !weights_ready) tc05::mbar_wait(&bars[0], 0);  // never exit with a bulk copy in flight
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem, 512);
}

// tiles staged per launch pair (bounds the staging buffer: 8192 tiles = 545 MB)
constexpr long long CHUNK_TILES = 8192;
}  // namespace gtc

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int64_t desco_gossip_weight_floats(void) { return WG_TOTAL; }
int64_t desco_gossip_query_weight_floats(void) { return WQ_TOTAL; }

constexpr int64_t L1_TICKET_BYTES = 256;  // head of the layer-1 workspace: the chain kernel's tile ticket

int64_t desco_gossip_layer1_workspace_bytes(int32_t num_nodes, int32_t num_queries, int32_t precision) {
  if (precision != DESCO_PRECISION_BF16X3 || num_nodes <= 0 || num_queries <= 0) return 0;
  long long tiles = (((long long)num_nodes + gtc::TR - 1) / gtc::TR) * num_queries;
  if (tiles > gtc::CHUNK_TILES) tiles = gtc::CHUNK_TILES;
  return L1_TICKET_BYTES + (int64_t)align_up((size_t)tiles * gtc::TILE_BYTES);
}

int64_t desco_gossip_workspace_bytes(int32_t num_nodes, int32_t num_queries) {
  return (int64_t)(align_up((size_t)num_queries * QV * 4) + align_up((size_t)num_nodes * num_queries * sizeof(float4))) +
         desco_gossip_layer1_workspace_bytes(num_nodes, num_queries, DESCO_PRECISION_BF16X3);
}

int desco_gossip_tc_phase_cycles(uint64_t* out, int32_t reset) {
  if (!out) return DESCO_EINVAL;
  unsigned long long h[gtc::GPH_COUNT];
  DESCO_CUDA_TRY(cudaMemcpyFromSymbol(h, gtc::g_phase_cycles, sizeof(h)));
  for (int i = 0; i < gtc::GPH_COUNT; ++i) out[i] = h[i];
  if (reset) {
    for (int i = 0; i < gtc::GPH_COUNT; ++i) h[i] = 0;
    DESCO_CUDA_TRY(cudaMemcpyToSymbol(gtc::g_phase_cycles, h, sizeof(h)));
  }
  return DESCO_OK;
}

int desco_gossip_prepare_queries(const float* query_emb, int32_t num_queries, const float* w_gossip_query, float* qvec,
                                 float* out_gates, void* stream) {
  if (num_queries < 0) return DESCO_EINVAL;
  if (num_queries == 0) return DESCO_OK;
  if (!query_emb || !w_gossip_query || !qvec) return DESCO_EINVAL;
  desco_count_launches(1);
  gossip_query_kernel<<<num_queries, F, 0, (cudaStream_t)stream>>>(query_emb, num_queries, w_gossip_query, qvec, out_gates);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_layer0_grouped(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end,
                                const float* x, int32_t num_queries, const float* qvec, float* s4, int32_t group_size,
                                int64_t s4_rows, void* stream) {
  if (node_begin < 0 || node_end < node_begin || num_queries < 0 || group_size < 1 || s4_rows < node_end) return DESCO_EINVAL;
  if (node_end == node_begin || num_queries == 0) return DESCO_OK;
  if (!rowptr || !col || !x || !qvec || !s4) return DESCO_EINVAL;
  const long long threads = (long long)(node_end - node_begin) * 32;
  void* hub_count = nullptr;
  DESCO_CUDA_TRY(cudaGetSymbolAddress(&hub_count, g_l0_hub_count));
  DESCO_CUDA_TRY(cudaMemsetAsync(hub_count, 0, sizeof(int), (cudaStream_t)stream));
  DescoProfScope prof(DESCO_PROF_GOSSIP_L0, (cudaStream_t)stream, 2);
  gossip_layer0_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rowptr, col, node_begin, node_end, x, num_queries, qvec, reinterpret_cast<float4*>(s4), group_size,
      (long long)s4_rows);
  gossip_layer0_hub_kernel<<<2 * desco_num_sms(), L0H_THREADS, 0, (cudaStream_t)stream>>>(
      rowptr, col, x, num_queries, qvec, reinterpret_cast<float4*>(s4), group_size, (long long)s4_rows);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_layer0(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* x,
                        int32_t num_queries, const float* qvec, float* s4, void* stream) {
  return desco_gossip_layer0_grouped(rowptr, col, node_begin, node_end, x, num_queries, qvec, s4,
                                     num_queries > 0 ? num_queries : 1, node_end, stream);
}

int desco_gossip_layer1_group(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end,
                              const float* s4_group, int32_t query_begin, int32_t group_queries, const float* qvec,
                              const float* w_gossip, float* out, int32_t out_stride, int32_t precision, void* workspace,
                              int64_t workspace_bytes, void* stream) {
  const int32_t num_queries = group_queries;
  if (node_begin < 0 || node_end < node_begin || num_queries < 0 || query_begin < 0 || out_stride < num_queries)
    return DESCO_EINVAL;
  if (precision != DESCO_PRECISION_FP32 && precision != DESCO_PRECISION_BF16X3) return DESCO_EINVAL;
  if (node_end == node_begin || num_queries == 0) return DESCO_OK;
  if (!rowptr || !col || !s4_group || !qvec || !w_gossip || !out) return DESCO_EINVAL;
  const float4* s4 = reinterpret_cast<const float4*>(s4_group);
  if (precision == DESCO_PRECISION_BF16X3) {
    if (!workspace) return DESCO_EINVAL;
    const long long tc_tiles = (((long long)(node_end - node_begin) + gtc::TR - 1) / gtc::TR) * num_queries;
    const long long cap = (workspace_bytes - L1_TICKET_BYTES) / gtc::TILE_BYTES;  // tiles the staging buffer holds
    if (cap < 1) return DESCO_ENOMEM;
    int* ticket = (int*)workspace;
    uint8_t* stage = (uint8_t*)workspace + L1_TICKET_BYTES;
    static bool tc_attr_set = false;
    if (!tc_attr_set) {
      DESCO_CUDA_TRY(cudaFuncSetAttribute(gtc::gossip_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          gtc::SMEM_BYTES));
      tc_attr_set = true;
    }
    const int sms = desco_num_sms();
    for (long long t0 = 0; t0 < tc_tiles; t0 += cap) {
      const int n = (int)(tc_tiles - t0 < cap ? tc_tiles - t0 : cap);
      DESCO_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(int), (cudaStream_t)stream));
      {
        DescoProfScope prof(DESCO_PROF_GOSSIP_L1, (cudaStream_t)stream, 1);
        gtc::gossip_gather_kernel<<<n, gtc::G_THREADS, 0, (cudaStream_t)stream>>>(
            rowptr, col, node_begin, node_end, s4, num_queries, query_begin, qvec, w_gossip, t0, stage);
      }
      DescoProfScope prof(DESCO_PROF_GOSSIP_CHAIN, (cudaStream_t)stream, 1);
      gtc::gossip_chain_kernel<<<n < sms ? n : sms, gtc::THREADS_TC, gtc::SMEM_BYTES, (cudaStream_t)stream>>>(
          node_begin, node_end, num_queries, query_begin, qvec, w_gossip, t0, n, stage, out, out_stride, ticket);
      DESCO_LAUNCH_CHECK();
    }
    return DESCO_OK;
  }
  const size_t smem = (size_t)(TM * LDX + 2 * F * F + 2 * TM) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(gossip_layer1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const long long tiles = ((long long)(node_end - node_begin) + TM - 1) / TM;
  const long long blocks = tiles * num_queries;
  if (blocks > 0x7fffffffLL) return DESCO_ERANGE;
  DescoProfScope prof(DESCO_PROF_GOSSIP_L1, (cudaStream_t)stream);
  gossip_layer1_kernel<<<(unsigned)blocks, THREADS, smem, (cudaStream_t)stream>>>(
      rowptr, col, node_begin, node_end, s4, num_queries, query_begin, qvec, w_gossip, out, out_stride);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_layer1(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* s4,
                        int32_t num_queries, const float* qvec, const float* w_gossip, float* out, int32_t precision,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  return desco_gossip_layer1_group(rowptr, col, node_begin, node_end, s4, 0, num_queries, qvec, w_gossip, out, num_queries,
                                   precision, workspace, workspace_bytes, stream);
}

int desco_gossip_forward(const int32_t* rowptr, const int32_t* col, int32_t num_nodes, const float* x,
                         int32_t num_queries, const float* query_emb, const float* w_gossip,
                         const float* w_gossip_query, float* out, float* out_gates, void* workspace,
                         int64_t workspace_bytes, int32_t precision, void* stream) {
  if (num_nodes < 0 || num_queries < 0) return DESCO_EINVAL;
  if (num_nodes == 0 || num_queries == 0) return DESCO_OK;
  if (!workspace || workspace_bytes < desco_gossip_workspace_bytes(num_nodes, num_queries)) return DESCO_ENOMEM;
  float* qvec = (float*)workspace;
  float* s4 = (float*)((char*)workspace + align_up((size_t)num_queries * QV * 4));
  char* stage = (char*)s4 + align_up((size_t)num_nodes * num_queries * sizeof(float4));
  const int64_t stage_bytes = workspace_bytes - (int64_t)(stage - (char*)workspace);
  int rc;
  if ((rc = desco_gossip_prepare_queries(query_emb, num_queries, w_gossip_query, qvec, out_gates, stream))) return rc;
  if ((rc = desco_gossip_layer0(rowptr, col, 0, num_nodes, x, num_queries, qvec, s4, stream))) return rc;
  return desco_gossip_layer1(rowptr, col, 0, num_nodes, s4, num_queries, qvec, w_gossip, out, precision, stage, stage_bytes,
                             stream);
}

}  // extern "C"
