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
constexpr int WG_TOTAL = WG_B3 + 4;

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

// layer 0: scalar gated SpMV for all queries; one warp per node, lane = query
__global__ void gossip_layer0_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int node_begin,
                                     int node_end, const float* __restrict__ x, int Q, const float* __restrict__ qvec,
                                     float4* __restrict__ S4) {
  const int i = node_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= node_end) return;
  const int lane = lane_id();
  const int eb = rowptr[i], ee = rowptr[i + 1];
  for (int q0 = 0; q0 < Q; q0 += 32) {
    const int q = q0 + lane;
    float s_lt = 0.f, s_gt = 0.f;
    int d_lt = 0;
    for (int e = eb; e < ee; ++e) {
      const int j = col[e];
      const float c = (q < Q) ? x[(size_t)j * Q + q] : 0.f;
      if (j < i) { s_lt += c; ++d_lt; } else { s_gt += c; }
    }
    if (q < Q) {
      const float g0 = qvec[(size_t)q * QV + 3 * F], g1 = qvec[(size_t)q * QV + 3 * F + 1];
      const float dl = (float)d_lt, dg = (float)(ee - eb - d_lt);
      float4 o;
      o.x = g0 * dl + (1.f - g0) * dg;      // dmix (layer 0)
      o.y = g0 * s_lt + (1.f - g0) * s_gt;  // smix (layer 0)
      o.z = x[(size_t)i * Q + q];           // c_i
      o.w = g1 * dl + (1.f - g1) * dg;      // dmix (layer 1)
      S4[(size_t)i * Q + q] = o;
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
    const float4* __restrict__ S4, int Q, const float* __restrict__ qvec, const float* __restrict__ wg,
    float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* X = smem;                 // [TM][LDX]: cols 0-63 u / y1, 64-127 x1 / y2, 128-191 x2
  float* sW = X + TM * LDX;        // [128][64]
  float* s_c = sW + 2 * F * F;     // [TM]
  float* s_d1 = s_c + TM;          // [TM]

  const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
  const int ty = tid >> 4, tx = tid & 15;
  const int q = blockIdx.x % Q;
  const int i0 = node_begin + (blockIdx.x / Q) * TM;
  const float* qv = qvec + (size_t)q * QV;
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
      out[(size_t)(i0 + r) * Q + q] = s_c[r] + v + wg[WG_B3];  // neigh_pred + gossip_pred (lightning_model.py:625)
  }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" {

int64_t desco_gossip_weight_floats(void) { return WG_TOTAL; }
int64_t desco_gossip_query_weight_floats(void) { return WQ_TOTAL; }

int64_t desco_gossip_workspace_bytes(int32_t num_nodes, int32_t num_queries) {
  return (int64_t)(align_up((size_t)num_queries * QV * 4) + align_up((size_t)num_nodes * num_queries * sizeof(float4)));
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

int desco_gossip_layer0(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* x,
                        int32_t num_queries, const float* qvec, float* s4, void* stream) {
  if (node_begin < 0 || node_end < node_begin || num_queries < 0) return DESCO_EINVAL;
  if (node_end == node_begin || num_queries == 0) return DESCO_OK;
  if (!rowptr || !col || !x || !qvec || !s4) return DESCO_EINVAL;
  const long long threads = (long long)(node_end - node_begin) * 32;
  DescoProfScope prof(DESCO_PROF_GOSSIP_L0, (cudaStream_t)stream);
  gossip_layer0_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      rowptr, col, node_begin, node_end, x, num_queries, qvec, reinterpret_cast<float4*>(s4));
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_layer1(const int32_t* rowptr, const int32_t* col, int32_t node_begin, int32_t node_end, const float* s4,
                        int32_t num_queries, const float* qvec, const float* w_gossip, float* out, int32_t precision,
                        void* stream) {
  if (node_begin < 0 || node_end < node_begin || num_queries < 0) return DESCO_EINVAL;
  if (precision != DESCO_PRECISION_FP32) return DESCO_EINVAL;
  if (node_end == node_begin || num_queries == 0) return DESCO_OK;
  if (!rowptr || !col || !s4 || !qvec || !w_gossip || !out) return DESCO_EINVAL;
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
      rowptr, col, node_begin, node_end, reinterpret_cast<const float4*>(s4), num_queries, qvec, w_gossip, out);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
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
  int rc;
  if ((rc = desco_gossip_prepare_queries(query_emb, num_queries, w_gossip_query, qvec, out_gates, stream))) return rc;
  if ((rc = desco_gossip_layer0(rowptr, col, 0, num_nodes, x, num_queries, qvec, s4, stream))) return rc;
  return desco_gossip_layer1(rowptr, col, 0, num_nodes, s4, num_queries, qvec, w_gossip, out, precision, stream);
}

}  // extern "C"
