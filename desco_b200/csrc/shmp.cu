// SHMP neighborhood counting forward, sm_100a (fp32 parity path).
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/gnn_model.py:362-404  SAGEConv           sum-aggregate then Linear, per relation
//   subgraph_counting/gnn_model.py:230-277  BaseGNNCore.forward as expanded by to_hetero_old (lightning_model.py:371-421)
//   subgraph_counting/gnn_model.py:58-109   BaseGNN.forward    anchor_mlp, global_add_pool, post_mp
//   subgraph_counting/lightning_model.py:176-222  embed_to_count / graph_to_count  (query-conditioned count head)
//
// Formulation (exact in real arithmetic, see DESIGN.md "SHMP kernels"):
//   h'_i = relu( [ sum_{j in N_tri(i), j count} h_j | sum_{j in N_tride(i), j count} h_j | h_i ] . Wfused  + bias
//                + [i adjacent to its canonical node by a triangle/tride edge] * cvec_{tri/tride}[nbh(i)] )
// where Wfused = [U_m W_tri ; U_m W_tride ; U_h] is pre-multiplied on the host (fp64) and cvec = h_canonical . (U_m W_ac_*)
// is one 128-vector per neighborhood.  The skip-concat (576 wide) is never materialised: the pooled sum over count
// rows and the canonical row are accumulated layer by layer.
#include "common.cuh"
#include "shmp_internal.h"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;          // hidden width (config.py:250 neigh_hidden_dim)
constexpr int TM = 64;         // rows per tile
constexpr int KC = 3 * F;      // fused K of one SHMP layer
constexpr int LDA = KC + 4;    // padded smem row stride of the A tile
constexpr int THREADS = 256;

// ------------------------------------------------------------------------------------------------------------------
// plan: per-row metadata derived from the packed batch
// ------------------------------------------------------------------------------------------------------------------
// one thread per neighborhood: the row whose edge to the canonical node SAGEConv's remove_self_loops drops
__global__ void shmp_plan_kernel(const int32_t* __restrict__ nbh_ptr, int G, int hetero, int pyg_batch_size,
                                 int32_t* __restrict__ quirk_row) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  int quirk = -1;
  if (hetero) {
    // SAGEConv.forward runs remove_self_loops on the bipartite count<->canonical relations too (gnn_model.py:389-390):
    // the edge whose per-type local ids coincide inside one collated PyG batch is dropped.  In this layout that is the
    // edge canonical(g) -- first row of g, iff every earlier neighborhood of the PyG batch has exactly two rows.
    const int bs = pyg_batch_size > 0 ? pyg_batch_size : G;
    const int g0 = (g / bs) * bs;
    const int lo = nbh_ptr[g];
    if (pyg_batch_size >= 0 && lo - nbh_ptr[g0] == 2 * (g - g0)) quirk = lo;  // pyg_batch_size < 0: quirk off
  }
  quirk_row[g] = quirk;
}

// one thread per packed row (a neighborhood may hold 10^6 rows: config 5): its neighborhood by binary search, its compact
// count-row slot and how it touches its canonical node
__global__ void shmp_plan_rows_kernel(const int32_t* __restrict__ nbh_ptr, const int32_t* __restrict__ edge_ptr,
                                      const int32_t* __restrict__ edge_col, const uint8_t* __restrict__ edge_tri, int G, int V,
                                      int hetero, const int32_t* __restrict__ quirk_row, int32_t* __restrict__ row_nbh,
                                      int32_t* __restrict__ crow, uint8_t* __restrict__ canon_code) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= V) return;
  int lo = 0, hi = G;  // largest g with nbh_ptr[g] <= r
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (nbh_ptr[mid] <= r) lo = mid; else hi = mid;
  }
  const int g = lo;
  row_nbh[r] = g;
  const int canon = nbh_ptr[g + 1] - 1;
  if (!hetero) {
    crow[r] = r;
    canon_code[r] = 0;
  } else if (r < canon) {
    crow[r - g] = r;
    uint8_t code = 0;
    const int eb = edge_ptr[r], ee = edge_ptr[r + 1];
    if (ee > eb && edge_col[ee - 1] == canon && r != quirk_row[g]) code = edge_tri[ee - 1] ? 1 : 2;  // canon = max row of g
    canon_code[r] = code;
  } else {
    canon_code[r] = 0;
  }
}

// h0 = feat . Wpre + bpre  (gnn_model.py:231; feat == NULL means ZeroNodeFeat, workload.py:431-440)
__global__ void shmp_pre_kernel(const int32_t* __restrict__ nbh_ptr, const int32_t* __restrict__ row_nbh, int V, int hetero,
                                const float* __restrict__ feat, int input_dim, const float* __restrict__ w_pre,
                                float* __restrict__ h0, float* __restrict__ emb_a, int emb_ld) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)V * F) return;
  const int row = (int)(idx / F), f = (int)(idx % F);
  const int g = row_nbh[row];
  const bool canon = hetero && (row == nbh_ptr[g + 1] - 1);
  const float* W = w_pre + (canon ? (size_t)(input_dim + 1) * F : 0);  // [input_dim][F] then bias[F], per node type
  float v = W[(size_t)input_dim * F + f];
  if (feat)
    for (int d = 0; d < input_dim; ++d) v = fmaf(feat[(size_t)row * input_dim + d], W[(size_t)d * F + f], v);
  if (canon) emb_a[(size_t)g * emb_ld + f] = v; else h0[(size_t)row * F + f] = v;
}

// cvec[g] = h_canonical^l[g] . Cw^l   ([F] x [F][2F])
__global__ void shmp_cvec_kernel(const float* __restrict__ emb_a, int emb_ld, int layer, const float* __restrict__ Cw,
                                 int G, float* __restrict__ cvec) {
  __shared__ float s_h[8][F];
  const int w = warp_id(), lane = lane_id();
  const int g = blockIdx.x * 8 + w;
  if (g < G) {
    s_h[w][lane] = emb_a[(size_t)g * emb_ld + layer * F + lane];
    s_h[w][lane + 32] = emb_a[(size_t)g * emb_ld + layer * F + lane + 32];
  }
  __syncwarp();
  if (g >= G) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < F; ++k) {
    const float h = s_h[w][k];
    const float4 c = *reinterpret_cast<const float4*>(Cw + (size_t)k * 2 * F + lane * 4);
    acc[0] = fmaf(h, c.x, acc[0]);
    acc[1] = fmaf(h, c.y, acc[1]);
    acc[2] = fmaf(h, c.z, acc[2]);
    acc[3] = fmaf(h, c.w, acc[3]);
  }
  *reinterpret_cast<float4*>(cvec + (size_t)g * 2 * F + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// ------------------------------------------------------------------------------------------------------------------
// one SHMP layer: persistent CTAs over [count tiles | canonical tiles]
// ------------------------------------------------------------------------------------------------------------------
struct LayerArgs {
  const int32_t* nbh_ptr; const int32_t* edge_ptr; const int32_t* edge_col; const uint8_t* edge_tri;
  const int32_t* row_nbh; const int32_t* crow; const uint8_t* canon_code; const int32_t* quirk_row;
  int G, Vc, hetero, layer, emb_ld;
  const float* h_in; float* h_out;        // [V][F] count rows
  float* emb_a;                           // [G][emb_ld] canonical rows, all layers (skip-concat of the canonical node)
  float* pool;                            // [G][emb_ld] sum over count rows, all layers
  const float* cvec;                      // [G][2F] canonical -> count contribution of this layer
  const float* Wc; const float* bias_c;   // [KC][F], [F]
  const float* Wa; const float* bias_a;   // [KC][F], [F]
};

__device__ __forceinline__ void add2(float2& a, const float2 b) { a.x += b.x; a.y += b.y; }

// sum over the edges of `row`, split by SHMP type, skipping source row `skip`
__device__ __forceinline__ void aggregate_row(const LayerArgs& p, int row, int skip, int lane, float2& at, float2& ad) {
  const int eb = p.edge_ptr[row], ee = p.edge_ptr[row + 1];
  for (int base = eb; base < ee; base += 32) {
    const int e = base + lane;
    const int my_src = (e < ee) ? p.edge_col[e] : -1;
    const int my_tri = (e < ee) ? (int)p.edge_tri[e] : 0;
    const int n = min(32, ee - base);
    for (int j = 0; j < n; ++j) {
      const int s = __shfl_sync(FULL_MASK, my_src, j);
      const int t = __shfl_sync(FULL_MASK, my_tri, j);
      if (s == skip) continue;
      const float2 v = *reinterpret_cast<const float2*>(p.h_in + (size_t)s * F + 2 * lane);
      if (t) add2(at, v); else add2(ad, v);
    }
  }
}

__device__ __forceinline__ void load_weights(float* sW, const float* __restrict__ W) {
  const float4* src = reinterpret_cast<const float4*>(W);
  float4* dst = reinterpret_cast<float4*>(sW);
  for (int i = threadIdx.x; i < KC * F / 4; i += THREADS) dst[i] = src[i];
}

__global__ void __launch_bounds__(THREADS, 2) shmp_layer_kernel(const LayerArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                    // [TM][LDA]
  float* sW = sA + TM * LDA;           // [KC][F]
  int* s_row = reinterpret_cast<int*>(sW + KC * F);  // [TM]
  int* s_g = s_row + TM;
  int* s_code = s_g + TM;

  const int tid = threadIdx.x, lane = lane_id(), w = warp_id();
  const int ty = tid >> 4, tx = tid & 15;
  const int n_ct = (p.Vc + TM - 1) / TM;
  const int n_at = (p.G + TM - 1) / TM;
  int loaded = -1;  // 0 = Wc, 1 = Wa

  for (int tile = blockIdx.x; tile < n_ct + n_at; tile += gridDim.x) {
    const bool ctile = tile < n_ct;
    const bool need_gemm = ctile || p.hetero;
    __syncthreads();  // previous tile's epilogue is done with sA / s_* / sW
    if (need_gemm && loaded != (ctile ? 0 : 1)) {
      load_weights(sW, ctile ? p.Wc : p.Wa);
      loaded = ctile ? 0 : 1;
    }
    if (tid < TM) {
      int row = -1, g = -1, code = 0;
      if (ctile) {
        const int k = tile * TM + tid;
        if (k < p.Vc) {
          row = p.crow[k];
          g = p.row_nbh[row];
          code = p.canon_code[row];
        }
      } else {
        g = (tile - n_ct) * TM + tid;
        if (g < p.G) row = p.nbh_ptr[g + 1] - 1; else g = -1;
      }
      s_row[tid] = row;
      s_g[tid] = g;
      s_code[tid] = code;
    }
    __syncthreads();

    // ---- gather / aggregate into the A tile: [sum tri | sum tride | self] ----
    for (int r = w; r < TM; r += THREADS / 32) {
      const int row = s_row[r], g = s_g[r];
      float2 at = make_float2(0.f, 0.f), ad = at, self = at;
      if (row >= 0) {
        if (ctile) {
          self = *reinterpret_cast<const float2*>(p.h_in + (size_t)row * F + 2 * lane);
          const int canon = p.hetero ? p.nbh_ptr[g + 1] - 1 : -1;
          aggregate_row(p, row, canon, lane, at, ad);
        } else {
          const int lo = p.nbh_ptr[g];
          const int hi = p.hetero ? row : row + 1;  // count rows of g (all rows for single-type graphs)
          float2 ps = make_float2(0.f, 0.f);  // global_add_pool over the count rows of layer `layer` (gnn_model.py:107)
          for (int rr = lo; rr < hi; ++rr) add2(ps, *reinterpret_cast<const float2*>(p.h_in + (size_t)rr * F + 2 * lane));
          *reinterpret_cast<float2*>(p.pool + (size_t)g * p.emb_ld + p.layer * F + 2 * lane) = ps;
          if (p.hetero) {
            self = *reinterpret_cast<const float2*>(p.emb_a + (size_t)g * p.emb_ld + p.layer * F + 2 * lane);
            aggregate_row(p, row, p.quirk_row[g], lane, at, ad);
          }
        }
      }
      if (need_gemm) {
        *reinterpret_cast<float2*>(sA + r * LDA + 2 * lane) = at;
        *reinterpret_cast<float2*>(sA + r * LDA + F + 2 * lane) = ad;
        *reinterpret_cast<float2*>(sA + r * LDA + 2 * F + 2 * lane) = self;
      }
    }
    if (!need_gemm) continue;
    __syncthreads();

    // ---- [TM x KC] . [KC x F] on the FFMA pipe, 4x4 register tile per thread ----
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    tile_gemm<KC, LDA>(sA, sW, ty, tx, acc);

    const float4 b = *reinterpret_cast<const float4*>((ctile ? p.bias_c : p.bias_a) + tx * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty * 4 + i;
      const int row = s_row[r];
      if (row < 0) continue;
      const int g = s_g[r];
      float4 o = make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
      if (ctile) {
        const int code = s_code[r];
        if (code) {
          const float4 c = *reinterpret_cast<const float4*>(p.cvec + (size_t)g * 2 * F + (code - 1) * F + tx * 4);
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
      }
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);  // gnn_model.py:273
      if (ctile) *reinterpret_cast<float4*>(p.h_out + (size_t)row * F + tx * 4) = o;
      else *reinterpret_cast<float4*>(p.emb_a + (size_t)g * p.emb_ld + (p.layer + 1) * F + tx * 4) = o;
    }
  }
}

// pool of the LAST layer's output (the layer kernels pool their input)
__global__ void shmp_pool_last_kernel(const int32_t* __restrict__ nbh_ptr, int G, int hetero, const float* __restrict__ h,
                                      int layer, float* __restrict__ pool, int emb_ld) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const int lane = lane_id();
  const int lo = nbh_ptr[g], hi = nbh_ptr[g + 1] - (hetero ? 1 : 0);
  float2 ps = make_float2(0.f, 0.f);
  for (int r = lo; r < hi; ++r) add2(ps, *reinterpret_cast<const float2*>(h + (size_t)r * F + 2 * lane));
  *reinterpret_cast<float2*>(pool + (size_t)g * emb_ld + layer * F + 2 * lane) = ps;
}

// Homogeneous model (hetero_graph = False, gnn_model.py:74-83): the centre is an ordinary row of its neighborhood (the last
// one) whose skip-concat embedding goes through anchor_mlp before the pooling.  Its per-layer rows are collected here ...
__global__ void shmp_copy_last_rows_kernel(const int32_t* __restrict__ nbh_ptr, int G, const float* __restrict__ h, int layer,
                                           float* __restrict__ emb_a, int emb_ld) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const size_t row = (size_t)nbh_ptr[g + 1] - 1;
  reinterpret_cast<float2*>(emb_a + (size_t)g * emb_ld + layer * F)[lane_id()] = reinterpret_cast<const float2*>(h + row * F)[lane_id()];
}
// ... and taken out of the all-rows pool again: z = (pool - emb_centre) + anchor_mlp(emb_centre)
__global__ void shmp_sub_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] -= b[i];
}

// ------------------------------------------------------------------------------------------------------------------
// dense row-wise layer for the readout MLPs:  Y = act(X . W + b) (+ R)
// ------------------------------------------------------------------------------------------------------------------
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

__global__ void __launch_bounds__(THREADS) dense_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ R,
                                                        int ldr, float* __restrict__ Y, int ldy, int M, int K, int N,
                                                        int act, float slope) {
  __shared__ __align__(16) float sX[TM][F + 4];
  __shared__ __align__(16) float sWt[F][F];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * F;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += F) {
    __syncthreads();
    for (int i = tid; i < TM * F / 4; i += THREADS) {
      const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) v = *reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * ldx + k0 + c4);
      *reinterpret_cast<float4*>(&sX[r][c4]) = v;
    }
    for (int i = tid; i < F * F / 4; i += THREADS) {
      const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
      *reinterpret_cast<float4*>(&sWt[r][c4]) = *reinterpret_cast<const float4*>(W + (size_t)(k0 + r) * N + n0 + c4);
    }
    __syncthreads();
    tile_gemm<F, F + 4>(&sX[0][0], &sWt[0][0], ty, tx, acc);
  }
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) b = *reinterpret_cast<const float4*>(bias + n0 + tx * 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    float o[4] = {acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (act == ACT_RELU) o[j] = fmaxf(o[j], 0.f);
      else if (act == ACT_LEAKY) o[j] = o[j] > 0.f ? o[j] : o[j] * slope;
    }
    if (R) {
      const float4 r = *reinterpret_cast<const float4*>(R + (size_t)m * ldr + n0 + tx * 4);
      o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
    }
    *reinterpret_cast<float4*>(Y + (size_t)m * ldy + n0 + tx * 4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

int dense(const float* X, int ldx, const float* W, const float* bias, const float* R, int ldr, float* Y, int ldy, int M,
          int K, int N, int act, float slope, cudaStream_t s) {
  if (M == 0) return DESCO_OK;
  if (K % F || N % F) return DESCO_EINVAL;
  dim3 grid((M + TM - 1) / TM, N / F);
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  dense_kernel<<<grid, THREADS, 0, s>>>(X, ldx, W, bias, R, ldr, Y, ldy, M, K, N, act, slope);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// query-conditioned count head (lightning_model.py:127-131, 191-192, 212-221), factorised:
//   pred[g,q] = w2 . leaky_0.01( T[g] + Bq[q] ) + b2,  T = t . W1[:, :64]^T,  Bq = q . W1[:, 64:]^T + b1
// ------------------------------------------------------------------------------------------------------------------
constexpr int HEAD_H = 4 * F;   // 256
constexpr int HEAD_TG = 16;     // neighborhoods per CTA

__global__ void __launch_bounds__(THREADS) count_head_kernel(const float* __restrict__ T, const float* __restrict__ Bq,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             int G, int Q, float* __restrict__ pred,
                                                             float* __restrict__ count) {
  extern __shared__ __align__(16) float sm[];
  float* sT = sm;                          // [HEAD_TG][HEAD_H + 1]
  float* sW2 = sT + HEAD_TG * (HEAD_H + 1);  // [HEAD_H]
  float* sB = sW2 + HEAD_H;                // [Q][HEAD_H + 1]
  const int g0 = blockIdx.x * HEAD_TG;
  for (int i = threadIdx.x; i < HEAD_TG * HEAD_H; i += THREADS) {
    const int r = i / HEAD_H, c = i % HEAD_H;
    sT[r * (HEAD_H + 1) + c] = (g0 + r < G) ? T[(size_t)(g0 + r) * HEAD_H + c] : 0.f;
  }
  for (int i = threadIdx.x; i < HEAD_H; i += THREADS) sW2[i] = w2[i];
  for (int i = threadIdx.x; i < Q * HEAD_H; i += THREADS) sB[(i / HEAD_H) * (HEAD_H + 1) + i % HEAD_H] = Bq[i];
  __syncthreads();
  const float bias2 = b2[0];
  for (int i = threadIdx.x; i < HEAD_TG * Q; i += THREADS) {
    const int r = i / Q, q = i % Q;
    if (g0 + r >= G) continue;
    const float* t = sT + r * (HEAD_H + 1);
    const float* bq = sB + q * (HEAD_H + 1);
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < HEAD_H; ++j) {
      float v = t[j] + bq[j];
      v = v > 0.f ? v : 0.01f * v;  // nn.LeakyReLU() default slope
      acc = fmaf(v, sW2[j], acc);
    }
    acc += bias2;
    if (pred) pred[(size_t)(g0 + r) * Q + q] = acc;
    if (count) count[(size_t)(g0 + r) * Q + q] = exp2f(acc) - 1.f;  // 2**pred - 1 (lightning_model.py:221)
  }
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Workspace {
  int32_t* row_nbh; int32_t* crow; uint8_t* canon_code; int32_t* quirk_row;
  float *hA, *hB, *emb_a, *pool, *cvec, *z, *t1, *t2, *t3;
  void* fused;  // tile plan of the fused tcgen05 path
  void* emb_img;  // canonical rows of all layers as A operand images of the anchor GEMM (fused path)
  void* mt;     // pooling partials of the multi-tile tcgen05 path
  size_t bytes;
};

Workspace carve(void* base, int V, int G, int layers) {
  Workspace w;
  const int emb_ld = (layers + 1) * F;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += align_up(n); return (char*)base + o; };
  w.row_nbh = (int32_t*)take((size_t)V * 4);
  w.crow = (int32_t*)take((size_t)V * 4);
  w.canon_code = (uint8_t*)take((size_t)V);
  w.quirk_row = (int32_t*)take((size_t)G * 4);
  w.hA = (float*)take((size_t)V * F * 4);
  w.hB = (float*)take((size_t)V * F * 4);
  w.emb_a = (float*)take((size_t)G * emb_ld * 4);
  w.pool = (float*)take((size_t)G * emb_ld * 4);
  w.cvec = (float*)take((size_t)G * 2 * F * 4);
  w.z = (float*)take((size_t)G * emb_ld * 4);
  w.t1 = (float*)take((size_t)G * F * 4);
  w.t2 = (float*)take((size_t)G * F * 4);
  w.t3 = (float*)take((size_t)G * 4 * F * 4);
  w.fused = (void*)take((size_t)desco_internal_shmp_fused_workspace_bytes(G));
  w.emb_img = (void*)take((size_t)((G + 127) / 128) * (layers + 1) * 3 * 128 * 128);
  w.mt = (void*)take((size_t)desco_internal_shmp_mt_workspace_bytes(V, G));
  w.bytes = off;
  return w;
}

}  // namespace

void desco_internal_shmp_cvec(const float* emb_a, int emb_ld, int layer, const float* Cw, int G, float* cvec, cudaStream_t s) {
  shmp_cvec_kernel<<<(G + 7) / 8, 256, 0, s>>>(emb_a, emb_ld, layer, Cw, G, cvec);
}
void desco_internal_shmp_copy_last_rows(const int32_t* nbh_ptr, int G, const float* h, int layer, float* emb_a, int emb_ld,
                                        cudaStream_t s) {
  desco_count_launches(1);
  shmp_copy_last_rows_kernel<<<(G * 32 + 255) / 256, 256, 0, s>>>(nbh_ptr, G, h, layer, emb_a, emb_ld);
}

extern "C" {

int64_t desco_shmp_workspace_bytes(int32_t num_rows, int32_t num_neighborhoods, int32_t layers) {
  return (int64_t)carve(nullptr, num_rows, num_neighborhoods, layers).bytes;
}

int64_t desco_shmp_layer_weight_floats(void) { return (int64_t)2 * (KC * F + F) + (int64_t)F * 2 * F; }
int64_t desco_shmp_tc_layer_bytes(void) { return (int64_t)SHMP_TC_LAYER_BYTES; }
int64_t desco_shmp_mt_layer_bytes(void) { return (int64_t)SHMP_MT_LAYER_BYTES; }

static int shmp_forward_impl(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                             int32_t num_neighborhoods, int32_t num_rows, int32_t hetero, int32_t pyg_batch_size,
                             const float* feat, int32_t input_dim, const float* w_pre, const float* w_layers,
                             const void* w_layers_tc, const void* w_layers_mt, const float* w_readout,
                             const void* w_readout_tc, int32_t layers, int32_t hidden, float* out_emb, void* workspace,
                             int64_t workspace_bytes, int32_t precision, int32_t* status, const int32_t* sizes_dev,
                             void* stream) {
  // sizes_dev != NULL: the stream-ordered form - num_neighborhoods / num_rows are CAPACITIES and the batch's own
  // {G, V, E, max rows} are read on the device (fused tensor-core path only)
  const int G = num_neighborhoods, V = num_rows;
  if (hidden != F || layers < 1 || input_dim < 1 || G < 0 || V < 0 || hetero < 0 || hetero > 2) return DESCO_EINVAL;
  if (precision < DESCO_PRECISION_FP32 || precision > DESCO_PRECISION_BF16) return DESCO_EINVAL;
  if (G == 0) return DESCO_OK;
  if (!nbh_ptr || !edge_ptr || !edge_col || !edge_tri || !w_pre || !w_readout || !out_emb || !workspace) return DESCO_EINVAL;
  // hetero = 2: one node type, but the LAST row of every neighborhood (the centre, marked by node_feature = 1) goes through
  // anchor_mlp before the pooling - the homogeneous model of hetero_graph = False (gnn_model.py:74-83, workload.py:238-241)
  const bool anchored = hetero == 2;
  if (anchored) hetero = 0;
  const bool mt = w_layers_mt != nullptr && precision != DESCO_PRECISION_FP32;  // multi-tile tcgen05 path (any size)
  const bool fused = !mt && precision != DESCO_PRECISION_FP32;
  if (fused && (!hetero || !w_layers_tc || !w_readout_tc || !status)) return DESCO_EINVAL;  // tensor-core path: count/canonical batches
  if (mt && (!w_layers || !status || ((hetero || anchored) && !w_readout_tc))) return DESCO_EINVAL;
  if (sizes_dev && !fused) return DESCO_EINVAL;
  if (!fused && !w_layers) return DESCO_EINVAL;
  Workspace ws = carve(workspace, V, G, layers);
  if ((int64_t)ws.bytes > workspace_bytes) return DESCO_ENOMEM;
  cudaStream_t s = (cudaStream_t)stream;
  const int emb_ld = (layers + 1) * F;
  const int Vc = hetero ? V - G : V;

  if (fused) {
    const int rc = desco_internal_shmp_fused_layers(nbh_ptr, edge_ptr, edge_col, edge_tri, G, pyg_batch_size, feat, input_dim,
                                                    w_pre, w_layers_tc, layers, precision == DESCO_PRECISION_BF16X3 ? 3 : 1,
                                                    ws.emb_a, ws.emb_img, ws.pool, emb_ld, ws.fused, status, sizes_dev, s);
    if (rc) return rc;
  } else {
  {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s, 2);
    shmp_plan_kernel<<<(G + 255) / 256, 256, 0, s>>>(nbh_ptr, G, hetero, pyg_batch_size, ws.quirk_row);
    shmp_plan_rows_kernel<<<(V + 255) / 256, 256, 0, s>>>(nbh_ptr, edge_ptr, edge_col, edge_tri, G, V, hetero, ws.quirk_row,
                                                         ws.row_nbh, ws.crow, ws.canon_code);
    DESCO_LAUNCH_CHECK();
  }
  {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
    const long long n = (long long)V * F;
    shmp_pre_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(nbh_ptr, ws.row_nbh, V, hetero, feat, input_dim, w_pre,
                                                               ws.hA, ws.emb_a, emb_ld);
    DESCO_LAUNCH_CHECK();
  }
  if (mt) {
    const int rc = desco_internal_shmp_mt_layers(nbh_ptr, edge_ptr, edge_col, edge_tri, G, V, hetero, ws.row_nbh, ws.crow,
                                                 ws.canon_code, ws.quirk_row, ws.hA, ws.hB, ws.emb_a, ws.pool, ws.cvec, emb_ld,
                                                 w_layers, desco_shmp_layer_weight_floats(), w_layers_mt, layers,
                                                 precision == DESCO_PRECISION_BF16X3 ? 3 : 1, ws.mt, status, anchored ? 1 : 0, s);
    if (rc) return rc;
  } else {
  const size_t smem = (size_t)(TM * LDA + KC * F) * sizeof(float) + 3 * TM * sizeof(int);
  static bool attr_set = false;
  if (!attr_set) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(shmp_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int n_tiles = (Vc + TM - 1) / TM + (G + TM - 1) / TM;
  const int grid = n_tiles < 2 * desco_num_sms() ? n_tiles : 2 * desco_num_sms();
  const int64_t lw = desco_shmp_layer_weight_floats();
  float* h_in = ws.hA;
  float* h_out = ws.hB;
  for (int l = 0; l < layers; ++l) {
    const float* wl = w_layers + (size_t)l * lw;  // [Wc | bias_c | Cw | Wa | bias_a]
    LayerArgs a;
    a.nbh_ptr = nbh_ptr; a.edge_ptr = edge_ptr; a.edge_col = edge_col; a.edge_tri = edge_tri;
    a.row_nbh = ws.row_nbh; a.crow = ws.crow; a.canon_code = ws.canon_code; a.quirk_row = ws.quirk_row;
    a.G = G; a.Vc = Vc; a.hetero = hetero; a.layer = l; a.emb_ld = emb_ld;
    a.h_in = h_in; a.h_out = h_out; a.emb_a = ws.emb_a; a.pool = ws.pool; a.cvec = ws.cvec;
    a.Wc = wl; a.bias_c = wl + KC * F; a.Wa = wl + KC * F + F + F * 2 * F; a.bias_a = a.Wa + KC * F;
    if (hetero) {
      DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
      shmp_cvec_kernel<<<(G + 7) / 8, 256, 0, s>>>(ws.emb_a, emb_ld, l, wl + KC * F + F, G, ws.cvec);
      DESCO_LAUNCH_CHECK();
    }
    if (anchored) desco_internal_shmp_copy_last_rows(nbh_ptr, G, h_in, l, ws.emb_a, emb_ld, s);
    {
      DescoProfScope prof(DESCO_PROF_SHMP_LAYER, s);
      shmp_layer_kernel<<<grid, THREADS, smem, s>>>(a);
      DESCO_LAUNCH_CHECK();
    }
    float* t = h_in; h_in = h_out; h_out = t;
  }
  {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
    shmp_pool_last_kernel<<<(G * 32 + 255) / 256, 256, 0, s>>>(nbh_ptr, G, hetero, h_in, layers, ws.pool, emb_ld);
    DESCO_LAUNCH_CHECK();
  }
  if (anchored) desco_internal_shmp_copy_last_rows(nbh_ptr, G, h_in, layers, ws.emb_a, emb_ld, s);
  }  // layer-by-layer FFMA kernels
  }  // layered paths
  if (anchored) {  // the pool above covers every row: take the centre rows out, anchor_mlp puts them back below
    const long long n = (long long)G * emb_ld;
    desco_count_launches(1);
    shmp_sub_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ws.pool, ws.emb_a, n);
    DESCO_LAUNCH_CHECK();
  }

  // readout: [Wanc (emb_ld x emb_ld) | banc | P0 (emb_ld x F) | b0 | P1 (F x F) | b1 | P2 (F x 4F) | b2 | P3 (4F x F) | b3]
  const float* r = w_readout;
  const float* Wanc = r; r += (size_t)emb_ld * emb_ld;
  const float* banc = r; r += emb_ld;
  const float* P0 = r; r += (size_t)emb_ld * F;
  const float* b0 = r; r += F;
  const float* P1 = r; r += F * F;
  const float* b1 = r; r += F;
  const float* P2 = r; r += F * 4 * F;
  const float* b2 = r; r += 4 * F;
  const float* P3 = r; r += 4 * F * F;
  const float* b3 = r;
  int rc;
  const float* z = ws.pool;
  if (fused || (mt && (hetero || anchored))) {  // the same chain on the tensor pipe (csrc/dense_tc.cu); images in w_readout_tc, biases from w_readout
    const int passes = precision == DESCO_PRECISION_BF16X3 ? 6 : 1;  // the readout sums cancel heavily: full 3-way split
    // bf16 hi + mid + lo = 6 bytes per weight; the blob starts with the anchor_mlp images (the post_mp images that follow
    // are for a dense_tc post_mp chain; the default chain below is the single-launch fp32 kernel)
    const uint8_t* iWanc = (const uint8_t*)w_readout_tc;
    // 144-column blocks: 576 / 144 = 4 column blocks x 32 row blocks = 128 CTAs for 4096 neighborhoods, one wave on 148 SMs
    // (96-column blocks were 192 CTAs = two waves)
    if (emb_ld % 144) return DESCO_EINVAL;
    // (the fused kernel leaves the canonical rows as ready-made A operand images; the multi-tile path as fp32 rows)
    if ((rc = desco_internal_dense_tc(ws.emb_a, fused ? ws.emb_img : nullptr, emb_ld, iWanc, banc, ws.pool, emb_ld, ws.z, emb_ld,
                                      G, emb_ld, emb_ld, 144, 2, 0.1f, passes, status, sizes_dev, s))) return rc;
    return desco_internal_readout_chain(ws.z, emb_ld, emb_ld, G, P0, b0, P1, b1, P2, b2, P3, b3, out_emb, sizes_dev, s);
  }
  if (hetero || anchored) {  // z = pool_count + LeakyReLU_0.1(anchor(emb_canonical))  (gnn_model.py:69-73, 88-89, 107)
    rc = dense(ws.emb_a, emb_ld, Wanc, banc, ws.pool, emb_ld, ws.z, emb_ld, G, emb_ld, emb_ld, ACT_LEAKY, 0.1f, s);
    if (rc) return rc;
    z = ws.z;
  }
  // post_mp (gnn_model.py:44-53)
  return desco_internal_readout_chain(z, emb_ld, emb_ld, G, P0, b0, P1, b1, P2, b2, P3, b3, out_emb, nullptr, s);
}

int desco_shmp_forward(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                       int32_t num_neighborhoods, int32_t num_rows, int32_t hetero, int32_t pyg_batch_size,
                       const float* feat, int32_t input_dim, const float* w_pre, const float* w_layers,
                       const void* w_layers_tc, const float* w_readout, const void* w_readout_tc, int32_t layers,
                       int32_t hidden, float* out_emb, void* workspace, int64_t workspace_bytes, int32_t precision,
                       int32_t* status, void* stream) {
  return shmp_forward_impl(nbh_ptr, edge_ptr, edge_col, edge_tri, num_neighborhoods, num_rows, hetero, pyg_batch_size, feat,
                           input_dim, w_pre, w_layers, w_layers_tc, nullptr, w_readout, w_readout_tc, layers, hidden, out_emb,
                           workspace, workspace_bytes, precision, status, nullptr, stream);
}

int desco_shmp_forward_dev(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                           int32_t cap_neighborhoods, int32_t cap_rows, const int32_t* sizes_dev, int32_t pyg_batch_size,
                           const float* feat, int32_t input_dim, const float* w_pre, const void* w_layers_tc,
                           const float* w_readout, const void* w_readout_tc, int32_t layers, int32_t hidden, float* out_emb,
                           void* workspace, int64_t workspace_bytes, int32_t precision, int32_t* status, void* stream) {
  if (!sizes_dev || precision == DESCO_PRECISION_FP32) return DESCO_EINVAL;
  return shmp_forward_impl(nbh_ptr, edge_ptr, edge_col, edge_tri, cap_neighborhoods, cap_rows, 1, pyg_batch_size, feat,
                           input_dim, w_pre, nullptr, w_layers_tc, nullptr, w_readout, w_readout_tc, layers, hidden, out_emb,
                           workspace, workspace_bytes, precision, status, sizes_dev, stream);
}

int desco_shmp_forward_mt(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col, const uint8_t* edge_tri,
                          int32_t num_neighborhoods, int32_t num_rows, int32_t hetero, int32_t pyg_batch_size,
                          const float* feat, int32_t input_dim, const float* w_pre, const float* w_layers,
                          const void* w_layers_mt, const float* w_readout, const void* w_readout_tc, int32_t layers,
                          int32_t hidden, float* out_emb, void* workspace, int64_t workspace_bytes, int32_t precision,
                          int32_t* status, void* stream) {
  if (!w_layers_mt || precision == DESCO_PRECISION_FP32) return DESCO_EINVAL;
  return shmp_forward_impl(nbh_ptr, edge_ptr, edge_col, edge_tri, num_neighborhoods, num_rows, hetero, pyg_batch_size, feat,
                           input_dim, w_pre, w_layers, nullptr, w_layers_mt, w_readout, w_readout_tc, layers, hidden, out_emb,
                           workspace, workspace_bytes, precision, status, nullptr, stream);
}

int64_t desco_count_head_workspace_bytes(int32_t num_neighborhoods, int32_t num_queries) {
  return (int64_t)(align_up((size_t)num_neighborhoods * HEAD_H * 4) + align_up((size_t)num_queries * HEAD_H * 4));
}

static int count_head_impl(const float* emb_target, int32_t num_neighborhoods, const float* emb_query, int32_t num_queries,
                           const float* w_head, const void* w_head_tc, int32_t hidden, float* out_pred, float* out_count,
                           void* workspace, int64_t workspace_bytes, int32_t precision, int32_t* status,
                           const int32_t* g_dev, void* stream) {
  const int G = num_neighborhoods, Q = num_queries;
  if (g_dev && Q > 32) return DESCO_ERANGE;  // the stream-ordered form serves the one-launch head
  if (hidden != F || G < 0 || Q < 0) return DESCO_EINVAL;
  if (G == 0 || Q == 0) return DESCO_OK;
  if (!emb_target || !emb_query || !w_head || !workspace || (!out_pred && !out_count)) return DESCO_EINVAL;
  if (workspace_bytes < desco_count_head_workspace_bytes(G, Q)) return DESCO_ENOMEM;
  cudaStream_t s = (cudaStream_t)stream;
  float* T = (float*)workspace;
  float* Bq = (float*)((char*)workspace + align_up((size_t)G * HEAD_H * 4));
  // w_head: [W1a (F x 4F) | W1b (F x 4F) | b1 (4F) | w2 (4F) | b2 (1)]
  const float* W1a = w_head;
  const float* W1b = W1a + F * HEAD_H;
  const float* b1 = W1b + F * HEAD_H;
  const float* w2 = b1 + HEAD_H;
  const float* b2 = w2 + HEAD_H;
  if (Q <= 32)  // one fused launch (csrc/readout.cu); the multi-launch path below serves larger query sets
    return desco_internal_count_head_fused(emb_target, G, emb_query, Q, W1a, W1b, b1, w2, b2, out_pred, out_count, Bq, g_dev, s);
  int rc;
  if (precision != DESCO_PRECISION_FP32) {
    if (!w_head_tc || !status) return DESCO_EINVAL;
    const int passes = precision == DESCO_PRECISION_BF16X3 ? 6 : 1;
    const uint8_t* iW1a = (const uint8_t*)w_head_tc;
    const uint8_t* iW1b = iW1a + (size_t)F * HEAD_H * 6;
    if ((rc = desco_internal_dense_tc(emb_target, nullptr, F, iW1a, nullptr, nullptr, 0, T, HEAD_H, G, F, HEAD_H, 128, 0, 0.f, passes, status, nullptr, s))) return rc;
    if ((rc = desco_internal_dense_tc(emb_query, nullptr, F, iW1b, b1, nullptr, 0, Bq, HEAD_H, Q, F, HEAD_H, 128, 0, 0.f, passes, status, nullptr, s))) return rc;
  } else {
  if ((rc = dense(emb_target, F, W1a, nullptr, nullptr, 0, T, HEAD_H, G, F, HEAD_H, ACT_NONE, 0.f, s))) return rc;
  if ((rc = dense(emb_query, F, W1b, b1, nullptr, 0, Bq, HEAD_H, Q, F, HEAD_H, ACT_NONE, 0.f, s))) return rc;
  }
  const size_t smem = (size_t)(HEAD_TG * (HEAD_H + 1) + HEAD_H + Q * (HEAD_H + 1)) * sizeof(float);
  if (smem > 200 * 1024) return DESCO_ERANGE;
  DESCO_CUDA_TRY(cudaFuncSetAttribute(count_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  count_head_kernel<<<(G + HEAD_TG - 1) / HEAD_TG, THREADS, smem, s>>>(T, Bq, w2, b2, G, Q, out_pred, out_count);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_count_head(const float* emb_target, int32_t num_neighborhoods, const float* emb_query, int32_t num_queries,
                     const float* w_head, const void* w_head_tc, int32_t hidden, float* out_pred, float* out_count,
                     void* workspace, int64_t workspace_bytes, int32_t precision, int32_t* status, void* stream) {
  return count_head_impl(emb_target, num_neighborhoods, emb_query, num_queries, w_head, w_head_tc, hidden, out_pred, out_count,
                         workspace, workspace_bytes, precision, status, nullptr, stream);
}

int desco_count_head_dev(const float* emb_target, int32_t cap_neighborhoods, const int32_t* sizes_dev, const float* emb_query,
                         int32_t num_queries, const float* w_head, int32_t hidden, float* out_pred, float* out_count,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  if (!sizes_dev) return DESCO_EINVAL;
  return count_head_impl(emb_target, cap_neighborhoods, emb_query, num_queries, w_head, nullptr, hidden, out_pred, out_count,
                         workspace, workspace_bytes, DESCO_PRECISION_FP32, nullptr, sizes_dev, stream);
}

}  // extern "C"
