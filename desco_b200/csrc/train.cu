// Training step of the SHMP neighborhood-counting model (forward with saved activations, backward, Adam), sm_100a, fp32.
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/lightning_model.py:228-254  train_forward   (per-query smooth-L1 on log2(truth + 1), mean)
//   subgraph_counting/lightning_model.py:285-289  criterion       F.smooth_l1_loss
//   subgraph_counting/lightning_model.py:160-173  configure_optimizers  torch.optim.Adam
//   and the autograd graph PyTorch records through gnn_model.py:58-109, 230-277, 362-404 (SAGEConv gather/scatter,
//   Linear, ReLU/LeakyReLU, skip-concat, global_add_pool).
//
// Unlike the inference kernels (shmp_fused.cu folds U_m . W_rel on the host), training works on the ORIGINAL
// parameters in their torch layout ([out][in] row-major) so every gradient lands directly in the parameter's .grad:
//   * node features are stored per node type in compact matrices (count rows [Vc, .], canonical rows [G, .]); the
//     skip-concat IS the storage (emb[:, 64 l : 64 l + 64] = h^l), nothing is copied;
//   * train_aggregate_kernel: segmented, edge-type- and source-type-split neighbour sums (the four incoming relations
//     of a count row, the two of a canonical row); its transpose is the same gather with the roles swapped, because the
//     edge set and the SHMP types are symmetric - no atomics, deterministic;
//   * train_dense_kernel: Y (+)= act(sum_b X_b . W_b^T + bias) over a list of 64-wide K blocks, each with its own
//     source matrix and weight (relation weights are never concatenated); the same kernel with K-major weights is the
//     data gradient;
//   * train_wgrad_kernel: dW += dY^T X (split over row chunks, fp32 atomics into .grad), bias gradient fused;
//   * head kernels: query-conditioned count head + smooth-L1 loss forward and backward for all (neighborhood, query)
//     pairs at once; train_adam_kernel: one launch over the flat parameter buffer.
#include <math.h>

#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;
constexpr int TM = 64;
constexpr int THREADS = 256;
constexpr int MAXB = 12;   // K blocks of 64 per dense / wgrad call (anchor_mlp: 576 / 64 = 9)
constexpr int MAXBIAS = 4;

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

// ------------------------------------------------------------------------------------------------------------------
// plan: packed row -> neighborhood, compact count row -> neighborhood, the remove_self_loops quirk row (shmp.cu)
// ------------------------------------------------------------------------------------------------------------------
__global__ void train_plan_kernel(const int32_t* __restrict__ nbh_ptr, int G, int hetero, int pyg_batch_size,
                                  int32_t* __restrict__ row_nbh, int32_t* __restrict__ crow_nbh,
                                  int32_t* __restrict__ quirk_row) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const int lane = lane_id();
  const int lo = nbh_ptr[g], hi = nbh_ptr[g + 1];
  if (lane == 0) {
    int quirk = -1;
    if (hetero) {  // gnn_model.py:389-390 applied to the bipartite relations: see shmp.cu shmp_plan_kernel
      const int bs = pyg_batch_size > 0 ? pyg_batch_size : G;
      const int g0 = (g / bs) * bs;
      if (pyg_batch_size >= 0 && lo - nbh_ptr[g0] == 2 * (g - g0)) quirk = lo;  // pyg_batch_size < 0: quirk off
    }
    quirk_row[g] = quirk;
  }
  for (int r = lo + lane; r < hi; r += 32) {
    row_nbh[r] = g;
    if (!hetero) crow_nbh[r] = g;
    else if (r < hi - 1) crow_nbh[r - g] = g;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// aggregation and its transpose
// ------------------------------------------------------------------------------------------------------------------
struct AggArgs {
  const int32_t* nbh_ptr; const int32_t* edge_ptr; const int32_t* edge_col; const uint8_t* edge_tri;
  const int32_t* row_nbh; const int32_t* quirk_row;
  int V, hetero;
  float* Xc; int ldc;   // count rows (all rows when !hetero): input of the forward, += output of the transpose
  float* Xa; int lda;   // canonical rows
  float* Ac; int ldAc;  // [.., S*64]: slot = 2 * (source is canonical) + (tride ? 1 : 0); S = 4 hetero, 2 otherwise
  float* Aa; int ldAa;  // [G, 2*64]
};

template <bool TRANSPOSE>
__global__ void __launch_bounds__(256) train_aggregate_kernel(const AggArgs p) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= p.V) return;
  const int lane = lane_id();
  const int g = p.row_nbh[r];
  const int canon = p.hetero ? p.nbh_ptr[g + 1] - 1 : -1;
  const bool r_canon = r == canon;
  const int quirk = p.hetero ? p.quirk_row[g] : -1;
  float2 acc[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) acc[s] = make_float2(0.f, 0.f);
  const int eb = p.edge_ptr[r], ee = p.edge_ptr[r + 1];
  for (int base = eb; base < ee; base += 32) {
    const int e = base + lane;
    const int my_j = (e < ee) ? p.edge_col[e] : -1;
    const int my_t = (e < ee) ? (int)p.edge_tri[e] : 0;
    const int n = min(32, ee - base);
    for (int k = 0; k < n; ++k) {
      const int j = __shfl_sync(FULL_MASK, my_j, k);
      const int t = __shfl_sync(FULL_MASK, my_t, k);
      const bool j_canon = j == canon;
      if ((r_canon && j == quirk) || (j_canon && r == quirk)) continue;  // dropped in both directions
      const int tr = t ? 0 : 1;
      if (!TRANSPOSE) {
        const float* src = j_canon ? p.Xa + (size_t)g * p.lda : p.Xc + (size_t)(p.hetero ? j - g : j) * p.ldc;
        const float2 v = *reinterpret_cast<const float2*>(src + 2 * lane);
        const int s = (j_canon ? 2 : 0) + tr;
        // (static indexing keeps acc[] in registers)
        if (s == 0) { acc[0].x += v.x; acc[0].y += v.y; }
        else if (s == 1) { acc[1].x += v.x; acc[1].y += v.y; }
        else if (s == 2) { acc[2].x += v.x; acc[2].y += v.y; }
        else { acc[3].x += v.x; acc[3].y += v.y; }
      } else {
        // d x_r += d A[j][slot of the relation (type(r) -> type(j), tri)]
        const float* src = j_canon ? p.Aa + (size_t)g * p.ldAa + tr * F
                                   : p.Ac + (size_t)(p.hetero ? j - g : j) * p.ldAc + ((r_canon ? 2 : 0) + tr) * F;
        const float2 v = *reinterpret_cast<const float2*>(src + 2 * lane);
        acc[0].x += v.x; acc[0].y += v.y;
      }
    }
  }
  if (!TRANSPOSE) {
    if (r_canon) {
      float* dst = p.Aa + (size_t)g * p.ldAa;
      *reinterpret_cast<float2*>(dst + 2 * lane) = acc[0];
      *reinterpret_cast<float2*>(dst + F + 2 * lane) = acc[1];
    } else {
      float* dst = p.Ac + (size_t)(p.hetero ? r - g : r) * p.ldAc;
      *reinterpret_cast<float2*>(dst + 2 * lane) = acc[0];
      *reinterpret_cast<float2*>(dst + F + 2 * lane) = acc[1];
      if (p.hetero) {
        *reinterpret_cast<float2*>(dst + 2 * F + 2 * lane) = acc[2];
        *reinterpret_cast<float2*>(dst + 3 * F + 2 * lane) = acc[3];
      }
    }
  } else {
    float* dst = r_canon ? p.Xa + (size_t)g * p.lda : p.Xc + (size_t)(p.hetero ? r - g : r) * p.ldc;
    float2 o = *reinterpret_cast<float2*>(dst + 2 * lane);
    o.x += acc[0].x; o.y += acc[0].y;
    *reinterpret_cast<float2*>(dst + 2 * lane) = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// dense: Y (+)= act( sum_b X_b[M,64] . W_b^T + sum bias )
// ------------------------------------------------------------------------------------------------------------------
struct DenseArgs {
  const float* X[MAXB]; int ldx[MAXB];
  const float* W[MAXB]; int ldw[MAXB];
  const float* bias[MAXBIAS];
  int nblk, nbias, w_nmajor;   // w_nmajor 1: W_b[n * ldw + k] (torch Linear weight);  0: W_b[k * ldw + n] (data gradient)
  float* Y; int ldy, M, N, act, accumulate;
  float slope;
};

__global__ void __launch_bounds__(THREADS) train_dense_kernel(const DenseArgs p) {
  __shared__ __align__(16) float sX[TM][F + 4];
  __shared__ __align__(16) float sW[F][F];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * F;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int b = 0; b < p.nblk; ++b) {
    __syncthreads();
    const float* X = p.X[b];
    const int ldx = p.ldx[b];
    for (int i = tid; i < TM * F / 4; i += THREADS) {
      const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < p.M) v = *reinterpret_cast<const float4*>(X + (size_t)(m0 + r) * ldx + c4);
      *reinterpret_cast<float4*>(&sX[r][c4]) = v;
    }
    const float* W = p.W[b];
    const int ldw = p.ldw[b];
    if (p.w_nmajor) {
      for (int i = tid; i < F * F / 4; i += THREADS) {
        const int n = i % F, k4 = (i / F) * 4;
        const float4 v = *reinterpret_cast<const float4*>(W + (size_t)(n0 + n) * ldw + k4);
        sW[k4][n] = v.x; sW[k4 + 1][n] = v.y; sW[k4 + 2][n] = v.z; sW[k4 + 3][n] = v.w;
      }
    } else {
      for (int i = tid; i < F * F / 4; i += THREADS) {
        const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
        *reinterpret_cast<float4*>(&sW[r][c4]) = *reinterpret_cast<const float4*>(W + (size_t)r * ldw + n0 + c4);
      }
    }
    __syncthreads();
    tile_gemm<F, F + 4>(&sX[0][0], &sW[0][0], ty, tx, acc);
  }
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < p.nbias; ++i) {
    const float4 b = *reinterpret_cast<const float4*>(p.bias[i] + n0 + tx * 4);
    b4[0] += b.x; b4[1] += b.y; b4[2] += b.z; b4[3] += b.w;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float* y = p.Y + (size_t)m * p.ldy + n0 + tx * 4;
    float o[4] = {acc[i][0] + b4[0], acc[i][1] + b4[1], acc[i][2] + b4[2], acc[i][3] + b4[3]};
    if (p.accumulate) {
      const float4 old = *reinterpret_cast<const float4*>(y);
      o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p.act == ACT_RELU) o[j] = fmaxf(o[j], 0.f);
      else if (p.act == ACT_LEAKY) o[j] = o[j] > 0.f ? o[j] : o[j] * p.slope;
    }
    *reinterpret_cast<float4*>(y) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// weight gradient: dW_b[n][k] += sum_m dY[m][n] X[m][64 b + k];   db[n] += sum_m dY[m][n]
// ------------------------------------------------------------------------------------------------------------------
struct WgradArgs {
  const float* X; int ldx;
  const float* dY; int ldy;
  float* dW[MAXB]; int ldw[MAXB];
  float* db[MAXBIAS];
  int nblk, ndb, M, N, chunk;
};

__global__ void __launch_bounds__(THREADS) train_wgrad_kernel(const WgradArgs p) {
  __shared__ __align__(16) float sD[TM][F + 4];  // dY rows m, columns n
  __shared__ __align__(16) float sX[TM][F];      // X rows m, columns k
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int kb = blockIdx.x, n0 = blockIdx.y * F;
  const int mb = blockIdx.z * p.chunk, me = min(p.M, mb + p.chunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float colsum = 0.f;
  const bool want_db = kb == 0 && p.ndb > 0;
  for (int m0 = mb; m0 < me; m0 += TM) {
    __syncthreads();
    for (int i = tid; i < TM * F / 4; i += THREADS) {
      const int r = i / (F / 4), c4 = (i % (F / 4)) * 4;
      float4 d = make_float4(0.f, 0.f, 0.f, 0.f), x = d;
      if (m0 + r < me) {
        d = *reinterpret_cast<const float4*>(p.dY + (size_t)(m0 + r) * p.ldy + n0 + c4);
        x = *reinterpret_cast<const float4*>(p.X + (size_t)(m0 + r) * p.ldx + kb * F + c4);
      }
      *reinterpret_cast<float4*>(&sD[r][c4]) = d;
      *reinterpret_cast<float4*>(&sX[r][c4]) = x;
    }
    __syncthreads();
#pragma unroll 4
    for (int m = 0; m < TM; ++m) {
      const float4 d = *reinterpret_cast<const float4*>(&sD[m][ty * 4]);
      const float4 x = *reinterpret_cast<const float4*>(&sX[m][tx * 4]);
      const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(dv[i], x.x, acc[i][0]);
        acc[i][1] = fmaf(dv[i], x.y, acc[i][1]);
        acc[i][2] = fmaf(dv[i], x.z, acc[i][2]);
        acc[i][3] = fmaf(dv[i], x.w, acc[i][3]);
      }
    }
    if (want_db && tid < F) {
      float s = 0.f;
      for (int m = 0; m < TM; ++m) s += sD[m][tid];
      colsum += s;
    }
  }
  float* dW = p.dW[kb];
  const int ldw = p.ldw[kb];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) atomicAdd(dW + (size_t)(n0 + ty * 4 + i) * ldw + tx * 4 + j, acc[i][j]);
  if (want_db && tid < F)
    for (int i = 0; i < p.ndb; ++i) atomicAdd(p.db[i] + n0 + tid, colsum);
}

// dX[m][c] *= act'(fwd[m][c])   (derivative taken from the activation OUTPUT: slopes are positive, signs agree)
__global__ void train_act_backward_kernel(float* __restrict__ dX, int ldd, const float* __restrict__ fwd, int ldf, int M,
                                          int C, int act, float slope) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx % C);
  const float y = fwd[(size_t)m * ldf + c];
  if (y <= 0.f) dX[(size_t)m * ldd + c] *= (act == ACT_RELU) ? 0.f : slope;
}

// rows <- bias (pre_mp on ZeroNodeFeat inputs, gnn_model.py:231 with workload.py:431-440)
__global__ void train_fill_rows_kernel(float* __restrict__ Y, int ldy, int M, int C, const float* __restrict__ bias) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)M * C) return;
  const int m = (int)(idx / C), c = (int)(idx % C);
  Y[(size_t)m * ldy + c] = bias[c];
}

// db[c] += sum_m dY[m][c]
__global__ void train_colsum_kernel(const float* __restrict__ dY, int ldy, int M, int C, float* __restrict__ db) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int chunk = (M + gridDim.y - 1) / gridDim.y;
  const int mb = blockIdx.y * chunk, me = min(M, mb + chunk);
  float s = 0.f;
  for (int m = mb; m < me; ++m) s += dY[(size_t)m * ldy + c];
  atomicAdd(db + c, s);
}

// global_add_pool over the count rows (+ the anchor-transformed canonical row), gnn_model.py:88-89,107; and its backward
__global__ void train_pool_kernel(const int32_t* __restrict__ nbh_ptr, int G, int hetero, int C,
                                  const float* __restrict__ emb_c, int ldc, const float* __restrict__ z_a, int lda,
                                  float* __restrict__ pooled, int ldp) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)G * C) return;
  const int g = (int)(idx / C), c = (int)(idx % C);
  const int lo = nbh_ptr[g] - (hetero ? g : 0), hi = nbh_ptr[g + 1] - (hetero ? g + 1 : 0);
  float s = 0.f;
  for (int r = lo; r < hi; ++r) s += emb_c[(size_t)r * ldc + c];
  if (z_a) s += z_a[(size_t)g * lda + c];
  pooled[(size_t)g * ldp + c] = s;
}

__global__ void train_pool_backward_kernel(const int32_t* __restrict__ crow_nbh, int Vc, int C,
                                           const float* __restrict__ dpooled, int ldp, float* __restrict__ demb_c, int ldc) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)Vc * C) return;
  const int r = (int)(idx / C), c = (int)(idx % C);
  demb_c[(size_t)r * ldc + c] = dpooled[(size_t)crow_nbh[r] * ldp + c];
}

// ------------------------------------------------------------------------------------------------------------------
// count head + loss (lightning_model.py:127-131, 176-193, 228-254, 285-289)
//   pred[g,q] = w2 . leaky_0.01(T[g] + Bq[q]) + b2;  loss = mean_q mean_g smooth_l1(pred[g,q], log2(y[g,q] + 1))
// ------------------------------------------------------------------------------------------------------------------
constexpr int HEAD_H = 4 * F;
constexpr int HEAD_TG = 8;
constexpr int HEAD_MAXQ = 32;

__global__ void __launch_bounds__(THREADS) train_head_loss_kernel(const float* __restrict__ T, const float* __restrict__ Bq,
                                                                  const float* __restrict__ w2, const float* __restrict__ b2,
                                                                  const float* __restrict__ y, int G, int Q,
                                                                  float* __restrict__ pred, float* __restrict__ dpred,
                                                                  float* __restrict__ loss) {
  extern __shared__ __align__(16) float sm[];
  float* sT = sm;                               // [HEAD_TG][HEAD_H + 1]
  float* sW2 = sT + HEAD_TG * (HEAD_H + 1);     // [HEAD_H]
  float* sB = sW2 + HEAD_H;                     // [Q][HEAD_H + 1]
  __shared__ float s_loss;
  const int g0 = blockIdx.x * HEAD_TG;
  if (threadIdx.x == 0) s_loss = 0.f;
  for (int i = threadIdx.x; i < HEAD_TG * HEAD_H; i += THREADS) {
    const int r = i / HEAD_H, c = i % HEAD_H;
    sT[r * (HEAD_H + 1) + c] = (g0 + r < G) ? T[(size_t)(g0 + r) * HEAD_H + c] : 0.f;
  }
  for (int i = threadIdx.x; i < HEAD_H; i += THREADS) sW2[i] = w2[i];
  for (int i = threadIdx.x; i < Q * HEAD_H; i += THREADS) sB[(i / HEAD_H) * (HEAD_H + 1) + i % HEAD_H] = Bq[i];
  __syncthreads();
  const float bias2 = b2[0];
  const float inv = 1.f / ((float)G * (float)Q);
  float lsum = 0.f;
  for (int i = threadIdx.x; i < HEAD_TG * Q; i += THREADS) {
    const int r = i / Q, q = i % Q;
    if (g0 + r >= G) continue;
    const float* t = sT + r * (HEAD_H + 1);
    const float* bq = sB + q * (HEAD_H + 1);
    float acc = 0.f;
#pragma unroll 8
    for (int j = 0; j < HEAD_H; ++j) {
      float v = t[j] + bq[j];
      v = v > 0.f ? v : 0.01f * v;
      acc = fmaf(v, sW2[j], acc);
    }
    acc += bias2;
    const size_t o = (size_t)(g0 + r) * Q + q;
    if (pred) pred[o] = acc;
    if (y) {
      const float d = acc - log2f(y[o] + 1.f);            // :249
      const float ad = fabsf(d);
      lsum += ad < 1.f ? 0.5f * d * d : ad - 0.5f;        // smooth_l1, beta = 1
      dpred[o] = fminf(fmaxf(d, -1.f), 1.f) * inv;
    }
  }
  if (y) {
    lsum = warp_sum(lsum);
    if (lane_id() == 0) atomicAdd(&s_loss, lsum);
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(loss, s_loss * inv);
  }
}

// thread = hidden unit j; dT[g][j] = sum_q dpred[g,q] w2[j] leaky'(a);  dBq[q][j] = sum_g ...;  dw2[j] = sum dpred leaky(a)
__global__ void __launch_bounds__(HEAD_H) train_head_backward_kernel(const float* __restrict__ T, const float* __restrict__ Bq,
                                                                     const float* __restrict__ w2,
                                                                     const float* __restrict__ dpred, int G, int Q,
                                                                     float* __restrict__ dT, float* __restrict__ dBq,
                                                                     float* __restrict__ dw2, float* __restrict__ db2) {
  __shared__ float sD[HEAD_TG][HEAD_MAXQ];
  const int j = threadIdx.x;
  const int g0 = blockIdx.x * HEAD_TG;
  for (int i = threadIdx.x; i < HEAD_TG * HEAD_MAXQ; i += HEAD_H) {
    const int r = i / HEAD_MAXQ, q = i % HEAD_MAXQ;
    sD[r][q] = (g0 + r < G && q < Q) ? dpred[(size_t)(g0 + r) * Q + q] : 0.f;
  }
  __syncthreads();
  float bq[HEAD_MAXQ], dB[HEAD_MAXQ];
#pragma unroll
  for (int q = 0; q < HEAD_MAXQ; ++q) {
    bq[q] = q < Q ? Bq[(size_t)q * HEAD_H + j] : 0.f;
    dB[q] = 0.f;
  }
  const float w = w2[j];
  float dw = 0.f, dbias = 0.f;
  for (int r = 0; r < HEAD_TG && g0 + r < G; ++r) {
    const float t = T[(size_t)(g0 + r) * HEAD_H + j];
    float dt = 0.f;
#pragma unroll
    for (int q = 0; q < HEAD_MAXQ; ++q) {
      const float d = sD[r][q];
      const float a = t + bq[q];
      const float der = a > 0.f ? 1.f : 0.01f;
      const float gsig = d * w * der;
      dt += gsig;
      dB[q] += gsig;
      dw = fmaf(d, a * der, dw);   // leaky(a) = a * der
      dbias += d;
    }
    dT[(size_t)(g0 + r) * HEAD_H + j] = dt;
  }
#pragma unroll
  for (int q = 0; q < HEAD_MAXQ; ++q)
    if (q < Q) atomicAdd(dBq + (size_t)q * HEAD_H + j, dB[q]);
  atomicAdd(dw2 + j, dw);
  if (j == 0) atomicAdd(db2, dbias);
}

// torch.optim.Adam (no amsgrad), one element per thread over the flat parameter buffer
__global__ void train_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, float bc1, float bc2_sqrt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = g[i];
  const float pi = p[i];
  if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] = pi - (lr / bc1) * (mi / denom);
}

inline unsigned blocks_for(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

}  // namespace

extern "C" {

int desco_train_plan(const int32_t* nbh_ptr, int32_t num_neighborhoods, int32_t hetero, int32_t pyg_batch_size,
                     int32_t* row_nbh, int32_t* crow_nbh, int32_t* quirk_row, void* stream) {
  if (num_neighborhoods < 0) return DESCO_EINVAL;
  if (num_neighborhoods == 0) return DESCO_OK;
  if (!nbh_ptr || !row_nbh || !crow_nbh || !quirk_row) return DESCO_EINVAL;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, (cudaStream_t)stream);
  train_plan_kernel<<<blocks_for((long long)num_neighborhoods * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      nbh_ptr, num_neighborhoods, hetero, pyg_batch_size, row_nbh, crow_nbh, quirk_row);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_aggregate(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                          const uint8_t* edge_tri, const int32_t* row_nbh, const int32_t* quirk_row, int32_t num_rows,
                          int32_t hetero, int32_t transpose, float* xc, int32_t ldc, float* xa, int32_t lda, float* ac,
                          int32_t ld_ac, float* aa, int32_t ld_aa, void* stream) {
  if (num_rows < 0) return DESCO_EINVAL;
  if (num_rows == 0) return DESCO_OK;
  if (!nbh_ptr || !edge_ptr || !edge_col || !edge_tri || !row_nbh || !quirk_row || !xc || !ac) return DESCO_EINVAL;
  if (hetero && (!xa || !aa)) return DESCO_EINVAL;
  if ((ldc | lda | ld_ac | ld_aa) & 1) return DESCO_EINVAL;  // float2 accesses
  AggArgs a;
  a.nbh_ptr = nbh_ptr; a.edge_ptr = edge_ptr; a.edge_col = edge_col; a.edge_tri = edge_tri;
  a.row_nbh = row_nbh; a.quirk_row = quirk_row; a.V = num_rows; a.hetero = hetero;
  a.Xc = xc; a.ldc = ldc; a.Xa = xa; a.lda = lda; a.Ac = ac; a.ldAc = ld_ac; a.Aa = aa; a.ldAa = ld_aa;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_LAYER, s);
  const unsigned grid = blocks_for((long long)num_rows * 32, 256);
  if (transpose) train_aggregate_kernel<true><<<grid, 256, 0, s>>>(a);
  else train_aggregate_kernel<false><<<grid, 256, 0, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_dense(const float* const* x, const int32_t* ldx, const float* const* w, const int32_t* ldw,
                      int32_t num_blocks, int32_t w_nmajor, const float* const* bias, int32_t num_bias, float* y,
                      int32_t ldy, int32_t m, int32_t n, int32_t act, float slope, int32_t accumulate, void* stream) {
  if (num_blocks < 1 || num_blocks > MAXB || num_bias < 0 || num_bias > MAXBIAS || m < 0 || n <= 0 || n % F) return DESCO_EINVAL;
  if (act < ACT_NONE || act > ACT_LEAKY || (accumulate && act != ACT_NONE)) return DESCO_EINVAL;
  if (m == 0) return DESCO_OK;
  if (!x || !ldx || !w || !ldw || !y || (num_bias && !bias) || (ldy & 3)) return DESCO_EINVAL;
  DenseArgs a;
  for (int b = 0; b < num_blocks; ++b) {
    if (!x[b] || !w[b] || (ldx[b] & 3) || (ldw[b] & 3)) return DESCO_EINVAL;
    a.X[b] = x[b]; a.ldx[b] = ldx[b]; a.W[b] = w[b]; a.ldw[b] = ldw[b];
  }
  for (int i = 0; i < num_bias; ++i) {
    if (!bias[i]) return DESCO_EINVAL;
    a.bias[i] = bias[i];
  }
  a.nblk = num_blocks; a.nbias = num_bias; a.w_nmajor = w_nmajor; a.Y = y; a.ldy = ldy; a.M = m; a.N = n; a.act = act;
  a.accumulate = accumulate; a.slope = slope;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_dense_kernel<<<dim3((m + TM - 1) / TM, n / F), THREADS, 0, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_wgrad(const float* x, int32_t ldx, int32_t num_blocks, const float* dy, int32_t ldy, int32_t n, int32_t m,
                      float* const* dw, const int32_t* ldw, float* const* db, int32_t num_db, void* stream) {
  if (num_blocks < 1 || num_blocks > MAXB || num_db < 0 || num_db > MAXBIAS || m < 0 || n <= 0 || n % F) return DESCO_EINVAL;
  if (m == 0) return DESCO_OK;
  if (!x || !dy || !dw || !ldw || (num_db && !db) || (ldx & 3) || (ldy & 3)) return DESCO_EINVAL;
  WgradArgs a;
  a.X = x; a.ldx = ldx; a.dY = dy; a.ldy = ldy;
  for (int b = 0; b < num_blocks; ++b) {
    if (!dw[b]) return DESCO_EINVAL;
    a.dW[b] = dw[b]; a.ldw[b] = ldw[b];
  }
  for (int i = 0; i < num_db; ++i) {
    if (!db[i]) return DESCO_EINVAL;
    a.db[i] = db[i];
  }
  a.nblk = num_blocks; a.ndb = num_db; a.M = m; a.N = n;
  // enough row chunks to fill the GPU, each a multiple of the 64-row tile
  const int tiles = num_blocks * (n / F);
  int splits = (2 * desco_num_sms() + tiles - 1) / tiles;
  const int max_splits = (m + TM - 1) / TM;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  a.chunk = ((m + splits - 1) / splits + TM - 1) / TM * TM;
  splits = (m + a.chunk - 1) / a.chunk;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_wgrad_kernel<<<dim3(num_blocks, n / F, splits), THREADS, 0, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_act_backward(float* dx, int32_t ldd, const float* fwd, int32_t ldf, int32_t m, int32_t c, int32_t act,
                             float slope, void* stream) {
  if (m < 0 || c < 0 || act < ACT_RELU || act > ACT_LEAKY) return DESCO_EINVAL;
  if (m == 0 || c == 0) return DESCO_OK;
  if (!dx || !fwd) return DESCO_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_act_backward_kernel<<<blocks_for((long long)m * c, 256), 256, 0, s>>>(dx, ldd, fwd, ldf, m, c, act, slope);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_fill_rows(float* y, int32_t ldy, int32_t m, int32_t c, const float* bias, void* stream) {
  if (m < 0 || c < 0) return DESCO_EINVAL;
  if (m == 0 || c == 0) return DESCO_OK;
  if (!y || !bias) return DESCO_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_fill_rows_kernel<<<blocks_for((long long)m * c, 256), 256, 0, s>>>(y, ldy, m, c, bias);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_colsum(const float* dy, int32_t ldy, int32_t m, int32_t c, float* db, void* stream) {
  if (m < 0 || c < 0) return DESCO_EINVAL;
  if (m == 0 || c == 0) return DESCO_OK;
  if (!dy || !db) return DESCO_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  int splits = (m + 255) / 256;
  if (splits > 256) splits = 256;
  train_colsum_kernel<<<dim3((c + 63) / 64, splits), 64, 0, s>>>(dy, ldy, m, c, db);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_pool(const int32_t* nbh_ptr, const int32_t* crow_nbh, int32_t num_neighborhoods, int32_t num_count_rows,
                     int32_t hetero, int32_t c, int32_t backward, float* emb_c, int32_t ldc, const float* z_a,
                     int32_t lda, float* pooled, int32_t ldp, void* stream) {
  if (num_neighborhoods < 0 || num_count_rows < 0 || c < 0) return DESCO_EINVAL;
  if (num_neighborhoods == 0 || c == 0) return DESCO_OK;
  if (!nbh_ptr || !crow_nbh || !emb_c || !pooled) return DESCO_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  if (!backward) {
    train_pool_kernel<<<blocks_for((long long)num_neighborhoods * c, 256), 256, 0, s>>>(nbh_ptr, num_neighborhoods, hetero, c,
                                                                                       emb_c, ldc, z_a, lda, pooled, ldp);
  } else {
    if (num_count_rows == 0) return DESCO_OK;
    train_pool_backward_kernel<<<blocks_for((long long)num_count_rows * c, 256), 256, 0, s>>>(crow_nbh, num_count_rows, c,
                                                                                             pooled, ldp, emb_c, ldc);
  }
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_head_loss(const float* t, const float* bq, const float* w2, const float* b2, const float* y,
                          int32_t num_neighborhoods, int32_t num_queries, float* pred, float* dpred, float* loss,
                          void* stream) {
  const int G = num_neighborhoods, Q = num_queries;
  if (G < 0 || Q < 0 || Q > HEAD_MAXQ) return Q > HEAD_MAXQ ? DESCO_ERANGE : DESCO_EINVAL;
  if (G == 0 || Q == 0) return DESCO_OK;
  if (!t || !bq || !w2 || !b2 || (y && (!dpred || !loss)) || (!y && !pred)) return DESCO_EINVAL;
  const size_t smem = (size_t)(HEAD_TG * (HEAD_H + 1) + HEAD_H + Q * (HEAD_H + 1)) * sizeof(float);
  DESCO_CUDA_TRY(cudaFuncSetAttribute(train_head_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_head_loss_kernel<<<(G + HEAD_TG - 1) / HEAD_TG, THREADS, smem, s>>>(t, bq, w2, b2, y, G, Q, pred, dpred, loss);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_head_backward(const float* t, const float* bq, const float* w2, const float* dpred,
                              int32_t num_neighborhoods, int32_t num_queries, float* dt, float* dbq, float* dw2,
                              float* db2, void* stream) {
  const int G = num_neighborhoods, Q = num_queries;
  if (G < 0 || Q < 0 || Q > HEAD_MAXQ) return Q > HEAD_MAXQ ? DESCO_ERANGE : DESCO_EINVAL;
  if (G == 0 || Q == 0) return DESCO_OK;
  if (!t || !bq || !w2 || !dpred || !dt || !dbq || !dw2 || !db2) return DESCO_EINVAL;
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_head_backward_kernel<<<(G + HEAD_TG - 1) / HEAD_TG, HEAD_H, 0, s>>>(t, bq, w2, dpred, G, Q, dt, dbq, dw2, db2);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                     float eps, float weight_decay, int32_t step, void* stream) {
  if (n < 0 || step < 1) return DESCO_EINVAL;
  if (n == 0) return DESCO_OK;
  if (!p || !g || !m || !v) return DESCO_EINVAL;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  cudaStream_t s = (cudaStream_t)stream;
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  train_adam_kernel<<<blocks_for(n, 256), 256, 0, s>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, (float)bc1,
                                                      (float)sqrt(bc2));
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // extern "C"
