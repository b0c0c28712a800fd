// Stand-alone message-passing primitives behind the callable leaf modules SAGEConv / GossipConv, sm_100a.
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/gnn_model.py:392-394  SAGEConv:  propagate(aggr="add") = index_select + scatter_add
//   subgraph_counting/gnn_model.py:326-343  GossipConv: per-edge gated messages summed at the target
//   subgraph_counting/gnn_model.py:294-301  GossipConv.lin_gate (Linear, Sigmoid, Linear, Sigmoid, LeakyReLU)
// The hot path never calls these one by one (the whole layer stack is fused in shmp_fused.cu / shmp_mt.cu / gossip.cu);
// they exist so that code written against the reference's module API keeps working on device memory.
#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {

// out[i][:] = sum over the CSR row i of w_e * x[col[e]][:]  (w_e = 1 when edge_w == NULL); one warp per row
__global__ void spmm_sum_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                const float* __restrict__ edge_w, int n_dst, const float* __restrict__ x, int ldx, int F,
                                float* __restrict__ out, int ldo) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_dst) return;
  const int lane = lane_id();
  const int eb = rowptr[i], ee = rowptr[i + 1];
  for (int f0 = 0; f0 < F; f0 += 128) {  // 4 features per lane and pass
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int e = eb; e < ee; ++e) {
      const float w = edge_w ? edge_w[e] : 1.f;
      const float* src = x + (size_t)col[e] * ldx;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int f = f0 + 32 * u + lane;
        if (f < F) acc[u] = fmaf(w, src[f], acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int f = f0 + 32 * u + lane;
      if (f < F) out[(size_t)i * ldo + f] = acc[u];
    }
  }
}

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// gate[q] = leaky_0.01(sigmoid(w2 . sigmoid(W1 qemb[q] + b1) + b2)); one warp per query, hidden width H <= 1024
__global__ void gossip_gate_kernel(const float* __restrict__ qemb, int Q, int E, const float* __restrict__ W1,
                                   const float* __restrict__ b1, int H, const float* __restrict__ w2,
                                   const float* __restrict__ b2, float* __restrict__ gate) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const int lane = lane_id();
  float part = 0.f;
  for (int h = lane; h < H; h += 32) {
    float acc = b1[h];
    for (int k = 0; k < E; ++k) acc = fmaf(qemb[(size_t)q * E + k], W1[(size_t)h * E + k], acc);
    part = fmaf(sigmoidf_(acc), w2[h], part);
  }
  part = warp_sum(part);
  if (lane == 0) {
    const float g = sigmoidf_(part + b2[0]);
    gate[q] = g > 0.f ? g : 0.01f * g;
  }
}

}  // namespace

extern "C" {

int desco_spmm_sum(const int32_t* rowptr, const int32_t* col, const float* edge_w, int32_t n_dst, const float* x,
                   int32_t ldx, int32_t width, float* out, int32_t ldo, void* stream) {
  if (n_dst < 0 || width < 1 || ldx < width || ldo < width) return DESCO_EINVAL;
  if (n_dst == 0) return DESCO_OK;
  if (!rowptr || !out || (!col && false)) return DESCO_EINVAL;
  desco_count_launches(1);
  const long long threads = (long long)n_dst * 32;
  spmm_sum_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, edge_w, n_dst, x, ldx,
                                                                                        width, out, ldo);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_gate(const float* query_emb, int32_t num_queries, int32_t emb_channels, const float* w1, const float* b1,
                      int32_t hidden, const float* w2, const float* b2, float* gate, void* stream) {
  if (num_queries < 0 || emb_channels < 1 || hidden < 1) return DESCO_EINVAL;
  if (num_queries == 0) return DESCO_OK;
  if (!query_emb || !w1 || !b1 || !w2 || !b2 || !gate) return DESCO_EINVAL;
  desco_count_launches(1);
  gossip_gate_kernel<<<(num_queries * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(query_emb, num_queries, emb_channels,
                                                                                       w1, b1, hidden, w2, b2, gate);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // extern "C"
