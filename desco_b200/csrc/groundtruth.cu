// Ground-truth canonical counts on the GPU, sm_100a (integer / bitset work).
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/workload.py:327-348  MatchSubgraphWorker: networkx VF2 subgraph_isomorphisms_iter per (target,
//                                          query); every mapping is credited to canonical_node = max(vmap.keys())
//   subgraph_counting/workload.py:551-726  compute_groundtruth: count / SymmetricFactor(query) per node and query
//   subgraph_counting/data.py:61-88        SymmetricFactor = |Aut(query)| (number of self-mappings)
// VF2 enumerates every induced-subgraph isomorphism, i.e. |Aut(Q)| mappings per node set; dividing by the symmetry factor
// leaves   truth[v][q] = #{ node sets S : max(S) = v, G[S] isomorphic to query q }.
// That is computed directly: for every root v, ESU (Wernicke) enumerates each CONNECTED induced subgraph of 3..5 nodes
// whose largest node is v exactly once (extension only through nodes < v, exclusive-neighbourhood rule), and the
// k(k-1)/2 adjacency bits of the node tuple index a host-built table that maps the pattern to its query column.
// One CTA per target graph (adjacency bit-matrix in shared memory, graphs up to 1024 nodes), one warp per root, the 32
// lanes split the second extension level; every set is a W-word bitset.
#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {

constexpr int GT_THREADS = 256;
constexpr int MAXQ = 64;

template <int W>
struct Bits {
  uint32_t w[W];
  __device__ __forceinline__ bool any() const {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) o |= w[i];
    return o != 0;
  }
  __device__ __forceinline__ int count() const {
    int c = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) c += __popc(w[i]);
    return c;
  }
  __device__ __forceinline__ bool test(int b) const { return (w[b >> 5] >> (b & 31)) & 1u; }
  // remove and return the lowest set bit (-1 when empty)
  __device__ __forceinline__ int pop() {
#pragma unroll
    for (int i = 0; i < W; ++i)
      if (w[i]) {
        const int b = __ffs(w[i]) - 1;
        w[i] &= w[i] - 1;
        return 32 * i + b;
      }
    return -1;
  }
  // the t-th set bit (0-based), -1 when there are not that many
  __device__ __forceinline__ int nth(int t) const {
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int c = __popc(w[i]);
      if (t < c) return 32 * i + (int)__fns(w[i], 0, t + 1);
      t -= c;
    }
    return -1;
  }
};

template <int W>
__device__ __forceinline__ Bits<W> row_of(const uint32_t* adj, int u) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = adj[u * W + i];
  return r;
}

// bits strictly above position b
template <int W>
__device__ __forceinline__ Bits<W> above(const Bits<W>& x, int b) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) {
    const int lo = 32 * i;
    uint32_t m = 0xffffffffu;
    if (b >= lo + 31) m = 0u;
    else if (b >= lo) m = 0xffffffffu << ((b - lo) + 1);
    r.w[i] = x.w[i] & m;
  }
  return r;
}

// ESU extension: Vext' = (Vext minus the nodes already taken) | (N(w) & below(v) & ~closedN(Vsub))
template <int W>
__device__ __forceinline__ Bits<W> extend(const Bits<W>& rem, const Bits<W>& nw, const Bits<W>& below, const Bits<W>& closed) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = rem.w[i] | (nw.w[i] & below.w[i] & ~closed.w[i]);
  return r;
}

template <int W>
__device__ __forceinline__ Bits<W> unite(const Bits<W>& a, const Bits<W>& b) {
  Bits<W> r;
#pragma unroll
  for (int i = 0; i < W; ++i) r.w[i] = a.w[i] | b.w[i];
  return r;
}

__device__ __forceinline__ void credit(unsigned long long* cnt, const uint8_t* lut, int pattern) {
  const int q = lut[pattern];
  if (q != 255) atomicAdd(&cnt[q], 1ull);
}

template <int W>
__global__ void __launch_bounds__(GT_THREADS) groundtruth_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ graph_ptr,
    int num_graphs, const uint8_t* __restrict__ lut3, const uint8_t* __restrict__ lut4, const uint8_t* __restrict__ lut5,
    int Q, int max_k, long long* __restrict__ out, int32_t* __restrict__ status) {
  extern __shared__ uint32_t gsm[];
  __shared__ uint8_t s_lut3[8], s_lut4[64], s_lut5[1024];
  __shared__ unsigned long long s_cnt[GT_THREADS / 32][MAXQ];
  __shared__ int s_ticket;
  uint32_t* adj = gsm;  // [n][W]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 1024; i += GT_THREADS) {
    s_lut5[i] = lut5[i];
    if (i < 64) s_lut4[i] = lut4[i];
    if (i < 8) s_lut3[i] = lut3[i];
  }
  for (int gidx = blockIdx.x; gidx < num_graphs; gidx += gridDim.x) {
    const int base = graph_ptr[gidx], n = graph_ptr[gidx + 1] - base;
    __syncthreads();
    if (n > 32 * W) {
      if (tid == 0) atomicExch(status, DESCO_ERANGE);
      continue;
    }
    for (int i = tid; i < n * W; i += GT_THREADS) adj[i] = 0u;
    if (tid == 0) s_ticket = 0;
    __syncthreads();
    for (int u = warp; u < n; u += GT_THREADS / 32)
      for (int e = rowptr[base + u] + lane; e < rowptr[base + u + 1]; e += 32) {
        const int t = col[e] - base;
        if (t != u) atomicOr(&adj[u * W + (t >> 5)], 1u << (t & 31));
      }
    __syncthreads();

    for (;;) {  // one warp per root v (dealt by a ticket: roots with high degree cost far more)
      int v = 0;
      if (lane == 0) v = atomicAdd(&s_ticket, 1);
      v = __shfl_sync(FULL_MASK, v, 0);
      if (v >= n) break;
      unsigned long long* cnt = s_cnt[warp];
      for (int q = lane; q < Q; q += 32) cnt[q] = 0ull;
      __syncwarp();
      Bits<W> below;  // nodes < v
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const int lo = 32 * i;
        below.w[i] = v >= lo + 32 ? 0xffffffffu : (v > lo ? (0xffffffffu >> (32 - (v - lo))) : 0u);
      }
      const Bits<W> av = row_of<W>(adj, v);
      Bits<W> cl1 = av;
      cl1.w[v >> 5] |= 1u << (v & 31);
      Bits<W> ext1;
#pragma unroll
      for (int i = 0; i < W; ++i) ext1.w[i] = av.w[i] & below.w[i];
      Bits<W> it1 = ext1;
      for (int w1 = it1.pop(); w1 >= 0; w1 = it1.pop()) {
        const Bits<W> a1 = row_of<W>(adj, w1);
        const Bits<W> ext2 = extend<W>(it1, a1, below, cl1);  // it1 = the part of ext1 not taken yet
        const Bits<W> cl2 = unite<W>(cl1, a1);
        const int n2 = ext2.count();
        for (int t = lane; t < n2; t += 32) {  // the lanes split the second extension level
          const int w2 = ext2.nth(t);
          const int p3 = 1 | (av.test(w2) ? 2 : 0) | (a1.test(w2) ? 4 : 0);  // pairs (0,1) (0,2) (1,2)
          credit(cnt, s_lut3, p3);
          if (max_k < 4) continue;
          const Bits<W> a2 = row_of<W>(adj, w2);
          Bits<W> it3 = extend<W>(above<W>(ext2, w2), a2, below, cl2);
          const Bits<W> cl3 = unite<W>(cl2, a2);
          for (int w3 = it3.pop(); w3 >= 0; w3 = it3.pop()) {
            const int p4 = p3 | (av.test(w3) ? 8 : 0) | (a1.test(w3) ? 16 : 0) | (a2.test(w3) ? 32 : 0);
            credit(cnt, s_lut4, p4);
            if (max_k < 5) continue;
            const Bits<W> a3 = row_of<W>(adj, w3);
            Bits<W> it4 = extend<W>(it3, a3, below, cl3);
            for (int w4 = it4.pop(); w4 >= 0; w4 = it4.pop()) {
              const int p5 = p4 | (av.test(w4) ? 64 : 0) | (a1.test(w4) ? 128 : 0) | (a2.test(w4) ? 256 : 0) |
                             (a3.test(w4) ? 512 : 0);
              credit(cnt, s_lut5, p5);
            }
          }
        }
      }
      __syncwarp();
      for (int q = lane; q < Q; q += 32) out[(size_t)(base + v) * Q + q] = (long long)cnt[q];
      __syncwarp();
    }
  }
}

template <int W>
int launch_gt(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int num_graphs, int max_graph_nodes,
              const uint8_t* lut3, const uint8_t* lut4, const uint8_t* lut5, int Q, int max_k, long long* out,
              int32_t* status, cudaStream_t s) {
  const size_t smem = (size_t)max_graph_nodes * W * sizeof(uint32_t);
  DESCO_CUDA_TRY(cudaFuncSetAttribute(groundtruth_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = num_graphs < 4 * desco_num_sms() ? num_graphs : 4 * desco_num_sms();
  desco_count_launches(1);
  groundtruth_kernel<W><<<grid, GT_THREADS, smem, s>>>(rowptr, col, graph_ptr, num_graphs, lut3, lut4, lut5, Q, max_k, out,
                                                       status);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // namespace

extern "C" {

int desco_groundtruth_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                            int32_t max_graph_nodes, const uint8_t* lut3, const uint8_t* lut4, const uint8_t* lut5,
                            int32_t num_queries, int32_t max_query_nodes, int64_t* out_counts, int32_t* status,
                            void* stream) {
  if (num_graphs < 0 || max_graph_nodes < 1 || num_queries < 1 || num_queries > MAXQ || max_query_nodes < 3 ||
      max_query_nodes > 5)
    return DESCO_EINVAL;
  if (num_graphs == 0) return DESCO_OK;
  if (!rowptr || !col || !graph_ptr || !lut3 || !lut4 || !lut5 || !out_counts || !status) return DESCO_EINVAL;
  if (max_graph_nodes > 1024) return DESCO_ERANGE;  // the adjacency bit-matrix of a graph lives in shared memory
  const int words = (max_graph_nodes + 31) / 32;
  cudaStream_t s = (cudaStream_t)stream;
#define DESCO_GT(Wc)                                                                                                  \
  return launch_gt<Wc>(rowptr, col, graph_ptr, num_graphs, max_graph_nodes, lut3, lut4, lut5, num_queries,          \
                       max_query_nodes, (long long*)out_counts, status, s)
  if (words <= 1) DESCO_GT(1);
  if (words <= 2) DESCO_GT(2);
  if (words <= 4) DESCO_GT(4);
  if (words <= 8) DESCO_GT(8);
  if (words <= 16) DESCO_GT(16);
  DESCO_GT(32);
#undef DESCO_GT
}

}  // extern "C"
