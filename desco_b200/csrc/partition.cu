// Canonical neighborhood partition + SHMP edge typing, sm_100a.
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/data.py:329-338   k_neigh              unrestricted k-hop set BFS
//   subgraph_counting/data.py:341-350   k_neigh_canonical    BFS through nodes <= centre only
//   subgraph_counting/data.py:353-396   get_neigh_canonical / get_neigh_hetero  (<= filter, component of the centre)
//   subgraph_counting/workload.py:243-260  NeighborhoodDataset.process loop (drop edge-free neighborhoods, indicator, index)
//   subgraph_counting/transforms.py:319-412  NetworkxToHetero (emitted directly as a packed CSR batch)
//   subgraph_counting/transforms.py:180-255  ToTconvHetero   (triangle / tride flag per directed edge)
//
// One "group" of threads owns one centre: a warp when every target graph fits 2048 nodes (molecule / ego datasets),
// a whole CTA otherwise.  All set state (visited / frontier / next / component) is a bitset over the centre's OWN
// target graph [lo,hi) staged in shared memory; adjacency is read from the CSR in HBM, hubs (deg >= 32) with all 32
// lanes on consecutive col[] entries, low-degree nodes one lane each.  The neighborhood is emitted sorted by node id,
// so the canonical node (the max, workload.py:346) is always the last row of its neighborhood.
#include <cub/cub.cuh>

#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {

enum ExpandMode { EXP_ALL = 0, EXP_LE = 1, EXP_CAND = 2 };

template <bool CTA>
struct Group {
  __device__ static int size() { return CTA ? blockDim.x : 32; }
  __device__ static int tid() { return CTA ? threadIdx.x : lane_id(); }
  __device__ static int nwarps() { return CTA ? (blockDim.x >> 5) : 1; }
  __device__ static int warp() { return CTA ? warp_id() : 0; }
  __device__ static void sync() {
    if (CTA) __syncthreads(); else __syncwarp();
  }
  __device__ static bool any(bool p) {
    if (CTA) return __syncthreads_or(p ? 1 : 0) != 0;
    return __any_sync(FULL_MASK, p);
  }
};

__device__ __forceinline__ bool bit_test(const uint32_t* b, int i) { return (b[i >> 5] >> (i & 31)) & 1u; }

// next |= union of adj(u) for u in frontier, filtered:  EXP_ALL: every neighbour not yet in `seen`;
// EXP_LE: additionally neighbour <= centre;  EXP_CAND: additionally neighbour in `cand`.
template <bool CTA>
__device__ void expand(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const uint32_t* frontier,
                       uint32_t* next, const uint32_t* seen, const uint32_t* cand, int W, int lo, int centre, int mode) {
  const int lane = lane_id();
  for (int wbase = Group<CTA>::warp() * 32; wbase < W; wbase += 32 * Group<CTA>::nwarps()) {
    uint32_t word = (wbase + lane < W) ? frontier[wbase + lane] : 0u;
    while (__any_sync(FULL_MASK, word != 0u)) {
      int beg = 0, end = 0;
      if (word) {
        int b = __ffs(word) - 1;
        word &= word - 1;
        int u = lo + 32 * (wbase + lane) + b;
        beg = rowptr[u];
        end = rowptr[u + 1];
      }
      auto mark = [&](int v) {
        if (mode != EXP_ALL && v > centre) return;
        int i = v - lo;
        uint32_t bit = 1u << (i & 31);
        if (seen[i >> 5] & bit) return;
        if (mode == EXP_CAND && !(cand[i >> 5] & bit)) return;
        atomicOr(&next[i >> 5], bit);
      };
      // hubs: all 32 lanes stream one adjacency list (coalesced 128 B reads)
      uint32_t hubs = __ballot_sync(FULL_MASK, end - beg >= 32);
      while (hubs) {
        int src = __ffs(hubs) - 1;
        hubs &= hubs - 1;
        int hb = __shfl_sync(FULL_MASK, beg, src), he = __shfl_sync(FULL_MASK, end, src);
        for (int e = hb + lane; e < he; e += 32) {
          int v = col[e];
          if (mode != EXP_ALL && v > centre) break;  // sorted adjacency: nothing further can pass
          mark(v);
        }
      }
      if (end - beg < 32) {
        for (int e = beg; e < end; ++e) {
          int v = col[e];
          if (mode != EXP_ALL && v > centre) break;
          mark(v);
        }
      }
    }
  }
}

// frontier = next & ~seen ; seen |= frontier ; next = 0.  Returns group-wide "frontier non-empty".
template <bool CTA>
__device__ bool advance(uint32_t* seen, uint32_t* frontier, uint32_t* next, int W) {
  bool any = false;
  for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
    uint32_t nf = next[w] & ~seen[w];
    next[w] = 0u;
    seen[w] |= nf;
    frontier[w] = nf;
    any |= (nf != 0u);
  }
  Group<CTA>::sync();  // vote functions do not order shared-memory traffic by themselves
  return Group<CTA>::any(any);
}

// Runs the partition of one centre; on return S (W words) holds the neighborhood node set and pref[w] the number of
// set bits in words < w.  Returns |S|.
template <bool CTA>
__device__ int partition_one(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, uint32_t* S,
                             uint32_t* F, uint32_t* Nx, uint32_t* aux, int W, int lo, int centre, int depth, int mode) {
  const int cl = centre - lo;
  for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
    uint32_t init = (w == (cl >> 5)) ? (1u << (cl & 31)) : 0u;
    S[w] = init;
    F[w] = init;
    Nx[w] = 0u;
  }
  Group<CTA>::sync();
  // phase A: k rounds of frontier expansion (data.py:332-337 / :344-349)
  for (int level = 0; level < depth; ++level) {
    int m = (mode == DESCO_MODE_KHOP) ? EXP_ALL
            : (mode == DESCO_MODE_CANONICAL || level == depth - 1) ? EXP_LE : EXP_ALL;  // last ring only matters <= centre
    expand<CTA>(rowptr, col, F, Nx, S, nullptr, W, lo, centre, m);
    Group<CTA>::sync();
    if (!advance<CTA>(S, F, Nx, W)) break;
  }
  if (mode == DESCO_MODE_HETERO) {
    // phase B: keep candidates <= centre (data.py:385); phase C: component of the centre inside them (:387-390)
    uint32_t* comp = aux;
    for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
      uint32_t keep = (w < (cl >> 5)) ? 0xffffffffu : (w == (cl >> 5) ? (0xffffffffu >> (31 - (cl & 31))) : 0u);
      S[w] &= keep;
      uint32_t init = (w == (cl >> 5)) ? (1u << (cl & 31)) : 0u;
      comp[w] = init;
      F[w] = init;
    }
    Group<CTA>::sync();
    while (true) {
      expand<CTA>(rowptr, col, F, Nx, comp, S, W, lo, centre, EXP_CAND);
      Group<CTA>::sync();
      if (!advance<CTA>(comp, F, Nx, W)) break;
    }
    for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) S[w] = comp[w];
    Group<CTA>::sync();
  }
  // phase D: popcount prefix -> dense local ids in ascending node order
  int* pref = reinterpret_cast<int*>(aux);
  int total = 0;
  if (Group<CTA>::warp() == 0) {
    const int lane = lane_id();
    int carry = 0;
    for (int base = 0; base < W; base += 32) {
      int c = (base + lane < W) ? __popc(S[base + lane]) : 0;
      int incl = warp_incl_scan(c);
      if (base + lane < W) pref[base + lane] = carry + incl - c;
      carry += __shfl_sync(FULL_MASK, incl, 31);
    }
    total = carry;
  }
  if (CTA) {
    __shared__ int s_total;
    if (threadIdx.x == 0) s_total = total;
    __syncthreads();
    total = s_total;
  } else {
    __syncwarp();
  }
  return total;
}

__device__ __forceinline__ int local_index(const uint32_t* S, const int* pref, int i) {
  return pref[i >> 5] + __popc(S[i >> 5] & ((1u << (i & 31)) - 1u));
}

// does u--v close a triangle inside S?  sorted-list intersection of adj(u), adj(v) restricted to S
// ( (A*A^2 + A)[u,v] > 1  <=>  common neighbour in the neighborhood; transforms.py:201-221 )
__device__ bool has_common_neighbour(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                     const uint32_t* S, int lo, int centre, int u, int v) {
  int a = rowptr[u], ae = rowptr[u + 1], b = rowptr[v], be = rowptr[v + 1];
  while (a < ae && b < be) {
    int x = col[a], y = col[b];
    if (x > centre || y > centre) return false;  // S has nothing above the centre
    if (x == y) {
      if (bit_test(S, x - lo)) return true;
      ++a;
      ++b;
    } else if (x < y) {
      ++a;
    } else {
      ++b;
    }
  }
  return false;
}

template <bool CTA>
__global__ void __launch_bounds__(256) partition_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ graph_ptr,
    int num_graphs, const int32_t* __restrict__ centres, int num_centres, int depth, int mode, int max_words,
    // count pass outputs (fill == 0)
    int32_t* __restrict__ out_nv, int32_t* __restrict__ out_ne, int32_t* __restrict__ centre_graph,
    // fill pass inputs / outputs (fill == 1)
    int fill, const int32_t* __restrict__ node_off, const int32_t* __restrict__ edge_off,
    int32_t* __restrict__ node_gid, int32_t* __restrict__ edge_ptr, int32_t* __restrict__ edge_col,
    uint8_t* __restrict__ edge_tri, int32_t* __restrict__ status,
    // dense tier of the large-graph path: bitsets in a per-CTA slice of global scratch, only centres of class `klass_want`
    uint32_t* __restrict__ gscratch = nullptr, const uint8_t* __restrict__ klass = nullptr, int klass_want = 0) {
  extern __shared__ uint32_t smem[];
  const int groups_per_cta = CTA ? 1 : (blockDim.x >> 5);
  const int g_in_cta = CTA ? 0 : warp_id();
  uint32_t* S = gscratch ? gscratch + (size_t)blockIdx.x * 4 * max_words : smem + (size_t)g_in_cta * 4 * max_words;
  uint32_t* F = S + max_words;
  uint32_t* Nx = F + max_words;
  uint32_t* aux = Nx + max_words;

  for (int ci = blockIdx.x * groups_per_cta + g_in_cta; ci < num_centres; ci += gridDim.x * groups_per_cta) {
    if (klass && klass[ci] != klass_want) continue;
    const int centre = centres[ci];
    int gid;
    if (fill) {
      gid = centre_graph[ci];
      if (out_ne[ci] == 0) continue;  // dropped neighborhood (workload.py:253-256); uniform across the group
    } else {
      int a = 0, b = num_graphs;  // largest a with graph_ptr[a] <= centre
      while (b - a > 1) {
        int mid = (a + b) >> 1;
        if (graph_ptr[mid] <= centre) a = mid; else b = mid;
      }
      gid = a;
    }
    const int lo = graph_ptr[gid], hi = graph_ptr[gid + 1];
    const int W = (hi - lo + 31) >> 5;
    if (W > max_words) {  // caller promised an upper bound that does not hold
      if (Group<CTA>::tid() == 0) atomicExch(status, DESCO_ERANGE);
      if (!fill && Group<CTA>::tid() == 0) {
        out_nv[ci] = 0;
        out_ne[ci] = 0;
        centre_graph[ci] = gid;
      }
      continue;
    }
    const int nv = partition_one<CTA>(rowptr, col, S, F, Nx, aux, W, lo, centre, depth, mode);
    const int* pref = reinterpret_cast<const int*>(aux);
    const int limit = (mode == DESCO_MODE_KHOP) ? 0x7fffffff : centre;  // S has nothing above `limit`

    if (!fill) {
      // induced directed edge count
      int cnt = 0;
      for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
        uint32_t word = S[w];
        while (word) {
          int u = lo + 32 * w + __ffs(word) - 1;
          word &= word - 1;
          for (int e = rowptr[u], ee = rowptr[u + 1]; e < ee; ++e) {
            int v = col[e];
            if (v > limit) break;
            cnt += bit_test(S, v - lo);
          }
        }
      }
      cnt = warp_sum(cnt);
      if (CTA) {
        __shared__ int s_cnt;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        if (lane_id() == 0) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        cnt = s_cnt;
      }
      if (Group<CTA>::tid() == 0) {
        out_ne[ci] = cnt;
        out_nv[ci] = cnt > 0 ? nv : 0;
        centre_graph[ci] = gid;
      }
      Group<CTA>::sync();
      continue;
    }

    // ---- fill pass ----
    const int n0 = node_off[ci], e0 = edge_off[ci];
    if (n0 == 0 && Group<CTA>::tid() == 0) edge_ptr[0] = 0;
    // rows: node ids + per-row induced degree (parked in edge_ptr[row+1])
    for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
      uint32_t word = S[w];
      int k = 0;
      while (word) {
        int u = lo + 32 * w + __ffs(word) - 1;
        word &= word - 1;
        int cnt = 0;
        for (int e = rowptr[u], ee = rowptr[u + 1]; e < ee; ++e) {
          int v = col[e];
          if (v > limit) break;
          cnt += bit_test(S, v - lo);
        }
        int row = n0 + pref[w] + k;
        node_gid[row] = u;
        edge_ptr[row + 1] = cnt;
        ++k;
      }
    }
    Group<CTA>::sync();
    if (Group<CTA>::warp() == 0) {  // in-place inclusive scan of the nv row degrees
      const int lane = lane_id();
      int carry = e0;
      for (int base = 0; base < nv; base += 32) {
        int x = (base + lane < nv) ? edge_ptr[n0 + 1 + base + lane] : 0;
        int incl = warp_incl_scan(x);
        if (base + lane < nv) edge_ptr[n0 + 1 + base + lane] = carry + incl;
        carry += __shfl_sync(FULL_MASK, incl, 31);
      }
    }
    Group<CTA>::sync();
    // edges: batch-global row index of the other endpoint + SHMP type
    for (int w = Group<CTA>::tid(); w < W; w += Group<CTA>::size()) {
      uint32_t word = S[w];
      int k = 0;
      while (word) {
        int u = lo + 32 * w + __ffs(word) - 1;
        word &= word - 1;
        int row = n0 + pref[w] + k;
        int out = (row == n0) ? e0 : edge_ptr[row];
        for (int e = rowptr[u], ee = rowptr[u + 1]; e < ee; ++e) {
          int v = col[e];
          if (v > limit) break;
          if (!bit_test(S, v - lo)) continue;
          edge_col[out] = n0 + local_index(S, pref, v - lo);
          edge_tri[out] = (gscratch == nullptr && has_common_neighbour(rowptr, col, S, lo, limit, u, v)) ? 1 : 0;
          ++out;
        }
        ++k;
      }
    }
    Group<CTA>::sync();
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Small-graph regime (every target graph <= 256 nodes: molecule / ego / protein datasets, configs 1-4).
// One warp per centre, but the whole target graph is first turned into an adjacency BIT-MATRIX in shared memory
// (lane-per-node: row u = W words), shared by the consecutive centres a warp processes.  Every set operation of the
// partition is then a handful of word-wide ANDs/ORs held REPLICATED in registers across the warp:
//   frontier expansion  next = OR_{u in F} adj[u]      lane-per-node + one __reduce_or_sync per word
//   <= centre filter    one mask per word;   component of the centre: the same expansion restricted to the mask
//   SHMP edge type      tri(u,v) = (adj[u] & adj[v] & S) != 0
//   emission            lane-per-node: local ids by popcount prefix, per-row offsets by a warp scan
// ------------------------------------------------------------------------------------------------------------------
template <int W>
__global__ void __launch_bounds__(256) partition_small_kernel(
    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ graph_ptr,
    int num_graphs, const int32_t* __restrict__ centres, int num_centres, int depth, int mode, int run,
    int32_t* __restrict__ out_nv, int32_t* __restrict__ out_ne, int32_t* __restrict__ centre_graph, int fill,
    const int32_t* __restrict__ node_off, const int32_t* __restrict__ edge_off, int32_t* __restrict__ node_gid,
    int32_t* __restrict__ edge_ptr, int32_t* __restrict__ edge_col, uint8_t* __restrict__ edge_tri,
    int32_t* __restrict__ status) {
  constexpr int PW = (W == 1) ? 1 : W + 1;  // odd row pitch: lane-per-row reads are bank-conflict free
  extern __shared__ uint32_t smem[];
  const int lane = lane_id(), warp = warp_id();
  uint32_t* adj = smem + (size_t)warp * (32 * W) * PW;
  const uint32_t lt_mask = (1u << lane) - 1u;

  int cached_gid = -1, lo = 0, hi = 0;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const int c_begin = (int)min((long long)num_centres, gw * run);
  const int c_end = (int)min((long long)num_centres, (gw + 1) * run);
  for (int ci = c_begin; ci < c_end; ++ci) {
    const int centre = centres[ci];
    if (fill && out_ne[ci] == 0) continue;  // dropped neighborhood (workload.py:253-256)
    if (centre < lo || centre >= hi) {      // a different target graph: locate it and rebuild the bit-matrix
      int gid;
      if (fill) {
        gid = centre_graph[ci];
      } else {  // 32-ary search: largest gid with graph_ptr[gid] <= centre
        int a = 0, b = num_graphs;
        while (b - a > 1) {
          const int step = (b - a + 31) >> 5;
          const int probe = min(a + lane * step, b);
          const bool le = probe < b && graph_ptr[probe] <= centre;
          const int k = __popc(__ballot_sync(FULL_MASK, le)) - 1;  // predicates are monotone in the lane index
          const int na = a + k * step;
          b = min(b, na + step);
          a = na;
        }
        gid = a;
      }
      cached_gid = gid;
      lo = graph_ptr[gid];
      hi = graph_ptr[gid + 1];
      if (hi - lo > 32 * W) {  // caller promised an upper bound that does not hold
        if (lane == 0) {
          atomicExch(status, DESCO_ERANGE);
          if (!fill) { out_nv[ci] = 0; out_ne[ci] = 0; centre_graph[ci] = gid; }
        }
        hi = lo;  // invalidate the cache
        continue;
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const int u = 32 * i + lane;
        uint32_t row[W];
#pragma unroll
        for (int j = 0; j < W; ++j) row[j] = 0u;
        if (lo + u < hi) {
          for (int e = rowptr[lo + u], ee = rowptr[lo + u + 1]; e < ee; ++e) {
            const int v = col[e] - lo;
#pragma unroll
            for (int j = 0; j < W; ++j)
              if ((v >> 5) == j) row[j] |= 1u << (v & 31);
          }
        }
#pragma unroll
        for (int j = 0; j < W; ++j) adj[u * PW + j] = row[j];
      }
      __syncwarp();
    }
    const int gid = cached_gid;
    const int cl = centre - lo;

    // next = OR of adj[u] over the nodes u of F (lane-per-node), identical in every lane afterwards
    auto expand = [&](const uint32_t (&Fr)[W], uint32_t (&nx)[W]) {
      uint32_t acc[W];
#pragma unroll
      for (int j = 0; j < W; ++j) acc[j] = 0u;
#pragma unroll
      for (int i = 0; i < W; ++i) {
        if ((Fr[i] >> lane) & 1u) {
#pragma unroll
          for (int j = 0; j < W; ++j) acc[j] |= adj[(32 * i + lane) * PW + j];
        }
      }
#pragma unroll
      for (int j = 0; j < W; ++j) nx[j] = __reduce_or_sync(FULL_MASK, acc[j]);
    };

    uint32_t S[W], Fr[W], le[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const uint32_t init = (j == (cl >> 5)) ? (1u << (cl & 31)) : 0u;
      S[j] = init;
      Fr[j] = init;
      le[j] = (mode == DESCO_MODE_KHOP) ? 0xffffffffu
              : (j < (cl >> 5)) ? 0xffffffffu : (j == (cl >> 5) ? (0xffffffffu >> (31 - (cl & 31))) : 0u);
    }
    // phase A: k rounds of frontier expansion (data.py:332-337 unrestricted / :344-349 through nodes <= centre)
    for (int level = 0; level < depth; ++level) {
      uint32_t nx[W];
      expand(Fr, nx);
      uint32_t any = 0u;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        uint32_t f = nx[j] & ~S[j];
        if (mode == DESCO_MODE_CANONICAL) f &= le[j];
        Fr[j] = f;
        S[j] |= f;
        any |= f;
      }
      if (!any) break;
    }
    if (mode == DESCO_MODE_HETERO) {
      // phase B: keep candidates <= centre (data.py:385); phase C: component of the centre inside them (:387-390)
      uint32_t comp[W];
#pragma unroll
      for (int j = 0; j < W; ++j) {
        S[j] &= le[j];
        const uint32_t init = (j == (cl >> 5)) ? (1u << (cl & 31)) : 0u;
        comp[j] = init;
        Fr[j] = init;
      }
      while (true) {
        uint32_t nx[W];
        expand(Fr, nx);
        uint32_t any = 0u;
#pragma unroll
        for (int j = 0; j < W; ++j) {
          const uint32_t f = nx[j] & S[j] & ~comp[j];
          Fr[j] = f;
          comp[j] |= f;
          any |= f;
        }
        if (!any) break;
      }
#pragma unroll
      for (int j = 0; j < W; ++j) S[j] = comp[j];
    }
    // phase D: popcount prefix -> dense local ids in ascending node order; induced degrees, lane-per-node
    int pref[W];
    int nv = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
      pref[j] = nv;
      nv += __popc(S[j]);
    }
    uint32_t nb[W][W];  // nb[i] = induced adjacency row of node 32 i + lane
    int deg[W];
    int my_edges = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const bool in = (S[i] >> lane) & 1u;
      deg[i] = 0;
#pragma unroll
      for (int j = 0; j < W; ++j) {
        nb[i][j] = in ? (adj[(32 * i + lane) * PW + j] & S[j]) : 0u;
        deg[i] += __popc(nb[i][j]);
      }
      my_edges += deg[i];
    }
    const int ne = warp_sum(my_edges);
    if (!fill) {
      if (lane == 0) {
        out_ne[ci] = ne;
        out_nv[ci] = ne > 0 ? nv : 0;
        centre_graph[ci] = gid;
      }
      continue;
    }
    // ---- fill pass ----
    const int n0 = node_off[ci], e0 = edge_off[ci];
    if (n0 == 0 && lane == 0) edge_ptr[0] = 0;
    int carry = e0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int incl = warp_incl_scan(deg[i]);
      int out = carry + incl - deg[i];
      carry += __shfl_sync(FULL_MASK, incl, 31);
      if ((S[i] >> lane) & 1u) {
        const int u = 32 * i + lane;
        const int row = n0 + pref[i] + __popc(S[i] & lt_mask);
        node_gid[row] = lo + u;
        edge_ptr[row + 1] = out + deg[i];
#pragma unroll
        for (int j = 0; j < W; ++j) {
          uint32_t bits = nb[i][j];
          while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int v = 32 * j + b;
            uint32_t common = 0u;  // (A*A^2 + A)[u,v] > 1  <=>  common neighbour inside S (transforms.py:201-221)
#pragma unroll
            for (int k = 0; k < W; ++k) common |= nb[i][k] & adj[v * PW + k];
            edge_col[out] = n0 + pref[j] + __popc(S[j] & ((1u << b) - 1u));
            edge_tri[out] = common ? 1 : 0;
            ++out;
          }
        }
      }
    }
  }
}

template <int W>
int launch_partition_small(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int num_graphs,
                           const int32_t* centres, int num_centres, int depth, int mode, int32_t* nv, int32_t* ne,
                           int32_t* centre_graph, int fill, const int32_t* node_off, const int32_t* edge_off,
                           int32_t* node_gid, int32_t* edge_ptr, int32_t* edge_col, uint8_t* edge_tri, int32_t* status,
                           cudaStream_t stream) {
  constexpr int PW = (W == 1) ? 1 : W + 1;
  const int warps = 8, threads = warps * 32;
  const size_t smem = (size_t)warps * 32 * W * PW * sizeof(uint32_t);
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(partition_small_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int sms = desco_num_sms();
  // consecutive centres usually share a target graph: a warp keeps its bit-matrix across a short run of them, but
  // never so long that the grid drops below ~4 warps per scheduler
  int run = num_centres / (sms * warps * 4);
  run = run < 1 ? 1 : (run > 8 ? 8 : run);
  const long long total_warps = ((long long)num_centres + run - 1) / run;
  const unsigned blocks = (unsigned)((total_warps + warps - 1) / warps);
  partition_small_kernel<W><<<blocks, threads, smem, stream>>>(rowptr, col, graph_ptr, num_graphs, centres, num_centres,
                                                              depth, mode, run, nv, ne, centre_graph, fill, node_off,
                                                              edge_off, node_gid, edge_ptr, edge_col, edge_tri, status);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

// do two ascending index lists share an element?  Similar lengths: linear merge; a short list against a long one
// (a leaf against a hub row - power-law targets): gallop through the long list by binary search, O(short * log long).
__device__ __forceinline__ bool sorted_lists_intersect(const int32_t* __restrict__ idx, int a, int ae, int b, int be) {
  if (ae - a > be - b) {
    int t = a; a = b; b = t;
    t = ae; ae = be; be = t;
  }
  if (ae == a) return false;
  if (be - b < 8 * (ae - a)) {
    while (a < ae && b < be) {
      const int x = idx[a], y = idx[b];
      if (x == y) return true;
      if (x < y) ++a; else ++b;
    }
    return false;
  }
  for (; a < ae && b < be; ++a) {
    const int x = idx[a];
    int lo = b, hi = be;  // first position with idx[pos] >= x
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (idx[mid] < x) lo = mid + 1; else hi = mid;
    }
    if (lo < be && idx[lo] == x) return true;
    b = lo;
  }
  return false;
}

// ToTconvHetero on an existing packed batch: one warp per row, one lane per incident edge.
__global__ void edge_types_kernel(const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ edge_col,
                                  int num_rows, uint8_t* __restrict__ edge_tri) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= num_rows) return;
  const int rb = edge_ptr[row], re = edge_ptr[row + 1];
  for (int e = rb + lane_id(); e < re; e += 32) {
    const int v = edge_col[e];
    edge_tri[e] = sorted_lists_intersect(edge_col, rb, re, edge_ptr[v], edge_ptr[v + 1]) ? 1 : 0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Large-graph regime (config 5: one 10M-node / 200M-directed-edge power-law target).  A bitset over the target no
// longer fits shared memory and would cost O(N/32) per centre, so the set state of a centre is SPARSE:
//   * an open-addressing hash set of node ids (linear probing, atomicCAS insert; bit 31 of a key = "reached" flag),
//   * the member list L in discovery order (BFS levels are contiguous slices of it),
//   * the reached list R (component of the centre), bitonic-sorted at the end so that rows come out in ascending node
//     id and a local id is a binary search.
// Tier 0 keeps all three in shared memory; a centre whose ball overflows them is re-run by tier 1 (same code, tables in
// a per-CTA slice of global scratch, L2-resident), and a ball that overflows that too by the dense tier (the bitset
// kernel above with its four bitsets in global scratch).  The sorted adjacency makes "<= centre" a PREFIX of every
// row, so the restricted passes stop at the first neighbour above the centre.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t SP_EMPTY = 0xffffffffu;
constexpr uint32_t SP_FLAG = 0x80000000u;
constexpr uint32_t SP_MASK = 0x7fffffffu;
constexpr int SP_THREADS = 256;

struct SparseArgs {
  const int32_t* rowptr; const int32_t* col; const int32_t* graph_ptr; int num_graphs;
  const int32_t* centres; int num_centres, depth, mode;
  int32_t* out_nv; int32_t* out_ne; int32_t* centre_graph;
  int fill; const int32_t* node_off; const int32_t* edge_off;
  int32_t* node_gid; int32_t* edge_ptr; int32_t* edge_col;
  uint8_t* klass; int tier;      // centre class: 0 = tier 0 (shared memory), 1 = tier 1 (global scratch), 2 = dense tier
  uint32_t* gscratch;            // tier 1: per-CTA slice of (H + capL + capR) words
  int log2H, capL, capR;         // hash slots (power of two), member-list / reached-list capacities (capR power of two)
};

struct SparseSet {
  uint32_t* keys; uint32_t* L; uint32_t* R;
  int H, log2H, capL, capR;
  int* nL; int* nR; int* over;   // shared counters

  __device__ __forceinline__ uint32_t slot_of(int v) const { return ((uint32_t)v * 2654435761u) >> (32 - log2H); }
  // returns true if v was not a member yet (and appends it to L)
  __device__ __forceinline__ bool insert(int v) {
    uint32_t h = slot_of(v);
    while (true) {
      const uint32_t old = atomicCAS(&keys[h], SP_EMPTY, (uint32_t)v);
      if (old == SP_EMPTY) {
        const int idx = atomicAdd(nL, 1);
        if (idx < capL) L[idx] = (uint32_t)v; else *over = 1;
        return true;
      }
      if ((old & SP_MASK) == (uint32_t)v) return false;
      if (*reinterpret_cast<volatile int*>(over)) return false;
      h = (h + 1) & (H - 1);
    }
  }
  __device__ __forceinline__ int find(int v) const {  // slot or -1
    uint32_t h = slot_of(v);
    while (true) {
      const uint32_t k = keys[h];
      if (k == SP_EMPTY) return -1;
      if ((k & SP_MASK) == (uint32_t)v) return (int)h;
      h = (h + 1) & (H - 1);
    }
  }
  __device__ __forceinline__ bool reached(int v) const {
    const int s = find(v);
    return s >= 0 && (keys[s] & SP_FLAG);
  }
  __device__ __forceinline__ int local_id(int v, int n) const {  // index of v in the sorted R[0,n)
    int a = 0, b = n;
    while (a < b) {
      const int m = (a + b) >> 1;
      if ((int)R[m] < v) a = m + 1; else b = m;
    }
    return a;
  }
};

template <bool GLOBAL>
__global__ void __launch_bounds__(SP_THREADS) partition_sparse_kernel(const SparseArgs p) {
  extern __shared__ uint32_t sp_smem[];
  __shared__ int s_nL, s_nR, s_over, s_cnt, s_gid;
  SparseSet set;
  set.log2H = p.log2H; set.H = 1 << p.log2H; set.capL = p.capL; set.capR = p.capR;
  set.keys = GLOBAL ? p.gscratch + (size_t)blockIdx.x * ((size_t)set.H + p.capL + p.capR) : sp_smem;
  set.L = set.keys + set.H;
  set.R = set.L + p.capL;
  set.nL = &s_nL; set.nR = &s_nR; set.over = &s_over;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  constexpr int NW = SP_THREADS / 32;
  const int32_t* __restrict__ rowptr = p.rowptr;
  const int32_t* __restrict__ col = p.col;

  bool table_ready = false;  // tier 1 wipes its 2 MB table only if it really owns a centre

  for (int ci = blockIdx.x; ci < p.num_centres; ci += gridDim.x) {
    if (p.fill) {
      if (p.klass[ci] != p.tier || p.out_ne[ci] == 0) continue;
    } else if (p.tier != 0 && p.klass[ci] != p.tier) {
      continue;
    }
    if (!table_ready) {
      for (int i = tid; i < set.H; i += SP_THREADS) set.keys[i] = SP_EMPTY;
      table_ready = true;
      __syncthreads();
    }
    const int centre = p.centres[ci];
    const int limit = (p.mode == DESCO_MODE_KHOP) ? 0x7fffffff : centre;
    if (tid == 0) {
      s_nL = 0; s_nR = 0; s_over = 0; s_cnt = 0;
      int a = 0, b = p.num_graphs;  // largest a with graph_ptr[a] <= centre
      while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (p.graph_ptr[mid] <= centre) a = mid; else b = mid;
      }
      s_gid = a;
    }
    __syncthreads();
    if (tid == 0) set.insert(centre);
    __syncthreads();

    // ---- phase A: k levels of frontier expansion (data.py:329-350); one warp per frontier node ----
    int lb = 0, le = 1;
    bool over = false;
    for (int level = 0; level < p.depth && lb < le && !over; ++level) {
      const bool restricted = p.mode == DESCO_MODE_CANONICAL || (p.mode == DESCO_MODE_HETERO && level == p.depth - 1);
      for (int i = lb + warp; i < le; i += NW) {
        const int u = (int)set.L[i];
        const int rb = rowptr[u], re = rowptr[u + 1];
        for (int e0 = rb; e0 < re; e0 += 32) {
          const int e = e0 + lane;
          const int v = (e < re) ? col[e] : 0x7fffffff;
          const bool above = restricted && e < re && v > centre;
          if (e < re && !above) set.insert(v);
          if (__any_sync(FULL_MASK, above)) break;  // sorted row: nothing further passes
        }
      }
      __syncthreads();
      lb = le;
      le = min(s_nL, p.capL);
      over = s_over != 0;
      __syncthreads();
    }

    // ---- phase B + C: candidates <= centre (data.py:385), component of the centre inside them (:387-390) ----
    if (!over) {
      if (p.mode == DESCO_MODE_HETERO) {
        if (tid == 0) {
          const int s = set.find(centre);
          set.keys[s] |= SP_FLAG;
          set.R[0] = (uint32_t)centre;
          s_nR = 1;
        }
        __syncthreads();
        int qb = 0, qe = 1;
        while (qb < qe && !over) {
          for (int i = qb + warp; i < qe; i += NW) {
            const int u = (int)set.R[i];
            const int rb = rowptr[u], re = rowptr[u + 1];
            for (int e0 = rb; e0 < re; e0 += 32) {
              const int e = e0 + lane;
              const int v = (e < re) ? col[e] : 0x7fffffff;
              if (e < re && v <= centre) {
                const int s = set.find(v);
                if (s >= 0 && !(set.keys[s] & SP_FLAG)) {
                  const uint32_t old = atomicOr(&set.keys[s], SP_FLAG);
                  if (!(old & SP_FLAG)) {
                    const int idx = atomicAdd(&s_nR, 1);
                    if (idx < set.capR) set.R[idx] = (uint32_t)v; else s_over = 1;
                  }
                }
              }
              if (__any_sync(FULL_MASK, e < re && v > centre)) break;
            }
          }
          __syncthreads();
          qb = qe;
          qe = min(s_nR, set.capR);
          over = s_over != 0;
          __syncthreads();
        }
      } else {  // restricted BFS / plain k-hop ball: every member is in the result
        const int n = s_nL;
        if (n > set.capR) {
          if (tid == 0) s_over = 1;
        } else {
          for (int i = tid; i < n; i += SP_THREADS) {
            const int v = (int)set.L[i];
            set.R[i] = (uint32_t)v;
            set.keys[set.find(v)] |= SP_FLAG;
          }
          if (tid == 0) s_nR = n;
        }
        __syncthreads();
        over = s_over != 0;
      }
    }

    if (over) {
      // this tier cannot hold the ball: hand the centre to the next tier and wipe the table
      if (tid == 0 && !p.fill) {
        p.klass[ci] = (uint8_t)(p.tier + 1);
        p.out_nv[ci] = 0;
        p.out_ne[ci] = 0;
        p.centre_graph[ci] = s_gid;
      }
      __syncthreads();
      for (int i = tid; i < set.H; i += SP_THREADS) set.keys[i] = SP_EMPTY;
      __syncthreads();
      continue;
    }

    // ---- phase D: sort the reached list -> ascending node ids (canonical node = last row) ----
    const int nv = s_nR;
    int n2 = 1;
    while (n2 < nv) n2 <<= 1;
    for (int i = nv + tid; i < n2; i += SP_THREADS) set.R[i] = 0x7fffffffu;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < n2; i += SP_THREADS) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const uint32_t a = set.R[i], b = set.R[ixj];
            const bool up = (i & k) == 0;
            if ((a > b) == up) {
              set.R[i] = b;
              set.R[ixj] = a;
            }
          }
        }
        __syncthreads();
      }
    }

    // ---- induced degrees (count pass: just the total) ----
    const int n0 = p.fill ? p.node_off[ci] : 0;
    const int eo = p.fill ? p.edge_off[ci] : 0;
    for (int i = warp; i < nv; i += NW) {
      const int u = (int)set.R[i];
      const int rb = rowptr[u], re = rowptr[u + 1];
      int cnt = 0;
      for (int e0 = rb; e0 < re; e0 += 32) {
        const int e = e0 + lane;
        const int v = (e < re) ? col[e] : 0x7fffffff;
        if (e < re && v <= limit && set.reached(v)) ++cnt;
        if (__any_sync(FULL_MASK, e < re && v > limit)) break;
      }
      cnt = warp_sum(cnt);
      if (lane == 0) {
        if (p.fill) {
          p.node_gid[n0 + i] = u;
          p.edge_ptr[n0 + 1 + i] = cnt;
        } else {
          atomicAdd(&s_cnt, cnt);
        }
      }
    }
    __syncthreads();
    if (!p.fill) {
      if (tid == 0) {
        const int ne = s_cnt;
        p.out_ne[ci] = ne;
        p.out_nv[ci] = ne > 0 ? nv : 0;  // edge-free neighborhoods are dropped (workload.py:253-256)
        p.centre_graph[ci] = s_gid;
        if (p.tier == 0) p.klass[ci] = 0;
      }
    } else {
      if (n0 == 0 && tid == 0) p.edge_ptr[0] = 0;
      if (warp == 0) {  // in-place inclusive scan of the row degrees
        int carry = eo;
        for (int base = 0; base < nv; base += 32) {
          const int x = (base + lane < nv) ? p.edge_ptr[n0 + 1 + base + lane] : 0;
          const int incl = warp_incl_scan(x);
          if (base + lane < nv) p.edge_ptr[n0 + 1 + base + lane] = carry + incl;
          carry += __shfl_sync(FULL_MASK, incl, 31);
        }
      }
      __syncthreads();
      // edges in adjacency order (ascending node id == ascending local id); types follow in edge_types_kernel
      for (int i = warp; i < nv; i += NW) {
        const int u = (int)set.R[i];
        const int rb = rowptr[u], re = rowptr[u + 1];
        int out = (i == 0) ? eo : p.edge_ptr[n0 + i];
        for (int e0 = rb; e0 < re; e0 += 32) {
          const int e = e0 + lane;
          const int v = (e < re) ? col[e] : 0x7fffffff;
          const bool ok = e < re && v <= limit && set.reached(v);
          const uint32_t m = __ballot_sync(FULL_MASK, ok);
          if (ok) p.edge_col[out + __popc(m & ((1u << lane) - 1u))] = n0 + set.local_id(v, nv);
          out += __popc(m);
          if (__any_sync(FULL_MASK, e < re && v > limit)) break;
        }
      }
    }
    __syncthreads();

    // ---- wipe the table: tier 0 clears all slots, tier 1 only the members' ----
    if (!GLOBAL || s_nL > set.H / 16) {
      for (int i = tid; i < set.H; i += SP_THREADS) set.keys[i] = SP_EMPTY;
    } else {
      const int n = s_nL;
      for (int i = tid; i < n; i += SP_THREADS) set.L[i] = (uint32_t)set.find((int)set.L[i]);
      __syncthreads();
      for (int i = tid; i < n; i += SP_THREADS) set.keys[set.L[i]] = SP_EMPTY;
    }
    __syncthreads();
  }
}

// tunables of the large-graph path (desco_partition_large_set_caps lets the tests force every tier on small graphs)
int g_sp_log2h0 = 13, g_sp_capl0 = 5120, g_sp_capr0 = 4096;      // tier 0: 32 + 20 + 16 KB of shared memory
int g_sp_log2h1 = 19, g_sp_capl1 = 1 << 18, g_sp_capr1 = 1 << 18;  // tier 1: 4 MB of global scratch per CTA

struct LargeLayout {
  size_t klass_off, t1_off, t2_off, bytes;
  int t1_ctas, t2_ctas, max_words;
  size_t t1_words;
};

LargeLayout large_layout(int max_graph_nodes, int num_centres) {
  LargeLayout l;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const int sms = desco_num_sms();
  l.max_words = (max_graph_nodes + 31) / 32;
  l.t1_ctas = 4 * sms;  // tier 1 is latency bound on L2 hash probes: four 256-thread CTAs per SM
  l.t2_ctas = sms;
  l.t1_words = ((size_t)1 << g_sp_log2h1) + g_sp_capl1 + g_sp_capr1;
  l.klass_off = 0;
  l.t1_off = up((size_t)(num_centres > 0 ? num_centres : 1));
  l.t2_off = l.t1_off + up(l.t1_words * 4 * l.t1_ctas);
  l.bytes = l.t2_off + up((size_t)4 * l.max_words * 4 * l.t2_ctas);
  return l;
}

int launch_partition_large(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int num_graphs,
                           const int32_t* centres, int num_centres, int depth, int mode, int max_graph_nodes,
                           int32_t* nv, int32_t* ne, int32_t* centre_graph, int fill, const int32_t* node_off,
                           const int32_t* edge_off, int32_t* node_gid, int32_t* edge_ptr, int32_t* edge_col,
                           uint8_t* edge_tri, int32_t* status, void* workspace, int64_t workspace_bytes,
                           cudaStream_t stream) {
  if (depth < 0 || mode < DESCO_MODE_HETERO || mode > DESCO_MODE_KHOP || max_graph_nodes <= 0 || num_centres < 0)
    return DESCO_EINVAL;
  if (num_centres == 0) return DESCO_OK;
  if (!rowptr || !col || !graph_ptr || !centres || !nv || !ne || !centre_graph || !status || !workspace) return DESCO_EINVAL;
  const LargeLayout l = large_layout(max_graph_nodes, num_centres);
  if ((int64_t)l.bytes > workspace_bytes) return DESCO_ENOMEM;
  uint8_t* base = (uint8_t*)workspace;
  const int sms = desco_num_sms();
  SparseArgs a;
  a.rowptr = rowptr; a.col = col; a.graph_ptr = graph_ptr; a.num_graphs = num_graphs;
  a.centres = centres; a.num_centres = num_centres; a.depth = depth; a.mode = mode;
  a.out_nv = nv; a.out_ne = ne; a.centre_graph = centre_graph;
  a.fill = fill; a.node_off = node_off; a.edge_off = edge_off; a.node_gid = node_gid; a.edge_ptr = edge_ptr; a.edge_col = edge_col;
  a.klass = base + l.klass_off;
  DescoProfScope prof(DESCO_PROF_PARTITION, stream, 3);
  {  // tier 0: shared memory
    a.tier = 0; a.gscratch = nullptr; a.log2H = g_sp_log2h0; a.capL = g_sp_capl0; a.capR = g_sp_capr0;
    const size_t smem = (((size_t)1 << a.log2H) + a.capL + a.capR) * 4;
    if (smem > 200 * 1024) return DESCO_EINVAL;
    DESCO_CUDA_TRY(cudaFuncSetAttribute(partition_sparse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const int blocks = num_centres < sms * per_sm ? num_centres : sms * per_sm;
    partition_sparse_kernel<false><<<blocks, SP_THREADS, smem, stream>>>(a);
    DESCO_LAUNCH_CHECK();
  }
  {  // tier 1: global scratch
    a.tier = 1; a.gscratch = (uint32_t*)(base + l.t1_off); a.log2H = g_sp_log2h1; a.capL = g_sp_capl1; a.capR = g_sp_capr1;
    const int blocks = num_centres < l.t1_ctas ? num_centres : l.t1_ctas;
    partition_sparse_kernel<true><<<blocks, SP_THREADS, 0, stream>>>(a);
    DESCO_LAUNCH_CHECK();
  }
  {  // dense tier: bitsets over the whole target graph in global scratch
    const int blocks = num_centres < l.t2_ctas ? num_centres : l.t2_ctas;
    partition_kernel<true><<<blocks, 256, 0, stream>>>(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode,
                                                      l.max_words, nv, ne, centre_graph, fill, node_off, edge_off, node_gid,
                                                      edge_ptr, edge_col, edge_tri, status,
                                                      (uint32_t*)(base + l.t2_off), a.klass, 2);
    DESCO_LAUNCH_CHECK();
  }
  return DESCO_OK;
}

struct KeepFlag {
  __host__ __device__ int operator()(const int32_t& ne) const { return ne > 0 ? 1 : 0; }
};

__global__ void scan_finalize_kernel(const int32_t* __restrict__ centres, const int32_t* __restrict__ nv,
                                     const int32_t* __restrict__ ne, const int32_t* __restrict__ rank,
                                     const int32_t* __restrict__ node_off, const int32_t* __restrict__ edge_off, int C,
                                     int32_t* __restrict__ nbh_ptr, int32_t* __restrict__ centre_out,
                                     uint8_t* __restrict__ indicator, int32_t* __restrict__ totals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C) return;
  bool keep = ne[i] > 0;
  if (indicator) indicator[i] = keep ? 1 : 0;
  if (keep) {
    nbh_ptr[rank[i]] = node_off[i];
    centre_out[rank[i]] = centres[i];
  }
  if (i == C - 1) {
    int G = rank[i] + (keep ? 1 : 0), V = node_off[i] + nv[i], E = edge_off[i] + ne[i];
    nbh_ptr[G] = V;
    totals[0] = G;
    totals[1] = V;
    totals[2] = E;
  }
}

int launch_partition(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int num_graphs,
                     const int32_t* centres, int num_centres, int depth, int mode, int max_graph_nodes, int32_t* nv,
                     int32_t* ne, int32_t* centre_graph, int fill, const int32_t* node_off, const int32_t* edge_off,
                     int32_t* node_gid, int32_t* edge_ptr, int32_t* edge_col, uint8_t* edge_tri, int32_t* status,
                     cudaStream_t stream) {
  if (depth < 0 || mode < DESCO_MODE_HETERO || mode > DESCO_MODE_KHOP || max_graph_nodes <= 0 || num_centres < 0)
    return DESCO_EINVAL;
  if (num_centres == 0) return DESCO_OK;
  if (!rowptr || !col || !graph_ptr || !centres || !nv || !ne || !centre_graph || !status) return DESCO_EINVAL;
  const int max_words = (max_graph_nodes + 31) / 32;
  const int sms = desco_num_sms();
  DescoProfScope prof(DESCO_PROF_PARTITION, stream);
#define DESCO_SMALL(Wc)                                                                                                  \
  return launch_partition_small<Wc>(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, nv, ne,       \
                                    centre_graph, fill, node_off, edge_off, node_gid, edge_ptr, edge_col, edge_tri,      \
                                    status, stream)
  if (max_words <= 1) DESCO_SMALL(1);
  if (max_words <= 2) DESCO_SMALL(2);
  if (max_words <= 4) DESCO_SMALL(4);
  if (max_words <= 8) DESCO_SMALL(8);
#undef DESCO_SMALL
  if (max_words <= 64) {  // warp per centre
    const int threads = 256, groups = threads / 32;
    size_t smem = (size_t)groups * 4 * max_words * sizeof(uint32_t);
    int blocks = (num_centres + groups - 1) / groups;
    int cap = sms * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > cap) blocks = cap;
    partition_kernel<false><<<blocks, threads, smem, stream>>>(rowptr, col, graph_ptr, num_graphs, centres, num_centres,
                                                              depth, mode, max_words, nv, ne, centre_graph, fill, node_off,
                                                              edge_off, node_gid, edge_ptr, edge_col, edge_tri, status);
  } else {
    size_t smem = (size_t)4 * max_words * sizeof(uint32_t);
    if (smem > 200 * 1024) return DESCO_ERANGE;  // large-graph regime: use desco_partition_large_* (hash-set frontier)
    DESCO_CUDA_TRY(cudaFuncSetAttribute(partition_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = 256;
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int blocks = num_centres < sms * per_sm ? num_centres : sms * per_sm;
    partition_kernel<true><<<blocks, threads, smem, stream>>>(rowptr, col, graph_ptr, num_graphs, centres, num_centres,
                                                             depth, mode, max_words, nv, ne, centre_graph, fill, node_off,
                                                             edge_off, node_gid, edge_ptr, edge_col, edge_tri, status);
  }
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // namespace

extern "C" {

int desco_partition_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                          const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                          int32_t max_graph_nodes, int32_t* out_nv, int32_t* out_ne, int32_t* out_centre_graph,
                          int32_t* status, void* stream) {
  return launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes, out_nv,
                          out_ne, out_centre_graph, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, status,
                          (cudaStream_t)stream);
}

int64_t desco_partition_scan_workspace_bytes(int32_t num_centres) {
  size_t a = 0, b = 0;
  cub::TransformInputIterator<int, KeepFlag, const int32_t*> it((const int32_t*)nullptr, KeepFlag());
  cub::DeviceScan::ExclusiveSum(nullptr, a, it, (int32_t*)nullptr, num_centres);
  cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t*)nullptr, (int32_t*)nullptr, num_centres);
  return (int64_t)(a > b ? a : b) + 256;
}

int desco_partition_scan(const int32_t* centres, const int32_t* nv, const int32_t* ne, int32_t num_centres,
                         int32_t* keep_rank, int32_t* node_off, int32_t* edge_off, int32_t* nbh_ptr,
                         int32_t* centre_out, uint8_t* indicator, int32_t* totals, void* workspace,
                         int64_t workspace_bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!nbh_ptr || !totals || num_centres < 0) return DESCO_EINVAL;
  if (num_centres > 0 && (!centres || !nv || !ne || !keep_rank || !node_off || !edge_off || !centre_out || !workspace))
    return DESCO_EINVAL;
  if (num_centres == 0) {
    DESCO_CUDA_TRY(cudaMemsetAsync(totals, 0, 3 * sizeof(int32_t), s));
    DESCO_CUDA_TRY(cudaMemsetAsync(nbh_ptr, 0, sizeof(int32_t), s));
    return DESCO_OK;
  }
  size_t bytes = (size_t)workspace_bytes;
  cub::TransformInputIterator<int, KeepFlag, const int32_t*> it(ne, KeepFlag());
  DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(workspace, bytes, it, keep_rank, num_centres, s));
  bytes = (size_t)workspace_bytes;
  DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(workspace, bytes, nv, node_off, num_centres, s));
  bytes = (size_t)workspace_bytes;
  DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(workspace, bytes, ne, edge_off, num_centres, s));
  desco_count_launches(1);
  scan_finalize_kernel<<<(num_centres + 255) / 256, 256, 0, s>>>(centres, nv, ne, keep_rank, node_off, edge_off,
                                                                num_centres, nbh_ptr, centre_out, indicator, totals);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_partition_fill(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                         const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                         int32_t max_graph_nodes, const int32_t* nv, const int32_t* ne, const int32_t* centre_graph,
                         const int32_t* node_off, const int32_t* edge_off, int32_t* node_gid, int32_t* edge_ptr,
                         int32_t* edge_col, uint8_t* edge_tri, int32_t* status, void* stream) {
  if (!node_off || !edge_off || !node_gid || !edge_ptr || !edge_col || !edge_tri) return DESCO_EINVAL;
  return launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes,
                          const_cast<int32_t*>(nv), const_cast<int32_t*>(ne), const_cast<int32_t*>(centre_graph), 1,
                          node_off, edge_off, node_gid, edge_ptr, edge_col, edge_tri, status, (cudaStream_t)stream);
}

int desco_shmp_edge_types(const int32_t* edge_ptr, const int32_t* edge_col, int32_t num_rows, uint8_t* edge_tri,
                          void* stream) {
  if (!edge_ptr || !edge_col || !edge_tri) return DESCO_EINVAL;
  if (num_rows == 0) return DESCO_OK;
  const int threads = 256;
  long long total = (long long)num_rows * 32;
  DescoProfScope prof(DESCO_PROF_PARTITION, (cudaStream_t)stream);
  edge_types_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(edge_ptr, edge_col,
                                                                                                       num_rows, edge_tri);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int64_t desco_partition_large_workspace_bytes(int32_t max_graph_nodes, int32_t num_centres) {
  return (int64_t)large_layout(max_graph_nodes, num_centres).bytes;
}

int desco_partition_large_set_caps(int32_t log2_slots0, int32_t members0, int32_t reached0, int32_t log2_slots1,
                                   int32_t members1, int32_t reached1) {
  auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
  if (log2_slots0 < 4 || log2_slots0 > 15 || log2_slots1 < 4 || log2_slots1 > 26 || !pow2(reached0) || !pow2(reached1) ||
      members0 < 1 || members1 < 1 || members0 + SP_THREADS > (1 << log2_slots0) || members1 + SP_THREADS > (1 << log2_slots1))
    return DESCO_EINVAL;
  g_sp_log2h0 = log2_slots0; g_sp_capl0 = members0; g_sp_capr0 = reached0;
  g_sp_log2h1 = log2_slots1; g_sp_capl1 = members1; g_sp_capr1 = reached1;
  return DESCO_OK;
}

int desco_partition_large_count(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                                const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                                int32_t max_graph_nodes, int32_t* out_nv, int32_t* out_ne, int32_t* out_centre_graph,
                                int32_t* status, void* workspace, int64_t workspace_bytes, void* stream) {
  return launch_partition_large(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes,
                                out_nv, out_ne, out_centre_graph, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                status, workspace, workspace_bytes, (cudaStream_t)stream);
}

int desco_partition_large_fill(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                               const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                               int32_t max_graph_nodes, const int32_t* nv, const int32_t* ne, const int32_t* centre_graph,
                               const int32_t* node_off, const int32_t* edge_off, int32_t num_rows, int32_t* node_gid,
                               int32_t* edge_ptr, int32_t* edge_col, uint8_t* edge_tri, int32_t* status, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  if (!node_off || !edge_off || !node_gid || !edge_ptr || !edge_col || !edge_tri || num_rows < 0) return DESCO_EINVAL;
  const int rc = launch_partition_large(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode,
                                        max_graph_nodes, const_cast<int32_t*>(nv), const_cast<int32_t*>(ne),
                                        const_cast<int32_t*>(centre_graph), 1, node_off, edge_off, node_gid, edge_ptr,
                                        edge_col, edge_tri, status, workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc) return rc;
  return desco_shmp_edge_types(edge_ptr, edge_col, num_rows, edge_tri, stream);  // SHMP types of the whole packed batch
}

const char* desco_version(void) { return "desco_b200 0.1 sm_100a"; }

}  // extern "C"
