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
    uint8_t* __restrict__ edge_tri, int32_t* __restrict__ status) {
  extern __shared__ uint32_t smem[];
  const int groups_per_cta = CTA ? 1 : (blockDim.x >> 5);
  const int g_in_cta = CTA ? 0 : warp_id();
  uint32_t* S = smem + (size_t)g_in_cta * 4 * max_words;
  uint32_t* F = S + max_words;
  uint32_t* Nx = F + max_words;
  uint32_t* aux = Nx + max_words;

  for (int ci = blockIdx.x * groups_per_cta + g_in_cta; ci < num_centres; ci += gridDim.x * groups_per_cta) {
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
          edge_tri[out] = has_common_neighbour(rowptr, col, S, lo, limit, u, v) ? 1 : 0;
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

// ToTconvHetero on an existing packed batch: one warp per row, one lane per incident edge.  The type of an edge is
// symmetric (a common neighbour of u and v), so only the u < v direction intersects the two rows; it then writes the
// reverse entry too (its position in row v by binary search - rows are sorted and the batch is symmetric).  When BOTH
// rows are long (hub-hub edges of a power-law ball: a lane would walk thousands of dependent loads) the whole warp
// takes the edge: 32 elements of the shorter row at a time, each lane one binary search in the longer row.
constexpr int TYPES_COOP_MIN = 48;  // both rows at least this long -> warp-cooperative intersection

__global__ void edge_types_kernel(const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ edge_col,
                                  int num_rows, uint8_t* __restrict__ edge_tri) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= num_rows) return;
  const int lane = lane_id();
  const int rb = edge_ptr[row], re = edge_ptr[row + 1];
  for (int e0 = rb; e0 < re; e0 += 32) {
    const int e = e0 + lane;
    int v = -1, vb = 0, ve = 0, rev = -1;
    bool mine = false;  // this lane owns the undirected edge (row, v)
    if (e < re) {
      v = edge_col[e];
      vb = edge_ptr[v];
      ve = edge_ptr[v + 1];
      int lo = vb, hi = ve;  // position of `row` in row v
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (edge_col[mid] < row) lo = mid + 1; else hi = mid;
      }
      const bool mirrored = lo < ve && edge_col[lo] == row;  // always, for the undirected batches of the partition
      mine = !(mirrored && v < row);                         // otherwise written by row v
      if (mirrored && v > row) rev = lo;
    }
    const bool coop = mine && (re - rb) >= TYPES_COOP_MIN && (ve - vb) >= TYPES_COOP_MIN;
    if (mine && !coop) {
      const uint8_t t = sorted_lists_intersect(edge_col, rb, re, vb, ve) ? 1 : 0;
      edge_tri[e] = t;
      if (rev >= 0) edge_tri[rev] = t;
    }
    uint32_t todo = __ballot_sync(FULL_MASK, coop);
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1;
      int sb = rb, se = re;  // shorter list
      int lb = __shfl_sync(FULL_MASK, vb, src), le = __shfl_sync(FULL_MASK, ve, src);  // longer list
      if (se - sb > le - lb) {
        int t0 = sb; sb = lb; lb = t0;
        t0 = se; se = le; le = t0;
      }
      bool hit = false;
      for (int k0 = sb; k0 < se && !hit; k0 += 32) {
        bool found = false;
        if (k0 + lane < se) {
          const int x = edge_col[k0 + lane];
          int lo = lb, hi = le;
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (edge_col[mid] < x) lo = mid + 1; else hi = mid;
          }
          found = lo < le && edge_col[lo] == x;
        }
        hit = __any_sync(FULL_MASK, found) != 0;
      }
      if (lane == src) {
        const uint8_t t = hit ? 1 : 0;
        edge_tri[e] = t;
        if (rev >= 0) edge_tri[rev] = t;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Large-graph regime (config 5: one 10M-node / 200M-directed-edge power-law target).  A bitset over the target no
// longer fits shared memory and would cost O(N/32) per centre, so the set state of an ordinary centre is SPARSE and
// lives in shared memory (tier 0):
//   * an open-addressing hash set of node ids (linear probing, atomicCAS insert; bit 31 of a key = "reached" flag),
//   * the member list L in discovery order (BFS levels are contiguous slices of it),
//   * the reached list R (component of the centre), bitonic-sorted at the end so that rows come out in ascending node
//     id; the local id of a node is then parked beside its hash slot (uint16 table aliasing the dead member list).
// A centre whose ball overflows these tables is appended to the big-centre list and served by the team tier below.
// The sorted adjacency makes "<= centre" a PREFIX of every row, so the restricted passes stop at the first neighbour
// above the centre.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t SP_EMPTY = 0xffffffffu;
constexpr uint32_t SP_FLAG = 0x80000000u;
constexpr uint32_t SP_MASK = 0x7fffffffu;
constexpr int SP_THREADS = 1024;

// Streams col[rb, re) of one adjacency row through a warp, 32 entries per call of f(v, valid), with PF independent
// 128-byte loads in flight: every chunk of a cold row is its own ~1 us L2 / HBM access and the early exit of a sorted row
// (f returns true, warp-uniformly) would otherwise make them a dependent chain.
template <int PF, typename Fn>
__device__ __forceinline__ void stream_row(const int32_t* __restrict__ col, int rb, int re, Fn f) {
  const int lane = lane_id();
  for (int e0 = rb; e0 < re; e0 += 32 * PF) {
    int v[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      const int e = e0 + 32 * k + lane;
      v[k] = (e < re) ? col[e] : 0x7fffffff;
    }
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      if (e0 + 32 * k >= re) return;
      if (f(v[k], e0 + 32 * k + lane < re)) return;
    }
  }
}
constexpr int ROW_PF = 4;

struct SparseArgs {
  const int32_t* rowptr; const int32_t* col; const int32_t* graph_ptr; int num_graphs;
  const int32_t* centres; int num_centres, depth, mode;
  int32_t* out_nv; int32_t* out_ne; int32_t* centre_graph;
  int fill; const int32_t* node_off; const int32_t* edge_off;
  int32_t* node_gid; int32_t* edge_ptr; int32_t* edge_col;
  uint8_t* klass;                // centre class: 0 = served by this kernel (shared memory), 1 = handed to the team tier
  int32_t* big_list; int* big_count;  // centres handed to the team tier (appended by the count pass)
  // sorted reached list + induced degrees of every served centre, kept from the count pass (hetero mode) so that fill
  // goes straight to the emission; a centre that does not fit the cache (cache_off = -1) is recomputed
  unsigned long long* cache_cursor; int32_t* cache_off; int32_t* cache; long long cache_cap;
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
  // v is known to be absent: park it with the "reached" flag set, return its slot
  __device__ __forceinline__ int insert_reached(int v) {
    uint32_t h = slot_of(v);
    while (atomicCAS(&keys[h], SP_EMPTY, (uint32_t)v | SP_FLAG) != SP_EMPTY) h = (h + 1) & (H - 1);
    return (int)h;
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
};

// phase timing of the shared-memory tier (thread 0 of every CTA adds its clock64 deltas; desco_partition_large_phase_cycles)
enum { SPH_EXPAND = 0, SPH_COMPONENT, SPH_SORT, SPH_DEGREE, SPH_EMIT, SPH_WIPE, SPH_COUNT };
__device__ unsigned long long g_sparse_phase_cycles[SPH_COUNT];

__global__ void __launch_bounds__(SP_THREADS) partition_sparse_kernel(const SparseArgs p) {
  extern __shared__ uint32_t sp_smem[];
  __shared__ int s_nL, s_nR, s_over, s_cnt, s_gid, s_coff;
  SparseSet set;
  set.log2H = p.log2H; set.H = 1 << p.log2H; set.capL = p.capL; set.capR = p.capR;
  set.keys = sp_smem;
  set.L = set.keys + set.H;
  set.R = set.L + max(p.capL, set.H / 2);
  // local id of the node in hash slot s, written after the sort; aliases the member list (dead by then)
  uint16_t* rank16 = reinterpret_cast<uint16_t*>(set.L);
  set.nL = &s_nL; set.nR = &s_nR; set.over = &s_over;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  constexpr int NW = SP_THREADS / 32;
  const int32_t* __restrict__ rowptr = p.rowptr;
  const int32_t* __restrict__ col = p.col;

  bool table_ready = false;

  for (int ci = blockIdx.x; ci < p.num_centres; ci += gridDim.x) {
    if (p.fill && (p.klass[ci] != 0 || p.out_ne[ci] == 0)) continue;
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
    const int n0 = p.fill ? p.node_off[ci] : 0;
    const int eo = p.fill ? p.edge_off[ci] : 0;
    const int coff = p.fill ? p.cache_off[ci] : -1;
    int nv = 0;
    long long tick = clock64();
    auto lap = [&](int phase) {
      if (tid == 0) {
        const long long now = clock64();
        atomicAdd(&g_sparse_phase_cycles[phase], (unsigned long long)(now - tick));
        tick = now;
      }
    };
    if (coff >= 0) {  // fill from the cache: rebuild the hash from the sorted reached list, rows and degrees are at hand
      nv = p.out_nv[ci];
      for (int i = tid; i < nv; i += SP_THREADS) {
        const int v = p.cache[coff + i];
        set.R[i] = (uint32_t)v;
        rank16[set.insert_reached(v)] = (uint16_t)i;
        p.node_gid[n0 + i] = v;
        p.edge_ptr[n0 + 1 + i] = p.cache[coff + nv + i];
      }
      __syncthreads();
    } else {
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
        stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
          const bool above = restricted && valid && v > centre;
          if (valid && !above) set.insert(v);
          return __any_sync(FULL_MASK, above) != 0;  // sorted row: nothing further passes
        });
      }
      __syncthreads();
      lb = le;
      le = min(s_nL, p.capL);
      over = s_over != 0;
      __syncthreads();
    }

    lap(SPH_EXPAND);
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
            int cnt = 0;  // every member <= centre next to a reached node is reached too: this IS the induced degree
            stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
              if (valid && v <= centre) {
                const int s = set.find(v);
                if (s >= 0) {
                  ++cnt;
                  if (!(set.keys[s] & SP_FLAG)) {
                    const uint32_t old = atomicOr(&set.keys[s], SP_FLAG);
                    if (!(old & SP_FLAG)) {
                      const int idx = atomicAdd(&s_nR, 1);
                      if (idx < set.capR) set.R[idx] = (uint32_t)v; else s_over = 1;
                    }
                  }
                }
              }
              return __any_sync(FULL_MASK, valid && v > centre) != 0;
            });
            cnt = warp_sum(cnt);
            if (lane == 0) {
              rank16[set.find(u)] = (uint16_t)min(cnt, 0xffff);  // parked beside the hash slot until the sort
              atomicAdd(&s_cnt, cnt);
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
      // shared memory cannot hold the ball: hand the centre to the team tier and wipe the table
      if (tid == 0 && !p.fill) {
        p.klass[ci] = 1;
        p.out_nv[ci] = 0;
        p.out_ne[ci] = 0;
        p.centre_graph[ci] = s_gid;
        p.big_list[atomicAdd(p.big_count, 1)] = ci;
      }
      __syncthreads();
      for (int i = tid; i < set.H; i += SP_THREADS) set.keys[i] = SP_EMPTY;
      __syncthreads();
      continue;
    }

    lap(SPH_COMPONENT);
    // ---- phase D: sort the reached list -> ascending node ids (canonical node = last row) ----
    nv = s_nR;
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

    lap(SPH_SORT);
    // ---- induced degrees (count pass: just the total) ----
    if (p.mode == DESCO_MODE_HETERO) {  // degrees were counted by the component BFS
      if (p.fill)
        for (int i = tid; i < nv; i += SP_THREADS) {
          const int u = (int)set.R[i];
          p.node_gid[n0 + i] = u;
          p.edge_ptr[n0 + 1 + i] = (int)rank16[set.find(u)];
        }
    } else
    for (int i = warp; i < nv; i += NW) {
      const int u = (int)set.R[i];
      const int rb = rowptr[u], re = rowptr[u + 1];
      int cnt = 0;
      stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
        if (valid && v <= limit && set.reached(v)) ++cnt;
        return __any_sync(FULL_MASK, valid && v > limit) != 0;
      });
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
    }  // not cached
    __syncthreads();
    lap(SPH_DEGREE);
    if (!p.fill) {
      if (tid == 0) {
        const int ne = s_cnt;
        p.out_ne[ci] = ne;
        p.out_nv[ci] = ne > 0 ? nv : 0;  // edge-free neighborhoods are dropped (workload.py:253-256)
        p.centre_graph[ci] = s_gid;
        p.klass[ci] = 0;
        long long off = -1;
        if (ne > 0 && p.mode == DESCO_MODE_HETERO) {
          off = (long long)atomicAdd(p.cache_cursor, (unsigned long long)(2 * nv));
          if (off + 2 * nv > p.cache_cap) off = -1;
        }
        p.cache_off[ci] = (int32_t)off;
        s_coff = (int)off;
      }
      __syncthreads();
      if (s_coff >= 0)
        for (int i = tid; i < nv; i += SP_THREADS) {
          const int u = (int)set.R[i];
          p.cache[s_coff + i] = u;
          p.cache[s_coff + nv + i] = (int)rank16[set.find(u)];  // induced degree, still parked beside the slot
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
      for (int i = tid; i < nv; i += SP_THREADS) rank16[set.find((int)set.R[i])] = (uint16_t)i;
      __syncthreads();
      for (int i = warp; i < nv; i += NW) {
        const int u = (int)set.R[i];
        const int rb = rowptr[u], re = rowptr[u + 1];
        int out = (i == 0) ? eo : p.edge_ptr[n0 + i];
        stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
          const int slot = (valid && v <= limit) ? set.find(v) : -1;
          const bool ok = slot >= 0 && (set.keys[slot] & SP_FLAG);
          const uint32_t m = __ballot_sync(FULL_MASK, ok);
          if (ok) p.edge_col[out + __popc(m & ((1u << lane) - 1u))] = n0 + (int)rank16[slot];
          out += __popc(m);
          return __any_sync(FULL_MASK, valid && v > limit) != 0;
        });
      }
    }
    __syncthreads();
    lap(SPH_EMIT);

    // ---- wipe the table ----
    for (int i = tid; i < set.H; i += SP_THREADS) set.keys[i] = SP_EMPTY;
    __syncthreads();
    lap(SPH_WIPE);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Team tier: balls that overflow the shared-memory tier (a few per cent of the centres of a power-law target, but most
// of its rows: a hub's depth-2 ball has 10^5..10^6 rows).  One CTA per such centre leaves ~256 L2 requests in flight for
// millions of adjacency probes, so a TEAM of TM_CTAS co-resident CTAs (cooperative launch) owns the centre instead:
//   * member set and reached set are BITMAPS over the centre's target graph in a per-team slice of global scratch
//     (L2-resident: 1.25 MB per set for 10M nodes); a reached word is a {bits, prefix} pair, so "is v in the
//     neighborhood" and "local id of v" are ONE 8-byte load, and ascending-node-id order needs no sort;
//   * frontier / reached nodes are appended to lists (warp-aggregated atomics), BFS levels are slices of the list;
//   * phases are separated by a team barrier (arrive counter + generation in global memory); the last arriver
//     snapshots the list lengths, so every CTA leaves the barrier with the same view;
//   * adjacency rows are streamed a warp per row, or - while a level has fewer rows than the team has warps (level 0 is
//     the centre alone, possibly a hub) - 32-entry chunks of each row dealt over all warps of the team.
// Big centres are taken from a list the tier-0 kernel appends to, through an atomic ticket (dynamic balancing).
// ------------------------------------------------------------------------------------------------------------------
constexpr int TM_THREADS = 256;
constexpr int TM_CTAS = 16;
constexpr int TM_WARPS = TM_THREADS / 32;

struct TeamCtrl {
  unsigned bar_count, bar_gen;
  int cur[2];          // ticket (index into big_list) of the centre of this / the next iteration (by parity)
  int nL, nR, cnt;     // member-list length, reached-list length, induced directed edges
  int abort;
  int snap[2][4];      // barrier snapshot (nL, nR, cnt, abort) by generation parity
  int chunk_tot[TM_CTAS];
  int pad[4];
};
static_assert(sizeof(TeamCtrl) % 16 == 0, "TeamCtrl must keep the slices aligned");

struct TeamArgs {
  const int32_t* rowptr; const int32_t* col; const int32_t* graph_ptr; int num_graphs;
  const int32_t* centres; int num_centres, depth, mode;
  int32_t* out_nv; int32_t* out_ne; int32_t* centre_graph;
  int fill; const int32_t* node_off; const int32_t* edge_off;
  int32_t* node_gid; int32_t* edge_ptr; int32_t* edge_col;
  const int32_t* big_list; int* big_count; int* ticket;
  TeamCtrl* ctrl; uint8_t* slices; size_t slice_bytes;
  int max_words, max_nodes;
  // reached lists of the big centres, kept from the count pass so that fill skips both BFS phases (two of its four
  // adjacency passes); a centre that does not fit (cache_off = -1) is simply recomputed
  unsigned long long* cache_cursor; int32_t* cache_off; int32_t* cache; long long cache_cap;
  int32_t* status;
};

struct TeamSnap { int nL, nR, cnt, abort; };

__device__ __forceinline__ TeamSnap team_sync(TeamCtrl* c, int* s_snap) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = *reinterpret_cast<volatile unsigned*>(&c->bar_gen);
    const int par = (int)(gen & 1u);
    if (atomicAdd(&c->bar_count, 1u) == (unsigned)(TM_CTAS - 1)) {
      c->snap[par][0] = __ldcg(&c->nL);
      c->snap[par][1] = __ldcg(&c->nR);
      c->snap[par][2] = __ldcg(&c->cnt);
      c->snap[par][3] = __ldcg(&c->abort);
      *reinterpret_cast<volatile unsigned*>(&c->bar_count) = 0u;
      __threadfence();
      atomicAdd(&c->bar_gen, 1u);
    } else {
      bool ok = false;  // bounded: a lost team-mate must never hang the GPU box
      for (unsigned spin = 0; spin < (1u << 25); ++spin) {
        if (*reinterpret_cast<volatile unsigned*>(&c->bar_gen) != gen) { ok = true; break; }
        __nanosleep(40);
      }
      if (!ok) atomicExch(&c->abort, 1);
    }
    __threadfence();
    s_snap[0] = __ldcg(&c->snap[par][0]);
    s_snap[1] = __ldcg(&c->snap[par][1]);
    s_snap[2] = __ldcg(&c->snap[par][2]);
    s_snap[3] = __ldcg(&c->snap[par][3]) | __ldcg(&c->abort);
  }
  __syncthreads();
  TeamSnap r;
  r.nL = s_snap[0]; r.nR = s_snap[1]; r.cnt = s_snap[2]; r.abort = s_snap[3];
  return r;
}

// f(v, ok) is called warp-uniformly for 32 adjacency entries at a time of every node of list[lb, le); ok = the entry
// exists and is <= limit (rows are sorted, so a warp stops a row at its first failing chunk).
template <typename Fn>
__device__ __forceinline__ void team_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                          const int32_t* list, int lb, int le, int limit, int rank, Fn f) {
  const int lane = lane_id();
  const int gw = rank * TM_WARPS + warp_id();
  constexpr int TW = TM_CTAS * TM_WARPS;
  if (le - lb >= TW) {
    for (int i = lb + gw; i < le; i += TW) {
      const int u = __ldcg(list + i);
      const int rb = rowptr[u], re = rowptr[u + 1];
      stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
        const bool ok = valid && v <= limit;
        f(v, ok);
        return !__all_sync(FULL_MASK, ok);
      });
    }
  } else {
    for (int i = lb; i < le; ++i) {
      const int u = __ldcg(list + i);
      const int rb = rowptr[u], re = rowptr[u + 1];
      for (int e0 = rb + gw * 32; e0 < re; e0 += TW * 32) {
        const int e = e0 + lane;
        const int v = (e < re) ? col[e] : 0;
        const bool ok = e < re && v <= limit;
        f(v, ok);
        if (!__all_sync(FULL_MASK, ok)) break;
      }
    }
  }
}

// warp-aggregated append of the lanes with `take` set
__device__ __forceinline__ void team_append(int32_t* list, int* counter, int cap, int v, bool take) {
  const uint32_t m = __ballot_sync(FULL_MASK, take);
  if (!m) return;
  int base = 0;
  if (lane_id() == 0) base = atomicAdd(counter, __popc(m));
  base = __shfl_sync(FULL_MASK, base, 0);
  const int idx = base + __popc(m & ((1u << lane_id()) - 1u));
  if (take && idx < cap) list[idx] = v;
}

__global__ void __launch_bounds__(TM_THREADS) partition_team_kernel(const TeamArgs p) {
  __shared__ int s_snap[4];
  __shared__ int s_wtot[TM_WARPS];
  __shared__ int s_gid;
  const int team = blockIdx.x / TM_CTAS, rank = blockIdx.x % TM_CTAS;
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  constexpr int TT = TM_CTAS * TM_THREADS;
  const int tt = rank * TM_THREADS + tid;
  TeamCtrl* c = p.ctrl + team;
  uint8_t* slice = p.slices + (size_t)team * p.slice_bytes;
  uint2* RP = reinterpret_cast<uint2*>(slice);                                 // [max_words] {reached bits, prefix}
  uint32_t* Mb = reinterpret_cast<uint32_t*>(slice + (size_t)p.max_words * 8); // [max_words] member bits
  int32_t* L = reinterpret_cast<int32_t*>(slice + (size_t)p.max_words * 12);   // [max_nodes] members, discovery order
  int32_t* R = L + p.max_nodes;                                                // [max_nodes] reached, discovery order
  const int32_t* __restrict__ rowptr = p.rowptr;
  const int32_t* __restrict__ col = p.col;
  const int nbig = __ldcg(p.big_count);

  if (rank == 0 && tid == 0) c->cur[0] = atomicAdd(p.ticket, 1);
  TeamSnap sn = team_sync(c, s_snap);
  for (int iter = 0; !sn.abort; ++iter) {
    // two slots: an iteration without inner barriers (dropped centre) must not overwrite the ticket the slower
    // team-mates are still about to read
    const int item = __ldcg(&c->cur[iter & 1]);
    if (item >= nbig) break;
    const int ci = __ldcg(p.big_list + item);
    const int centre = p.centres[ci];
    const bool skip = p.fill && p.out_ne[ci] == 0;  // dropped neighborhood: nothing to emit
    if (tid == 0) {
      int a = 0, b = p.num_graphs;  // largest a with graph_ptr[a] <= centre
      while (b - a > 1) {
        const int mid = (a + b) >> 1;
        if (p.graph_ptr[mid] <= centre) a = mid; else b = mid;
      }
      s_gid = a;
    }
    __syncthreads();
    const int gid = s_gid;
    const int lo = p.graph_ptr[gid];
    const int W = (p.graph_ptr[gid + 1] - lo + 31) >> 5;
    const int limit = (p.mode == DESCO_MODE_KHOP) ? 0x7fffffff : centre;
    const int cl = centre - lo;
    if (W > p.max_words) {
      if (rank == 0 && tid == 0) {
        atomicExch(p.status, DESCO_ERANGE);
        if (!p.fill) { p.out_nv[ci] = 0; p.out_ne[ci] = 0; p.centre_graph[ci] = gid; }
      }
    } else if (!skip) {
      const int coff = p.fill ? p.cache_off[ci] : -1;
      const bool cached = coff >= 0;  // fill: the reached list of the count pass is at hand
      if (rank == 0 && tid == 0) {
        if (!cached) {
          Mb[cl >> 5] = 1u << (cl & 31);
          L[0] = centre;
        }
        c->nL = cached ? 0 : 1; c->nR = 0; c->cnt = 0;
      }
      sn = team_sync(c, s_snap);
      if (sn.abort) break;

      // ---- phase A: k levels of frontier expansion (data.py:329-350) ----
      int lb = 0, le = cached ? 0 : 1;
      for (int level = 0; level < p.depth && lb < le; ++level) {
        const bool restricted = p.mode == DESCO_MODE_CANONICAL || (p.mode == DESCO_MODE_HETERO && level == p.depth - 1);
        team_rows(rowptr, col, L, lb, le, restricted ? centre : 0x7fffffff, rank, [&](int v, bool ok) {
          bool take = false;
          if (ok) {
            const int i = v - lo;
            const uint32_t bit = 1u << (i & 31);
            if (!(__ldcg(Mb + (i >> 5)) & bit)) take = !(atomicOr(Mb + (i >> 5), bit) & bit);
          }
          team_append(L, &c->nL, p.max_nodes, v, take);
        });
        sn = team_sync(c, s_snap);
        lb = le;
        le = min(sn.nL, p.max_nodes);
      }
      const int nL = cached ? 0 : min(sn.nL, p.max_nodes);

      // ---- phase B + C: candidates <= centre (data.py:385), component of the centre inside them (:387-390) ----
      int bfs_cnt = 0;
      if (cached) {
        const int n = p.out_nv[ci];
        for (int i = tt; i < n; i += TT) {
          const int v = p.cache[coff + i];
          R[i] = v;
          atomicOr(&RP[(v - lo) >> 5].x, 1u << ((v - lo) & 31));
        }
        if (rank == 0 && tid == 0) c->nR = n;
        sn = team_sync(c, s_snap);
      } else if (p.mode == DESCO_MODE_HETERO) {
        if (rank == 0 && tid == 0) {
          RP[cl >> 5].x = 1u << (cl & 31);
          R[0] = centre;
          c->nR = 1;
        }
        sn = team_sync(c, s_snap);
        int qb = 0, qe = 1;
        while (qb < qe && !sn.abort) {
          team_rows(rowptr, col, R, qb, qe, centre, rank, [&](int v, bool ok) {
            bool take = false;
            if (ok) {
              const int i = v - lo;
              const uint32_t bit = 1u << (i & 31);
              if (__ldcg(Mb + (i >> 5)) & bit) {
                ++bfs_cnt;  // a member <= centre next to a reached node is reached too: an induced directed edge
                if (!(__ldcg(&RP[i >> 5].x) & bit)) take = !(atomicOr(&RP[i >> 5].x, bit) & bit);
              }
            }
            team_append(R, &c->nR, p.max_nodes, v, take);
          });
          sn = team_sync(c, s_snap);
          qb = qe;
          qe = min(sn.nR, p.max_nodes);
        }
      } else {  // restricted BFS / plain k-hop ball: every member is in the result
        for (int i = tt; i < nL; i += TT) {
          const int v = __ldcg(L + i);
          R[i] = v;
          atomicOr(&RP[(v - lo) >> 5].x, 1u << ((v - lo) & 31));
        }
        if (rank == 0 && tid == 0) c->nR = nL;
        sn = team_sync(c, s_snap);
      }
      if (sn.abort) break;
      const int nv = min(sn.nR, p.max_nodes);
      const int n0 = p.fill ? p.node_off[ci] : 0;
      const int eo = p.fill ? p.edge_off[ci] : 0;

      // ---- phase D: ranks = popcount prefix over the reached bitmap (ascending node id; canonical node = last row) ----
      if (p.fill) {
        const int wchunk = (W + TM_CTAS - 1) / TM_CTAS;               // words per CTA
        const int wper = ((wchunk + TM_WARPS - 1) / TM_WARPS + 31) & ~31;  // words per warp, whole 32-word trips
        const int wb = rank * wchunk + warp * wper;
        const int we = min(min(wb + wper, (rank + 1) * wchunk), W);
        int tot = 0;
        for (int w0 = wb; w0 < we; w0 += 32) tot += (w0 + lane < we) ? __popc(__ldcg(&RP[w0 + lane].x)) : 0;
        tot = warp_sum(tot);
        if (lane == 0) s_wtot[warp] = tot;
        __syncthreads();
        if (tid == 0) {
          int t = 0;
          for (int k = 0; k < TM_WARPS; ++k) t += s_wtot[k];
          c->chunk_tot[rank] = t;
        }
        sn = team_sync(c, s_snap);
        int carry = 0;
        for (int k = 0; k < rank; ++k) carry += __ldcg(&c->chunk_tot[k]);
        for (int k = 0; k < warp; ++k) carry += s_wtot[k];
        for (int w0 = wb; w0 < we; w0 += 32) {
          const int w = w0 + lane;
          uint32_t x = (w < we) ? __ldcg(&RP[w].x) : 0u;
          const int cpop = __popc(x);
          const int incl = warp_incl_scan(cpop);
          int pre = carry + incl - cpop;
          if (x) {
            RP[w].y = (uint32_t)pre;
            while (x) {
                p.node_gid[n0 + pre++] = lo + 32 * w + __ffs(x) - 1;
                x &= x - 1;
              }
          }
          carry += __shfl_sync(FULL_MASK, incl, 31);
        }
        __syncthreads();  // s_wtot is reused below
      }

      if (!p.fill) {
        // ---- induced directed edge total (hetero: already counted by the component BFS) ----
        int cnt = bfs_cnt;
        if (p.mode != DESCO_MODE_HETERO)
          team_rows(rowptr, col, R, 0, nv, limit, rank, [&](int v, bool ok) {
            if (ok) cnt += (__ldcg(&RP[(v - lo) >> 5].x) >> ((v - lo) & 31)) & 1u;
          });
        cnt = warp_sum(cnt);
        if (lane == 0 && cnt) atomicAdd(&c->cnt, cnt);
        sn = team_sync(c, s_snap);
        if (rank == 0 && tid == 0) {
          p.out_ne[ci] = sn.cnt;
          p.out_nv[ci] = sn.cnt > 0 ? nv : 0;  // edge-free neighborhoods are dropped (workload.py:253-256)
          p.centre_graph[ci] = gid;
          long long off = -1;
          if (sn.cnt > 0) {
            off = (long long)atomicAdd(p.cache_cursor, (unsigned long long)nv);
            if (off + nv > p.cache_cap) off = -1;
          }
          p.cache_off[ci] = (int32_t)off;
          c->chunk_tot[0] = (int)off;
        }
        sn = team_sync(c, s_snap);
        const int off = __ldcg(&c->chunk_tot[0]);
        if (off >= 0)
          for (int i = tt; i < nv; i += TT) p.cache[off + i] = __ldcg(R + i);
      } else {
        if (n0 == 0 && rank == 0 && tid == 0) p.edge_ptr[0] = 0;
        sn = team_sync(c, s_snap);  // node_gid of every row is in place
        // ---- induced degree of the rows of this CTA's rank chunk (warp per row) + chunk total ----
        const int rchunk = (nv + TM_CTAS - 1) / TM_CTAS;
        const int rb0 = rank * rchunk, re0 = min(rb0 + rchunk, nv);
        int wsum = 0;
        for (int i = rb0 + warp; i < re0; i += TM_WARPS) {
          const int u = __ldcg(p.node_gid + n0 + i);
          const int rb = rowptr[u], re = rowptr[u + 1];
          int cnt = 0;
          stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
            const bool ok = valid && v <= limit;
            if (ok) cnt += (__ldcg(&RP[(v - lo) >> 5].x) >> ((v - lo) & 31)) & 1u;
            return !__all_sync(FULL_MASK, ok);
          });
          cnt = warp_sum(cnt);
          if (lane == 0) p.edge_ptr[n0 + 1 + i] = cnt;
          wsum += cnt;
        }
        if (lane == 0) s_wtot[warp] = wsum;
        __syncthreads();
        if (tid == 0) {
          int t = 0;
          for (int k = 0; k < TM_WARPS; ++k) t += s_wtot[k];
          c->chunk_tot[rank] = t;
        }
        sn = team_sync(c, s_snap);
        // ---- inclusive scan of the chunk's degrees (one warp; chunk base = totals of the lower chunks) ----
        if (warp == 0) {
          int carry = eo;
          for (int k = 0; k < rank; ++k) carry += __ldcg(&c->chunk_tot[k]);
          for (int base = rb0; base < re0; base += 32) {
            const int x = (base + lane < re0) ? __ldcg(p.edge_ptr + n0 + 1 + base + lane) : 0;
            const int incl = warp_incl_scan(x);
            if (base + lane < re0) p.edge_ptr[n0 + 1 + base + lane] = carry + incl;
            carry += __shfl_sync(FULL_MASK, incl, 31);
          }
        }
        sn = team_sync(c, s_snap);
        // ---- edges in adjacency order (ascending node id == ascending local id); types follow in edge_types_kernel ----
        constexpr int TW = TM_CTAS * TM_WARPS;
        for (int i = rank * TM_WARPS + warp; i < nv; i += TW) {
          const int u = __ldcg(p.node_gid + n0 + i);
          const int rb = rowptr[u], re = rowptr[u + 1];
          int out = (i == 0) ? eo : __ldcg(p.edge_ptr + n0 + i);
          stream_row<ROW_PF>(col, rb, re, [&](int v, bool valid) {
            const bool ok = valid && v <= limit;
            uint2 rp = make_uint2(0u, 0u);
            if (ok) rp = __ldcg(&RP[(v - lo) >> 5]);
            const uint32_t bit = 1u << ((v - lo) & 31);
            const bool in = ok && (rp.x & bit);
            const uint32_t m = __ballot_sync(FULL_MASK, in);
            if (in) p.edge_col[out + __popc(m & ((1u << lane) - 1u))] = n0 + (int)rp.y + __popc(rp.x & (bit - 1u));
            out += __popc(m);
            return !__all_sync(FULL_MASK, ok);
          });
        }
        sn = team_sync(c, s_snap);  // every team-mate is done reading the reached words
      }

      // ---- wipe the words this centre touched (nobody reads the member bits after phase C) ----
      for (int i = tt; i < nL; i += TT) Mb[(__ldcg(L + i) - lo) >> 5] = 0u;
      for (int i = tt; i < nv; i += TT) RP[(__ldcg(R + i) - lo) >> 5] = make_uint2(0u, 0u);
    }
    if (rank == 0 && tid == 0) c->cur[(iter + 1) & 1] = atomicAdd(p.ticket, 1);
    sn = team_sync(c, s_snap);
  }
  if (sn.abort && rank == 0 && tid == 0) atomicExch(p.status, DESCO_ECUDA);
}

// tunables of the large-graph path (desco_partition_large_set_caps lets the tests force the team tier on small graphs)
int g_sp_log2h0 = 13, g_sp_capl0 = 5120, g_sp_capr0 = 4096;      // tier 0: 32 + 20 + 16 KB of shared memory

struct LargeLayout {
  size_t klass_off, counters_off, list_off, coff_off, ctrl_off, slices_off, slice_bytes, cache_off, bytes;
  long long cache_cap;
  int teams, max_words;
};

int team_count() {  // co-resident teams: the cooperative launch needs every CTA of the grid on the device at once
  static int teams = 0;
  if (!teams) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, partition_team_kernel, TM_THREADS, 0) != cudaSuccess || per_sm < 1)
      per_sm = 1;
    if (per_sm > 5) per_sm = 5;
    teams = desco_num_sms() * per_sm / TM_CTAS;
    if (teams < 1) teams = 1;
  }
  return teams;
}

LargeLayout large_layout(int max_graph_nodes, int num_centres) {
  LargeLayout l;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t nc = (size_t)(num_centres > 0 ? num_centres : 1);
  l.teams = team_count();
  l.max_words = (max_graph_nodes + 31) / 32;
  l.klass_off = 0;
  l.counters_off = up(nc);
  l.list_off = l.counters_off + 256;
  l.coff_off = l.list_off + up(nc * 4);
  l.ctrl_off = l.coff_off + up(nc * 4);
  l.slices_off = l.ctrl_off + up(sizeof(TeamCtrl) * l.teams);
  l.slice_bytes = up((size_t)l.max_words * 12 + (size_t)max_graph_nodes * 8);
  l.cache_off = l.slices_off + l.slice_bytes * l.teams;
  // count -> fill cache of both tiers: 32 entries per node of the target graph, between 1M and 128M entries
  l.cache_cap = (long long)max_graph_nodes * 32;
  if (l.cache_cap < (1ll << 20)) l.cache_cap = 1ll << 20;
  if (l.cache_cap > (1ll << 27)) l.cache_cap = 1ll << 27;
  l.bytes = l.cache_off + up((size_t)l.cache_cap * 4);
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
  int* counters = (int*)(base + l.counters_off);  // [0] big centres, [1] ticket
  SparseArgs a;
  a.rowptr = rowptr; a.col = col; a.graph_ptr = graph_ptr; a.num_graphs = num_graphs;
  a.centres = centres; a.num_centres = num_centres; a.depth = depth; a.mode = mode;
  a.out_nv = nv; a.out_ne = ne; a.centre_graph = centre_graph;
  a.fill = fill; a.node_off = node_off; a.edge_off = edge_off; a.node_gid = node_gid; a.edge_ptr = edge_ptr; a.edge_col = edge_col;
  a.klass = base + l.klass_off;
  a.big_list = (int32_t*)(base + l.list_off); a.big_count = counters;
  a.cache_cursor = (unsigned long long*)(counters + 2);  // 8-byte aligned: counters is 256-byte aligned
  a.cache_off = (int32_t*)(base + l.coff_off); a.cache = (int32_t*)(base + l.cache_off); a.cache_cap = l.cache_cap;
  DescoProfScope prof(DESCO_PROF_PARTITION, stream, 2);
  // the count pass builds the list of big centres; fill reuses it.  Team state (barriers, bitmaps) starts from zero.
  if (!fill) DESCO_CUDA_TRY(cudaMemsetAsync(counters, 0, 256, stream));
  else DESCO_CUDA_TRY(cudaMemsetAsync(counters + 1, 0, sizeof(int), stream));
  DESCO_CUDA_TRY(cudaMemsetAsync(base + l.ctrl_off, 0, l.slices_off - l.ctrl_off, stream));
  for (int t = 0; t < l.teams; ++t)
    DESCO_CUDA_TRY(cudaMemsetAsync(base + l.slices_off + (size_t)t * l.slice_bytes, 0, (size_t)l.max_words * 12, stream));
  {  // tier 0: shared memory
    a.log2H = g_sp_log2h0; a.capL = g_sp_capl0; a.capR = g_sp_capr0;
    const size_t hslots = (size_t)1 << a.log2H;
    const size_t smem = (hslots + (hslots / 2 > (size_t)a.capL ? hslots / 2 : (size_t)a.capL) + a.capR) * 4;
    if (smem > 200 * 1024) return DESCO_EINVAL;
    DESCO_CUDA_TRY(cudaFuncSetAttribute(partition_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    const int blocks = num_centres < sms * per_sm ? num_centres : sms * per_sm;
    partition_sparse_kernel<<<blocks, SP_THREADS, smem, stream>>>(a);
    DESCO_LAUNCH_CHECK();
  }
  {  // team tier: global bitmaps, TM_CTAS co-resident CTAs per centre
    TeamArgs t;
    t.rowptr = rowptr; t.col = col; t.graph_ptr = graph_ptr; t.num_graphs = num_graphs;
    t.centres = centres; t.num_centres = num_centres; t.depth = depth; t.mode = mode;
    t.out_nv = nv; t.out_ne = ne; t.centre_graph = centre_graph;
    t.fill = fill; t.node_off = node_off; t.edge_off = edge_off; t.node_gid = node_gid; t.edge_ptr = edge_ptr; t.edge_col = edge_col;
    t.big_list = a.big_list; t.big_count = counters; t.ticket = counters + 1;
    t.ctrl = (TeamCtrl*)(base + l.ctrl_off); t.slices = base + l.slices_off; t.slice_bytes = l.slice_bytes;
    t.max_words = l.max_words; t.max_nodes = max_graph_nodes; t.status = status;
    t.cache_cursor = (unsigned long long*)(counters + 2);  // 8-byte aligned: counters is 256-byte aligned
    t.cache_off = (int32_t*)(base + l.coff_off); t.cache = (int32_t*)(base + l.cache_off); t.cache_cap = l.cache_cap;
    void* params[] = {(void*)&t};
    DESCO_CUDA_TRY(cudaLaunchCooperativeKernel((const void*)partition_team_kernel, dim3(l.teams * TM_CTAS), dim3(TM_THREADS),
                                               params, 0, stream));
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
                                     uint8_t* __restrict__ indicator, int32_t* __restrict__ totals,
                                     int32_t* __restrict__ max_rows, unsigned long long* __restrict__ sums64) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < C;
  if (sums64) {  // exact row / edge totals (caller-zeroed): the int32 scans above wrap silently past 2^31
    unsigned long long v = in ? (unsigned long long)nv[i] : 0ull, e = in ? (unsigned long long)ne[i] : 0ull;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      v += __shfl_xor_sync(FULL_MASK, v, d);
      e += __shfl_xor_sync(FULL_MASK, e, d);
    }
    if (lane_id() == 0 && (v | e)) {
      atomicAdd(&sums64[0], v);
      atomicAdd(&sums64[1], e);
    }
  }
  if (!in) i = C - 1;  // keep the warp whole for the shuffle below; the duplicate work is idempotent
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
  if (max_rows) {  // rows of the largest kept neighborhood (caller-zeroed)
    int m = keep ? nv[i] : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(FULL_MASK, m, d));
    if (lane_id() == 0 && m > 0) atomicMax(max_rows, m);
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
                                                                num_centres, nbh_ptr, centre_out, indicator, totals, nullptr,
                                                                nullptr);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // extern "C"

namespace {
// Pass 2 for batches of up to a few 10^4 centres in ONE single-CTA launch (instead of three cub scans + finalize =
// seven launches): exclusive scans of (kept, rows, edges), the neighborhood pointers, indicator and totals.
__global__ void __launch_bounds__(1024) scan_small_kernel(const int32_t* __restrict__ centres, const int32_t* __restrict__ nv,
                                                          const int32_t* __restrict__ ne, int C, int32_t* __restrict__ node_off,
                                                          int32_t* __restrict__ edge_off, int32_t* __restrict__ nbh_ptr,
                                                          int32_t* __restrict__ centre_out, uint8_t* __restrict__ indicator,
                                                          int32_t* __restrict__ totals) {
  __shared__ int s_w[3][32];
  __shared__ int s_carry[3];
  const int tid = threadIdx.x, lane = lane_id(), warp = warp_id();
  if (tid < 3) s_carry[tid] = 0;
  int mx = 0;
  unsigned long long v64 = 0, e64 = 0;  // exact totals -> totals[8..11] (two u64): the int32 scans wrap past 2^31
  __syncthreads();
  for (int base = 0; base < C; base += 1024) {
    const int i = base + tid;
    const int v = i < C ? nv[i] : 0, e = i < C ? ne[i] : 0, k = e > 0 ? 1 : 0;
    v64 += (unsigned long long)v;
    e64 += (unsigned long long)e;
    int sk = warp_incl_scan(k), sv = warp_incl_scan(v), se = warp_incl_scan(e);
    if (lane == 31) { s_w[0][warp] = sk; s_w[1][warp] = sv; s_w[2][warp] = se; }
    __syncthreads();
    if (warp < 3) {
      const int x = s_w[warp][lane];
      const int incl = warp_incl_scan(x);
      s_w[warp][lane] = incl - x;  // exclusive base of warp `lane` for quantity `warp`
    }
    __syncthreads();
    const int rk = s_carry[0] + s_w[0][warp] + sk - k;
    const int no = s_carry[1] + s_w[1][warp] + sv - v;
    const int eo = s_carry[2] + s_w[2][warp] + se - e;
    if (i < C) {
      node_off[i] = no;
      edge_off[i] = eo;
      if (indicator) indicator[i] = (uint8_t)k;
      if (k) {
        nbh_ptr[rk] = no;
        centre_out[rk] = centres[i];
        mx = max(mx, v);
      }
      if (i == C - 1) {
        nbh_ptr[rk + k] = no + v;
        totals[0] = rk + k;
        totals[1] = no + v;
        totals[2] = eo + e;
      }
    }
    __syncthreads();
    if (tid == 1023) { s_carry[0] = rk + k; s_carry[1] = no + v; s_carry[2] = eo + e; }
    __syncthreads();
  }
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 16));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 8));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 4));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 2));
  mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, 1));
  if (lane == 0 && mx > 0) atomicMax(&totals[3], mx);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v64 += __shfl_xor_sync(FULL_MASK, v64, d);
    e64 += __shfl_xor_sync(FULL_MASK, e64, d);
  }
  if (lane == 0) {
    unsigned long long* sums64 = reinterpret_cast<unsigned long long*>(totals + 8);
    atomicAdd(&sums64[0], v64);
    atomicAdd(&sums64[1], e64);
  }
}
}  // namespace

extern "C" {

int64_t desco_partition_batch_workspace_bytes(int32_t num_centres) {
  const int64_t c = num_centres > 0 ? num_centres : 1;
  return ((5 * c * 4 + 255) / 256) * 256 + 256 + desco_partition_scan_workspace_bytes(num_centres);
}

int desco_partition_batch(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                          const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                          int32_t max_graph_nodes, void* workspace, int64_t workspace_bytes, int32_t* nbh_ptr,
                          int32_t* centre_out, uint8_t* indicator, int32_t* centre_graph, int32_t* node_gid,
                          int32_t* edge_ptr, int64_t cap_rows, int32_t* edge_col, uint8_t* edge_tri, int64_t cap_edges,
                          int32_t* totals_host, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!totals_host || !nbh_ptr || num_centres < 0) return DESCO_EINVAL;
  totals_host[0] = totals_host[1] = totals_host[2] = totals_host[3] = 0;
  if (num_centres == 0) {
    DESCO_CUDA_TRY(cudaMemsetAsync(nbh_ptr, 0, sizeof(int32_t), s));
    return DESCO_OK;
  }
  if (!workspace || workspace_bytes < desco_partition_batch_workspace_bytes(num_centres) || !centre_out || !centre_graph)
    return DESCO_EINVAL;
  const size_t c = (size_t)num_centres;
  int32_t* nv = (int32_t*)workspace;
  int32_t* ne = nv + c;
  int32_t* rank = ne + c;
  int32_t* noff = rank + c;
  int32_t* eoff = noff + c;
  int32_t* small = (int32_t*)((char*)workspace + ((5 * c * 4 + 255) / 256) * 256);  // totals[3], max rows, status
  void* scan_ws = (char*)small + 256;
  const int64_t scan_bytes = desco_partition_scan_workspace_bytes(num_centres);
  DESCO_CUDA_TRY(cudaMemsetAsync(small, 0, 256, s));
  int rc = launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes, nv, ne,
                            centre_graph, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, small + 4, s);
  if (rc) return rc;
  desco_count_launches(1);
  if (num_centres <= 65536) {
    scan_small_kernel<<<1, 1024, 0, s>>>(centres, nv, ne, num_centres, noff, eoff, nbh_ptr, centre_out, indicator, small);
  } else {
    size_t bytes = (size_t)scan_bytes;
    cub::TransformInputIterator<int, KeepFlag, const int32_t*> it(ne, KeepFlag());
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, it, rank, num_centres, s));
    bytes = (size_t)scan_bytes;
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, nv, noff, num_centres, s));
    bytes = (size_t)scan_bytes;
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, ne, eoff, num_centres, s));
    scan_finalize_kernel<<<(num_centres + 255) / 256, 256, 0, s>>>(centres, nv, ne, rank, noff, eoff, num_centres, nbh_ptr,
                                                                  centre_out, indicator, small, small + 3,
                                                                  reinterpret_cast<unsigned long long*>(small + 8));
  }
  DESCO_LAUNCH_CHECK();
  // the one host sync of the path: output sizes (and the device status word) through a pinned staging buffer
  static thread_local int32_t* pinned = nullptr;
  if (!pinned) DESCO_CUDA_TRY(cudaHostAlloc((void**)&pinned, 64, cudaHostAllocDefault));
  DESCO_CUDA_TRY(cudaMemcpyAsync(pinned, small, 12 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  DESCO_CUDA_TRY(cudaStreamSynchronize(s));
  for (int i = 0; i < 4; ++i) totals_host[i] = pinned[i];
  if (pinned[4] != 0) return pinned[4];
  {  // exact totals: a packed batch addresses rows and edges with int32 - the caller must split the centre list
    const unsigned long long* sums64 = reinterpret_cast<const unsigned long long*>(pinned + 8);
    if (sums64[0] > 0x7fffffffull || sums64[1] > 0x7fffffffull) {
      totals_host[1] = totals_host[2] = -1;
      return DESCO_ERANGE;
    }
  }
  if (pinned[1] > cap_rows || pinned[2] > cap_edges) return DESCO_ENOBUFS;  // caller re-allocates from totals_host
  if (pinned[1] == 0) {
    if (edge_ptr) DESCO_CUDA_TRY(cudaMemsetAsync(edge_ptr, 0, sizeof(int32_t), s));
    return DESCO_OK;
  }
  if (!node_gid || !edge_ptr || !edge_col || !edge_tri) return DESCO_EINVAL;
  return launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes, nv, ne,
                          centre_graph, 1, noff, eoff, node_gid, edge_ptr, edge_col, edge_tri, small + 4, s);
}

}  // extern "C"

namespace {
// Capacity guard of the stream-ordered (no host sync) form: when the exact row / edge totals do not fit the caller's
// buffers, every neighborhood is marked dropped (the fill pass skips them), the effective sizes become 0 and the status
// word reports ENOBUFS / ERANGE.  small: [0..3] totals of the scan, [4] status, [8..11] exact 64-bit sums,
// [12..15] effective {G, V, E, max rows} that every later kernel of the step reads.
__global__ void partition_guard_kernel(int32_t* __restrict__ ne, int C, int32_t* __restrict__ small, long long cap_rows,
                                       long long cap_edges) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long* sums64 = reinterpret_cast<const unsigned long long*>(small + 8);
  const bool range = sums64[0] > 0x7fffffffull || sums64[1] > 0x7fffffffull;
  const bool over = range || (long long)sums64[0] > cap_rows || (long long)sums64[1] > cap_edges || small[4] != 0;
  if (i < C && over) ne[i] = 0;
  if (i == 0) {
    if (over && small[4] == 0) small[4] = range ? DESCO_ERANGE : DESCO_ENOBUFS;
    for (int k = 0; k < 4; ++k) small[12 + k] = over ? 0 : small[k];
  }
}
}  // namespace

extern "C" {

int desco_partition_batch_async(const int32_t* rowptr, const int32_t* col, const int32_t* graph_ptr, int32_t num_graphs,
                                const int32_t* centres, int32_t num_centres, int32_t depth, int32_t mode,
                                int32_t max_graph_nodes, void* workspace, int64_t workspace_bytes, int32_t* nbh_ptr,
                                int32_t* centre_out, uint8_t* indicator, int32_t* centre_graph, int32_t* node_gid,
                                int32_t* edge_ptr, int64_t cap_rows, int32_t* edge_col, uint8_t* edge_tri, int64_t cap_edges,
                                int32_t* sizes_dev, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (!sizes_dev || !nbh_ptr || num_centres <= 0 || cap_rows < 1 || cap_edges < 1) return DESCO_EINVAL;
  if (!workspace || workspace_bytes < desco_partition_batch_workspace_bytes(num_centres) || !centre_out || !centre_graph ||
      !node_gid || !edge_ptr || !edge_col || !edge_tri)
    return DESCO_EINVAL;
  const size_t c = (size_t)num_centres;
  int32_t* nv = (int32_t*)workspace;
  int32_t* ne = nv + c;
  int32_t* rank = ne + c;
  int32_t* noff = rank + c;
  int32_t* eoff = noff + c;
  void* scan_ws = (char*)workspace + ((5 * c * 4 + 255) / 256) * 256 + 256;
  const int64_t scan_bytes = desco_partition_scan_workspace_bytes(num_centres);
  int32_t* small = sizes_dev;
  DESCO_CUDA_TRY(cudaMemsetAsync(small, 0, 16 * sizeof(int32_t), s));
  DESCO_CUDA_TRY(cudaMemsetAsync(edge_ptr, 0, sizeof(int32_t), s));
  int rc = launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes, nv, ne,
                            centre_graph, 0, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, small + 4, s);
  if (rc) return rc;
  desco_count_launches(2);
  if (num_centres <= 65536) {
    scan_small_kernel<<<1, 1024, 0, s>>>(centres, nv, ne, num_centres, noff, eoff, nbh_ptr, centre_out, indicator, small);
  } else {
    size_t bytes = (size_t)scan_bytes;
    cub::TransformInputIterator<int, KeepFlag, const int32_t*> it(ne, KeepFlag());
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, it, rank, num_centres, s));
    bytes = (size_t)scan_bytes;
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, nv, noff, num_centres, s));
    bytes = (size_t)scan_bytes;
    DESCO_CUDA_TRY(cub::DeviceScan::ExclusiveSum(scan_ws, bytes, ne, eoff, num_centres, s));
    scan_finalize_kernel<<<(num_centres + 255) / 256, 256, 0, s>>>(centres, nv, ne, rank, noff, eoff, num_centres, nbh_ptr,
                                                                  centre_out, indicator, small, small + 3,
                                                                  reinterpret_cast<unsigned long long*>(small + 8));
  }
  partition_guard_kernel<<<(num_centres + 255) / 256, 256, 0, s>>>(ne, num_centres, small, (long long)cap_rows,
                                                                  (long long)cap_edges);
  DESCO_LAUNCH_CHECK();
  return launch_partition(rowptr, col, graph_ptr, num_graphs, centres, num_centres, depth, mode, max_graph_nodes, nv, ne,
                          centre_graph, 1, noff, eoff, node_gid, edge_ptr, edge_col, edge_tri, small + 4, s);
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
  (void)log2_slots1; (void)members1; (void)reached1;  // the team tier has no capacity limit (bitmaps over the target graph)
  if (log2_slots0 < 4 || log2_slots0 > 15 || !pow2(reached0) || reached0 > 32768 || members0 < 1 ||
      members0 + SP_THREADS > (1 << log2_slots0))
    return DESCO_EINVAL;
  g_sp_log2h0 = log2_slots0; g_sp_capl0 = members0; g_sp_capr0 = reached0;
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

int desco_partition_large_phase_cycles(uint64_t* out, int32_t reset) {
  if (!out) return DESCO_EINVAL;
  unsigned long long h[SPH_COUNT];
  DESCO_CUDA_TRY(cudaMemcpyFromSymbol(h, g_sparse_phase_cycles, sizeof(h)));
  for (int i = 0; i < SPH_COUNT; ++i) out[i] = h[i];
  if (reset) {
    for (int i = 0; i < SPH_COUNT; ++i) h[i] = 0;
    DESCO_CUDA_TRY(cudaMemcpyToSymbol(g_sparse_phase_cycles, h, sizeof(h)));
  }
  return DESCO_OK;
}

const char* desco_version(void) { return "desco_b200 0.1 sm_100a"; }

}  // extern "C"
