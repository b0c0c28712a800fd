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

// ToTconvHetero on an existing packed batch: one warp per row, one lane per incident edge.
__global__ void edge_types_kernel(const int32_t* __restrict__ edge_ptr, const int32_t* __restrict__ edge_col,
                                  int num_rows, uint8_t* __restrict__ edge_tri) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= num_rows) return;
  const int rb = edge_ptr[row], re = edge_ptr[row + 1];
  for (int e = rb + lane_id(); e < re; e += 32) {
    int v = edge_col[e];
    int a = rb, b = edge_ptr[v], be = edge_ptr[v + 1];
    bool tri = false;
    while (a < re && b < be) {
      int x = edge_col[a], y = edge_col[b];
      if (x == y) { tri = true; break; }
      if (x < y) ++a; else ++b;
    }
    edge_tri[e] = tri ? 1 : 0;
  }
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

const char* desco_version(void) { return "desco_b200 0.1 sm_100a"; }

}  // extern "C"
