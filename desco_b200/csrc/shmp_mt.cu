// SHMP message-passing layers for neighborhoods of ANY size on the tensor cores ("multi-tile" path), sm_100a.
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a), like shmp.cu / shmp_fused.cu:
//   subgraph_counting/gnn_model.py:362-404  SAGEConv (sum-aggregate then Linear, per relation)
//   subgraph_counting/gnn_model.py:230-277  BaseGNNCore.forward as expanded by to_hetero_old (lightning_model.py:371-421)
//   subgraph_counting/gnn_model.py:107      global_add_pool
// The reference has no size limit on a neighborhood.  shmp_fused.cu keeps a whole neighborhood inside one 128-row tile
// (configs 1, 2, 4); Syn_1827-shaped batches (up to ~700 rows) and the depth-2 balls of a power-law target (10^3 - 10^6
// rows, config 5) go through this file: features live in HBM between layers ([V][64] fp32), a tile is 128 consecutive
// count rows of the packed batch whatever neighborhoods they belong to, and one layer is
//   h'_i = relu( [ sum_{j in N_tri(i)} h_j | sum_{j in N_tride(i)} h_j | h_i ] . Wc + bias_c + cvec_{tri/tride}[nbh(i)] )
// per tile: the segmented, edge-type-split GATHER out of global memory (mostly L2 hits: a neighborhood's rows are
// contiguous) writes the [128 x 192] operand straight into shared memory as bf16 hi/lo SWIZZLE_128B images, the
// [128 x 192] . [192 x 64] product runs on tcgen05 (three hi/lo passes = fp32-grade, accumulator in TMEM) and the epilogue
// reads TMEM, adds bias / canonical-node term, applies relu and writes h'.  The kernel is HBM/L2-gather bound:
// algorithmic bytes per layer = 4F (E + 2V), which is also what it moves.
// The canonical rows (one per neighborhood, K = 192 GEMV with their own weights) and the per-layer pooled sums
// (global_add_pool, deterministic two-level reduction) are separate small kernels.
#include "common.cuh"
#include "shmp_internal.h"
#include "tc05.cuh"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;
constexpr int TR = 128;                      // rows per tile = UMMA M
constexpr int KB = 3;                        // 64-wide K blocks: tri | tride | self
constexpr int THREADS = 512;
constexpr int NW = THREADS / 32;             // 16 warps x 128 registers (16 row loads in flight each); rows of a tile are dealt by a ticket
constexpr int EPI_WARPS = 16;                // the warps that run the epilogue (4 lane quarters x 4 column groups)
constexpr int MID_DEG = 96;                  // rows with more edges are gathered by a GROUP of 4 warps (ncu: 41 % of the stall
                                             // samples sat at the barrier behind the one warp that drew a long row) ...
constexpr int HUB_DEG = 2048;                // ... and beyond this by the whole CTA
constexpr int GW = 4;                        // warps per group
constexpr int NG = 4;                        // groups per CTA
constexpr int IMG = TR * 128;                // one bf16 image of a [128 x 64] block
constexpr int B_IMG = F * 128;               // one bf16 image of a [64 n x 64 k] weight block
constexpr int SM_B = 0;                                  // [KB][hi | lo] weight images, 48 KB
constexpr int SM_A = SM_B + KB * 2 * B_IMG;              // [KB][hi | lo] operand images, 96 KB
constexpr int SM_ROW = SM_A + KB * 2 * IMG;              // 2 x { int s_row, s_g, s_code, s_eb, s_ee [TR] }
constexpr int SM_HUB = SM_ROW + 2 * 5 * TR * 4;          // float4 s_hub[NW][16][2]
constexpr int SM_BIAS = SM_HUB + NW * 16 * 32;           // float bias[64]
constexpr int SM_BARS = SM_BIAS + F * 4;                 // 2 mbarriers, tmem slot, hub / mid counts, tickets, hub list, mid list
constexpr int SM_TOTAL = SM_BARS + 64 + 2 * TR;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;
static_assert(SM_A % 1024 == 0 && B_IMG % 1024 == 0 && IMG % 1024 == 0, "UMMA tiles must be 1024-B aligned");
static_assert(SMEM_BYTES <= 232448, "multi-tile SHMP kernel exceeds the 227 KB shared-memory limit");

struct MtArgs {
  const int32_t* edge_ptr; const int32_t* edge_col; const uint8_t* edge_tri;
  const int32_t* row_nbh; const int32_t* crow; const uint8_t* canon_code;
  int Vc, num_tiles, passes;
  const float* h_in; float* h_out;   // [V][F]; canonical rows are kept at zero, so gathering them adds nothing
  const float* cvec;                 // [G][2F] canonical -> count term of this layer (NULL for single-type graphs)
  const uint8_t* w_img;              // [KB][hi | lo] images of Wc^T blocks
  const float* bias_c;               // [F]
  int32_t* status;
};

__device__ __forceinline__ void add4(float4& a, const float4 v) { a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }

// D rows (two per warp instruction: a half-warp per 256-byte row, 16 bytes per lane) of one 32-edge chunk in flight
template <int D>
__device__ __forceinline__ void gather_chunk(const float* __restrict__ h, int cur, int n, int half, int l16, float4& at,
                                             float4& ad) {
  float4 v[D];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int j = 2 * i + half;
    const int s = __shfl_sync(FULL_MASK, cur, j);
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n) v[i] = __ldg(reinterpret_cast<const float4*>(h + (size_t)(s & 0x7fffffff) * F) + l16);
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    const int s = __shfl_sync(FULL_MASK, cur, 2 * i + half);
    if (s < 0) add4(at, v[i]); else add4(ad, v[i]);  // bit 31 = triangle edge (a padded slot is 0 and adds 0 to tride)
  }
}

// sum over edges [eb, ee) in 32-edge chunks `first`, `first + step`, ..., split by SHMP type (bit 31 of the packed
// source word).  Lane = (edge parity, 4 features): on return every lane of BOTH halves holds the sums of its 4 features.
// The next chunk's sources load while the current chunk's (up to 32) rows are in flight.
__device__ __forceinline__ void gather_edges(const float* __restrict__ h, const int32_t* __restrict__ edge_col,
                                             const uint8_t* __restrict__ edge_tri, int eb, int ee, int first, int step,
                                             int lane, float4& at, float4& ad) {
  const int half = lane >> 4, l16 = lane & 15;
  int base = eb + 32 * first;
  int my = 0;
  if (base + lane < ee) my = edge_col[base + lane] | (edge_tri[base + lane] ? (int)0x80000000 : 0);
  while (base < ee) {
    const int n = min(32, ee - base);
    const int cur = my;
    base += 32 * step;
    my = 0;
    if (base + lane < ee) my = edge_col[base + lane] | (edge_tri[base + lane] ? (int)0x80000000 : 0);
    if (n <= 8) gather_chunk<4>(h, cur, n, half, l16, at, ad);
    else if (n <= 16) gather_chunk<8>(h, cur, n, half, l16, at, ad);
    else gather_chunk<16>(h, cur, n, half, l16, at, ad);
  }
#define DESCO_XH(a) a += __shfl_xor_sync(FULL_MASK, a, 16)
  DESCO_XH(at.x); DESCO_XH(at.y); DESCO_XH(at.z); DESCO_XH(at.w);
  DESCO_XH(ad.x); DESCO_XH(ad.y); DESCO_XH(ad.z); DESCO_XH(ad.w);
#undef DESCO_XH
}

// 4 consecutive features of row r as bf16 hi / lo, swizzled (8-byte stores; 16 lanes cover the 128-byte row)
__device__ __forceinline__ void store_quad(uint8_t* img, int r, int l16, float4 v) {
  const uint32_t off = tc05::sw128_offset(r, 4 * l16);
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
  *reinterpret_cast<uint2*>(img + off) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
  *reinterpret_cast<uint2*>(img + IMG + off) =
      make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
}

__global__ void __launch_bounds__(THREADS, 1) shmp_mt_layer_kernel(const MtArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem + SM_B;
  uint8_t* sA = smem + SM_A;
  int* s_row = reinterpret_cast<int*>(smem + SM_ROW);
  int* s_g = s_row + TR;
  int* s_code = s_g + TR;
  int* s_eb = s_code + TR;
  int* s_ee = s_eb + TR;
  float4* s_hub = reinterpret_cast<float4*>(smem + SM_HUB);
  float* s_bias = reinterpret_cast<float*>(smem + SM_BIAS);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);  // [0] weights, [1] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  int* s_nhub = reinterpret_cast<int*>(tmem_slot + 1);
  int* s_ticket = s_nhub + 1;  // [2], by tile parity
  int* s_nmid = s_ticket + 2;
  int* s_mid_ticket = s_nmid + 1;
  int* s_grow = s_mid_ticket + 1;  // [NG] the row a group is working on
  uint8_t* s_hubs = reinterpret_cast<uint8_t*>(bars) + 64;
  uint8_t* s_mids = s_hubs + TR;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  if ((int)blockIdx.x >= p.num_tiles) return;
  if (tid == 0) {
    tc05::mbar_init(&bars[0], 1);
    tc05::mbar_init(&bars[1], 1);
    tc05::fence_mbar_init();
  }
  if (warp == 0) tc05::tmem_alloc(tmem_slot, 64);
  if (tid < F) s_bias[tid] = p.bias_c[tid];
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {  // the layer's weight images once per CTA
    tc05::mbar_arrive_expect_tx(&bars[0], KB * 2 * B_IMG);
    for (int off = 0; off < KB * 2 * B_IMG; off += 16384) tc05::bulk_g2s(sB + off, p.w_img + off, 16384, &bars[0]);
  }
  const uint32_t idesc = tc05::make_idesc_bf16(TR, F);
  uint32_t mphase = 0;
  bool weights_ready = false;

  // per-row metadata of a tile, double-buffered by tile parity: the epilogue of tile t (warps 0..15) reads buffer b while
  // the other warps are already gathering tile t+1 out of buffer b^1
  auto setup = [&](int tile, int b) {
    if (tid < TR) {
      const int k = tile * TR + tid;
      int row = -1, g = -1, code = 0;
      if (tile < p.num_tiles && k < p.Vc) {
        row = p.crow[k];
        g = p.row_nbh[row];
        code = p.canon_code[row];
      }
      int eb = 0, ee = 0;
      if (row >= 0) { eb = p.edge_ptr[row]; ee = p.edge_ptr[row + 1]; }
      s_row[b * 5 * TR + tid] = row;
      s_g[b * 5 * TR + tid] = g;
      s_code[b * 5 * TR + tid] = code;
      s_eb[b * 5 * TR + tid] = eb;  // fetched one tile ahead: a row's gather starts with its source words
      s_ee[b * 5 * TR + tid] = ee;
    }
    if (tid == TR) s_ticket[b] = 0;
  };
  setup(blockIdx.x, 0);
  if (tid == 0) { *s_nhub = 0; *s_nmid = 0; *s_mid_ticket = 0; }
  __syncthreads();
  const int grp = warp / GW, gw = warp % GW;
  auto group_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(32 * GW) : "memory"); };
  int buf = 0;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, buf ^= 1) {
    const int* t_row = s_row + buf * 5 * TR;
    const int* t_g = s_g + buf * 5 * TR;
    const int* t_code = s_code + buf * 5 * TR;
    const int* t_eb = s_eb + buf * 5 * TR;
    const int* t_ee = s_ee + buf * 5 * TR;

    // ---- gather: [sum tri | sum tride | self] of every row -> bf16 hi/lo operand images; rows are dealt by a ticket, so a
    //      warp that drew a long row (or is still in the previous tile's epilogue) simply takes fewer of them ----
#pragma unroll 1
    for (;;) {
      int r = 0;
      if (lane == 0) r = atomicAdd(&s_ticket[buf], 1);
      r = __shfl_sync(FULL_MASK, r, 0);
      if (r >= TR) break;
      const int row = t_row[r];
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f), ad = at, self = at;
      bool hub = false;
      if (row >= 0) {
        if (half == 0) self = __ldg(reinterpret_cast<const float4*>(p.h_in + (size_t)row * F) + l16);
        const int eb = t_eb[r], ee = t_ee[r];
        hub = ee - eb > MID_DEG;
        if (!hub) gather_edges(p.h_in, p.edge_col, p.edge_tri, eb, ee, 0, 1, lane, at, ad);
        else if (lane == 0) {
          if (ee - eb > HUB_DEG) s_hubs[atomicAdd(s_nhub, 1)] = (uint8_t)r;
          else s_mids[atomicAdd(s_nmid, 1)] = (uint8_t)r;
        }
      }
      if (!hub) store_quad(sA + (half ? 2 * IMG : 0), r, l16, half ? ad : at);  // the two halves write one block each
      if (half == 0) store_quad(sA + 4 * IMG, r, l16, self);
    }
    __syncthreads();  // every warp is past the previous tile's epilogue here: the other metadata buffer is free
    setup(tile + gridDim.x, buf ^ 1);
    // mid rows: dealt by a ticket to the 4 groups of 4 warps; a group's warps take the row's 32-edge chunks round robin and
    // its first warp adds the four partial sums in warp order (deterministic), so long rows proceed four at a time
    for (const int nmid = *s_nmid;;) {
      if ((tid & (32 * GW - 1)) == 0) s_grow[grp] = atomicAdd(s_mid_ticket, 1);
      group_bar();
      const int t = s_grow[grp];
      if (t >= nmid) break;
      const int r = s_mids[t];
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f), ad = at;
      gather_edges(p.h_in, p.edge_col, p.edge_tri, t_eb[r], t_ee[r], gw, GW, lane, at, ad);
      if (half == 0) {
        s_hub[(warp * 16 + l16) * 2] = at;
        s_hub[(warp * 16 + l16) * 2 + 1] = ad;
      }
      group_bar();
      if (gw == 0) {
        float4 tt = s_hub[(warp * 16 + l16) * 2 + half];  // half 0 sums the triangle partials, half 1 the tride partials
        for (int w = 1; w < GW; ++w) add4(tt, s_hub[((warp + w) * 16 + l16) * 2 + half]);
        store_quad(sA + (half ? 2 * IMG : 0), r, l16, tt);
      }
      group_bar();
    }
    __syncthreads();  // (s_hub is shared with the whole-CTA pass below)
    for (int hh = 0, nh = *s_nhub; hh < nh; ++hh) {  // hub rows: 32-edge chunks dealt over all warps, summed in warp order
      const int r = s_hubs[hh], row = t_row[r];
      float4 at = make_float4(0.f, 0.f, 0.f, 0.f), ad = at;
      (void)row;
      gather_edges(p.h_in, p.edge_col, p.edge_tri, t_eb[r], t_ee[r], warp, NW, lane, at, ad);
      if (half == 0) {
        s_hub[(warp * 16 + l16) * 2] = at;
        s_hub[(warp * 16 + l16) * 2 + 1] = ad;
      }
      __syncthreads();
      if (warp == 0) {
        float4 t = s_hub[l16 * 2 + half];  // half 0 sums the triangle partials, half 1 the tride partials
        for (int w = 1; w < NW; ++w) add4(t, s_hub[(w * 16 + l16) * 2 + half]);
        store_quad(sA + (half ? 2 * IMG : 0), r, l16, t);
      }
      __syncthreads();
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before_sync();
    __syncthreads();

    // ---- [128 x 192] . [192 x 64] on tcgen05: hi.hi + lo.hi + hi.lo ----
    if (tid == 0) {
      *s_nhub = 0;  // every thread has read the hub / mid counts (barrier above); the next tile's gather, which pushes its
      *s_nmid = 0;  // long rows, starts only after the MMA issued below has completed
      *s_mid_ticket = 0;
      if (!weights_ready) {
        if (!tc05::mbar_wait(&bars[0], 0)) { atomicExch(p.status, DESCO_ECUDA); __trap(); }
        weights_ready = true;
      }
      tc05::fence_after_sync();
      bool acc = false;
      for (int pass = 0; pass < p.passes; ++pass) {
        for (int kb = 0; kb < KB; ++kb) {
          const uint64_t da = tc05::make_smem_desc(sA + kb * 2 * IMG + (pass == 1 ? IMG : 0));
          const uint64_t db = tc05::make_smem_desc(sB + kb * 2 * B_IMG + (pass == 2 ? B_IMG : 0));
#pragma unroll
          for (int k = 0; k < F / 16; ++k) {
            tc05::mma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, acc);
            acc = true;
          }
        }
      }
      tc05::mma_commit(&bars[1]);
    }
    if (!tc05::mbar_wait(&bars[1], mphase)) { atomicExch(p.status, DESCO_ECUDA); __trap(); }  // never hang the box
    mphase ^= 1;
    tc05::fence_after_sync();

    // ---- epilogue (warps 0..15; the others go straight on to the next tile's gather): bias, canonical term, relu ----
    if (warp < EPI_WARPS) {
      const int qd = warp & 3, cg = warp >> 2;
      const int r = 32 * qd + lane;
      float v[16];
      tc05::tmem_ld16(tmem + (static_cast<uint32_t>(32 * qd) << 16) + 16 * cg, v);
      const int row = t_row[r];
      if (row >= 0) {
        const int code = t_code[r];
        const float* cv = (code && p.cvec) ? p.cvec + (size_t)t_g[r] * 2 * F + (code - 1) * F + 16 * cg : nullptr;
        float4* dst = reinterpret_cast<float4*>(p.h_out + (size_t)row * F + 16 * cg);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float x = v[4 * q4 + i] + s_bias[16 * cg + 4 * q4 + i];
            if (cv) x += cv[4 * q4 + i];
            o[i] = fmaxf(x, 0.f);  // gnn_model.py:273
          }
          dst[q4] = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
      tc05::fence_before_sync();
    }
  }
  if (tid == 0 && !weights_ready) tc05::mbar_wait(&bars[0], 0);  // never exit with a bulk copy in flight
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------------------------
// canonical rows: emb_a[g][(l+1)F..] = relu( [sum_tri h_c | sum_tride h_c | h_a] . Wa + bias_a ), one CTA per neighborhood
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void add_sel(float2& at, float2& ad, const float2 v, int s) {
  if (s < 0) { at.x += v.x; at.y += v.y; } else { ad.x += v.x; ad.y += v.y; }
}

constexpr int CT = 256;
__global__ void __launch_bounds__(CT) shmp_mt_canon_kernel(const int32_t* __restrict__ nbh_ptr, const int32_t* __restrict__ edge_ptr,
                                                           const int32_t* __restrict__ edge_col, const uint8_t* __restrict__ edge_tri,
                                                           const int32_t* __restrict__ quirk_row, const float* __restrict__ h_in,
                                                           float* __restrict__ emb_a, int emb_ld, int layer,
                                                           const float* __restrict__ Wa, const float* __restrict__ bias_a) {
  __shared__ float4 s_part[CT / 32][32];
  __shared__ float s_a[3 * F];
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row = nbh_ptr[g + 1] - 1;
  const int quirk = quirk_row[g];
  const int eb = edge_ptr[row], ee = edge_ptr[row + 1];
  float2 at = make_float2(0.f, 0.f), ad = at;
  for (int base = eb + 32 * warp; base < ee; base += 32 * (CT / 32)) {
    const int e = base + lane;
    int my = 0;
    if (e < ee) my = edge_col[e] | (edge_tri[e] ? (int)0x80000000 : 0);
    const int n = min(32, ee - base);
    for (int j = 0; j < n; ++j) {
      const int s = __shfl_sync(FULL_MASK, my, j);
      const int src = s & 0x7fffffff;
      if (src == quirk) continue;  // the bipartite edge SAGEConv's remove_self_loops drops (gnn_model.py:389-390)
      add_sel(at, ad, __ldg(reinterpret_cast<const float2*>(h_in + (size_t)src * F) + lane), s);
    }
  }
  s_part[warp][lane] = make_float4(at.x, at.y, ad.x, ad.y);
  __syncthreads();
  if (warp == 0) {
    float4 t = s_part[0][lane];
    for (int w = 1; w < CT / 32; ++w) {
      const float4 o = s_part[w][lane];
      t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
    }
    s_a[2 * lane] = t.x; s_a[2 * lane + 1] = t.y;
    s_a[F + 2 * lane] = t.z; s_a[F + 2 * lane + 1] = t.w;
    const float2 ha = *reinterpret_cast<const float2*>(emb_a + (size_t)g * emb_ld + layer * F + 2 * lane);
    s_a[2 * F + 2 * lane] = ha.x; s_a[2 * F + 2 * lane + 1] = ha.y;
  }
  __syncthreads();
  if (tid < F) {
    float acc = bias_a[tid];
#pragma unroll 8
    for (int k = 0; k < 3 * F; ++k) acc = fmaf(s_a[k], Wa[k * F + tid], acc);
    emb_a[(size_t)g * emb_ld + (layer + 1) * F + tid] = fmaxf(acc, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// global_add_pool of one layer (gnn_model.py:107), deterministic: a warp sums a sub-chunk of 256 consecutive packed rows
// into one partial per neighborhood it meets (slot = sub-chunk + neighborhood: unique), then one warp per neighborhood
// adds its partials in sub-chunk order.  Canonical rows are zero in h, so they need no special case.
// ------------------------------------------------------------------------------------------------------------------
constexpr int PSUB = 256;
__global__ void shmp_mt_pool_partial_kernel(const int32_t* __restrict__ row_nbh, const float* __restrict__ h, int V,
                                            float* __restrict__ partial) {
  const long long sc = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long r0 = sc * PSUB;
  if (r0 >= V) return;
  const int lane = lane_id();
  const int r1 = (int)min((long long)V, r0 + PSUB);
  int g = row_nbh[r0];
  float2 acc = make_float2(0.f, 0.f);
  for (int r = (int)r0; r < r1; r += 4) {
    float2 v[4];
    int gg[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int rr = min(r + u, r1 - 1);
      v[u] = __ldg(reinterpret_cast<const float2*>(h + (size_t)rr * F) + lane);
      gg[u] = row_nbh[rr];
      if (r + u >= r1) v[u] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (gg[u] != g) {
        *reinterpret_cast<float2*>(partial + ((size_t)sc + g) * F + 2 * lane) = acc;
        acc = make_float2(0.f, 0.f);
        g = gg[u];
      }
      acc.x += v[u].x; acc.y += v[u].y;
    }
  }
  *reinterpret_cast<float2*>(partial + ((size_t)sc + g) * F + 2 * lane) = acc;
}

__global__ void shmp_mt_pool_reduce_kernel(const int32_t* __restrict__ nbh_ptr, int G, const float* __restrict__ partial,
                                           float* __restrict__ pool, int emb_ld, int layer) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const int lane = lane_id();
  const int lo = nbh_ptr[g], hi = nbh_ptr[g + 1];
  float2 acc = make_float2(0.f, 0.f);
  if (hi > lo)
    for (int sc = lo / PSUB; sc <= (hi - 1) / PSUB; ++sc) {
      const float2 v = *reinterpret_cast<const float2*>(partial + ((size_t)sc + g) * F + 2 * lane);
      acc.x += v.x; acc.y += v.y;
    }
  *reinterpret_cast<float2*>(pool + (size_t)g * emb_ld + layer * F + 2 * lane) = acc;
}

__global__ void shmp_mt_zero_canon_kernel(const int32_t* __restrict__ nbh_ptr, int G, float* __restrict__ hA,
                                          float* __restrict__ hB) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= G) return;
  const size_t row = (size_t)nbh_ptr[g + 1] - 1;
  reinterpret_cast<float2*>(hA + row * F)[lane_id()] = make_float2(0.f, 0.f);
  reinterpret_cast<float2*>(hB + row * F)[lane_id()] = make_float2(0.f, 0.f);
}

}  // namespace

int64_t desco_internal_shmp_mt_workspace_bytes(int num_rows, int num_neighborhoods) {
  return ((int64_t)((num_rows + PSUB - 1) / PSUB + num_neighborhoods + 1) * F * 4 + 255) / 256 * 256;
}

// Layers + pooling for a packed batch whose plan (row_nbh, crow, canon_code, quirk_row) and layer-0 features (hA count
// rows, emb_a[:, 0:F]) are ready.  w_layers: the fp32 blob of desco_shmp_forward (bias_c, Cw, Wa, bias_a are read from
// it); w_layers_mt: per layer KB x [hi | lo] images of the Wc^T blocks (SHMP_MT_LAYER_BYTES).  Leaves pool / emb_a
// ([G][emb_ld]) like the other paths.
int desco_internal_shmp_mt_layers(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                                  const uint8_t* edge_tri, int G, int V, int hetero, const int32_t* row_nbh,
                                  const int32_t* crow, const uint8_t* canon_code, const int32_t* quirk_row, float* hA,
                                  float* hB, float* emb_a, float* pool, float* cvec, int emb_ld, const float* w_layers,
                                  int64_t layer_floats, const void* w_layers_mt, int layers, int passes, void* workspace,
                                  int32_t* status, int anchored, cudaStream_t s) {
  const int Vc = hetero ? V - G : V;
  const int KC = 3 * F;
  float* partial = (float*)workspace;
  static bool attr_set = false;
  if (!attr_set) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(shmp_mt_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  if (hetero) {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
    shmp_mt_zero_canon_kernel<<<(G * 32 + 255) / 256, 256, 0, s>>>(nbh_ptr, G, hA, hB);
    DESCO_LAUNCH_CHECK();
  }
  auto pool_layer = [&](const float* h, int l) -> int {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s, 2);
    const long long warps = ((long long)V + PSUB - 1) / PSUB;
    shmp_mt_pool_partial_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(row_nbh, h, V, partial);
    shmp_mt_pool_reduce_kernel<<<(G * 32 + 255) / 256, 256, 0, s>>>(nbh_ptr, G, partial, pool, emb_ld, l);
    if (anchored) desco_internal_shmp_copy_last_rows(nbh_ptr, G, h, l, emb_a, emb_ld, s);  // homogeneous model: centre rows
    DESCO_LAUNCH_CHECK();
    return DESCO_OK;
  };
  const int num_tiles = (Vc + TR - 1) / TR;
  const int sms = desco_num_sms();
  float* h_in = hA;
  float* h_out = hB;
  for (int l = 0; l < layers; ++l) {
    const float* wl = w_layers + (size_t)l * layer_floats;  // [Wc | bias_c | Cw | Wa | bias_a]
    const float* bias_c = wl + KC * F;
    const float* Cw = bias_c + F;
    const float* Wa = Cw + F * 2 * F;
    const float* bias_a = Wa + KC * F;
    int rc = pool_layer(h_in, l);
    if (rc) return rc;
    if (hetero) {
      DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s, 2);
      desco_internal_shmp_cvec(emb_a, emb_ld, l, Cw, G, cvec, s);
      shmp_mt_canon_kernel<<<G, CT, 0, s>>>(nbh_ptr, edge_ptr, edge_col, edge_tri, quirk_row, h_in, emb_a, emb_ld, l, Wa,
                                            bias_a);
      DESCO_LAUNCH_CHECK();
    }
    if (num_tiles > 0) {
      MtArgs a;
      a.edge_ptr = edge_ptr; a.edge_col = edge_col; a.edge_tri = edge_tri;
      a.row_nbh = row_nbh; a.crow = crow; a.canon_code = canon_code;
      a.Vc = Vc; a.num_tiles = num_tiles; a.passes = passes;
      a.h_in = h_in; a.h_out = h_out; a.cvec = hetero ? cvec : nullptr;
      a.w_img = (const uint8_t*)w_layers_mt + (size_t)l * SHMP_MT_LAYER_BYTES;
      a.bias_c = bias_c; a.status = status;
      DescoProfScope prof(DESCO_PROF_SHMP_LAYER, s);
      shmp_mt_layer_kernel<<<num_tiles < sms ? num_tiles : sms, THREADS, SMEM_BYTES, s>>>(a);
      DESCO_LAUNCH_CHECK();
    }
    float* t = h_in; h_in = h_out; h_out = t;
  }
  return pool_layer(h_in, layers);
}
