// SHMP neighborhood counting, all message-passing layers fused in ONE persistent sm_100a kernel (tcgen05 + TMEM).
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/gnn_model.py:362-404  SAGEConv           sum-aggregate then Linear, per relation
//   subgraph_counting/gnn_model.py:230-277  BaseGNNCore.forward as expanded by to_hetero_old (lightning_model.py:371-421)
//   subgraph_counting/gnn_model.py:88-89,107  skip-concat + global_add_pool (written as per-layer pooled sums)
//
// Canonical neighborhoods are independent small graphs, so a CTA takes a TILE = a run of whole neighborhoods with at
// most 128 rows and keeps its features on chip for all layers (the reference round-trips [V,64] through HBM twice per
// relation per layer).  Per layer, with h the 128 x 64 tile of layer inputs:
//   1. transform-then-aggregate (SAGEConv is linear, so sum_j (h_j W) == (sum_j h_j) W):
//        P = h . [W_tri | W_tride | W_self]           128 x 64 x 192 GEMM on tcgen05, bf16 hi/lo operand split
//      (three passes hi.hi + lo.hi + hi.lo, fp32 accumulation in TMEM: ~3e-6 from the fp32 oracle), weights fetched
//      with one bulk async copy per image; the next layer's weights stream in while this layer's epilogue runs;
//   2. while the tensor pipe works, the CUDA cores do the few canonical rows of the tile in fp32:
//        h_a' = relu([sum_tri h_j | sum_tride h_j | h_a] . Wa + b_a),   cvec = h_a . [Cw_tri | Cw_tride];
//   3. TMEM -> shared memory, then the segmented, edge-type-split gather runs out of shared memory:
//        h_i' = relu(P_self[i] + sum_{j in N_tri(i)} P_tri[j] + sum_{j in N_tride(i)} P_tride[j] + cvec[type(i,a)] + b);
//   4. h' is written back as the next layer's bf16 hi/lo A operand (swizzled) and as fp32 rows for pooling.
// Only the pooled sums and the canonical rows ([G, (L+1)*64] each) ever leave the chip.
#include "common.cuh"
#include "tc05.cuh"
#include "shmp_internal.h"
#include "../../include/desco_b200.h"

namespace {

constexpr int F = 64;
constexpr int TR = SHMP_TILE_ROWS;      // rows per tile (one UMMA M)
constexpr int MAXC = SHMP_TILE_MAX_NBH; // neighborhoods per tile
constexpr int NB = 3 * F;               // GEMM N: [tri | tride | self]
constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int ISSUER = THREADS - 32;      // thread that talks to the tensor pipe / TMA engine (lane 0 of the last warp)
constexpr int LDP = 2 * F + 4;          // row pitch (floats) of the P_tri|P_tride buffer
constexpr int LDS_ = F + 4;             // row pitch (floats) of the fp32 staging rows (P_self, then h')
constexpr int EDGE_CAP = 8192;          // staged edges per tile (1 byte each); larger tiles read edges from global
constexpr int CH = SHMP_PLAN_CHUNK;     // neighborhoods per planning chunk
constexpr int CINP = NB + 8;            // row pitch (bf16) of the canonical-input rows: fragment loads of 8 rows hit 32 banks

// per-layer weight blob (bytes)
constexpr int OFF_BHI = 0;
constexpr int OFF_BIASC = 2 * NB * 128;
constexpr int OFF_BIASA = OFF_BIASC + F * 4;
// canonical-row weights as ready-made mma.sync B fragments (tcpack.pack_mma_b_frags): per (column tile of 8, k-step of
// 16, lane) one uint4 {hi(k..k+1), hi(k+8..k+9), lo(k..k+1), lo(k+8..k+9)} of column 8 nt + lane/4, k = 16 ks + 2 (lane%4)
constexpr int OFF_WAT = OFF_BIASA + F * 4;        // Wa  [192 k][64 n]  -> [8 nt][12 ks][32 lanes] uint4
constexpr int OFF_CWT = OFF_WAT + F * NB * 4;     // Cw  [64 k][128 n]  -> [16 nt][4 ks][32 lanes] uint4
constexpr int LAYER_BYTES = OFF_CWT + 2 * F * F * 4;
static_assert(LAYER_BYTES == SHMP_TC_LAYER_BYTES, "blob layout and header constant disagree");

// shared-memory carve-up (bytes from the 1024-aligned base)
constexpr int SM_BHI = 0;
constexpr int SM_BLO = SM_BHI + NB * 128;
constexpr int SM_AHI = SM_BLO + NB * 128;
constexpr int SM_ALO = SM_AHI + TR * 128;
constexpr int SM_P = SM_ALO + TR * 128;
constexpr int SM_STAGE = SM_P + TR * LDP * 4;
constexpr int SM_CIN = SM_STAGE + TR * LDS_ * 4;        // bf16 hi then lo, each [MAXC][CINP]: sum_tri h_j | sum_tride h_j | h_a
constexpr int SM_CVEC = SM_CIN + 2 * MAXC * CINP * 2;   // [MAXC][128]
constexpr int SM_CH = SM_CVEC + MAXC * 2 * F * 4;       // [MAXC][64]  h_a of the next layer
constexpr int SM_EDGE = SM_CH + MAXC * F * 4;           // [EDGE_CAP] local col | tri << 7
constexpr int SM_EPTR = SM_EDGE + EDGE_CAP;             // [TR + 1] int
constexpr int SM_ROWG = SM_EPTR + ((TR + 1) * 4 + 15) / 16 * 16;  // [TR] uint8 local neighborhood of the row
constexpr int SM_ROWCODE = SM_ROWG + TR;                // [TR] uint8: 0 none, 1 tri / 2 tride edge to canonical, 3 canonical
constexpr int SM_NBHLO = SM_ROWCODE + TR;               // [MAXC + 1] int local first row
constexpr int SM_QUIRK = SM_NBHLO + ((MAXC + 1) * 4 + 15) / 16 * 16;  // [MAXC] int local row or -1
constexpr int SM_BIASC = SM_QUIRK + (MAXC * 4 + 15) / 16 * 16;        // [64] float: bias of the count rows, this layer
constexpr int SM_BARS = SM_BIASC + F * 4;                             // 2 mbarriers + tmem slot
constexpr int SM_TOTAL = SM_BARS + 64;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;  // slack for the manual 1024-B alignment
static_assert(MAXC <= 32 && MAXC * CINP * 2 + MAXC * 3 * F * 4 >= 32 * CINP * 2, "the second mma row block reads (and ignores) rows past MAXC");
static_assert(SMEM_BYTES <= 232448, "fused SHMP kernel exceeds the 227 KB shared-memory limit");
static_assert(SM_AHI % 1024 == 0 && SM_ALO % 1024 == 0 && SM_BLO % 1024 == 0, "UMMA tiles must be 1024-B aligned");

// ------------------------------------------------------------------------------------------------------------------
// tile plan: greedy packing of consecutive neighborhoods into tiles (<= TR rows, <= MAXC neighborhoods), per chunk of
// CH neighborhoods, by pointer doubling over jump[g] = first neighborhood of the tile after the one starting at g.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) shmp_tile_plan_kernel(const int32_t* __restrict__ nbh_ptr, int G,
                                                              const int32_t* __restrict__ g_dev,
                                                              int32_t* __restrict__ tile_start,
                                                              int32_t* __restrict__ tile_count, int32_t* __restrict__ ticket,
                                                              int32_t* __restrict__ status) {
  __shared__ int sP[CH + 1], sJa[CH + 1], sJb[CH + 1], sS[CH + 1];
  __shared__ int s_cnt;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * CH;
  if (g_dev) G = min(G, *g_dev);  // stream-ordered form: G is a capacity, the batch's own count lives on the device
  const int n = min(CH, G - c0);
  if (n <= 0) {
    if (tid == 0) {
      tile_count[blockIdx.x] = 0;
      if (blockIdx.x == 0) *ticket = 0;
    }
    return;
  }
  for (int i = tid; i <= n; i += 1024) sP[i] = nbh_ptr[c0 + i];
  if (tid == 0) {
    s_cnt = n;
    sS[0] = 0;
  }
  __syncthreads();
  for (int g = tid; g <= n; g += 1024) {
    int j = n;
    if (g < n) {
      int lo = g + 1, hi = min(n, g + MAXC);
      if (sP[lo] - sP[g] > TR) {
        atomicExch(status, DESCO_ERANGE);  // a neighborhood larger than a tile: the caller must use the layered path
      } else {
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (sP[mid] - sP[g] <= TR) lo = mid; else hi = mid - 1;
        }
      }
      j = lo;
    }
    sJa[g] = j;
  }
  __syncthreads();
  int* J = sJa;
  int* Jn = sJb;
  for (int len = 1; len <= n; len <<= 1) {  // invariant: sS[t] = jump^t(0) for t < len, J = jump^len
    for (int t = tid; t < len && len + t <= n; t += 1024) sS[len + t] = J[sS[t]];
    for (int g = tid; g <= n; g += 1024) Jn[g] = J[J[g]];
    __syncthreads();
    int* tmp = J; J = Jn; Jn = tmp;
  }
  for (int t = tid; t <= n; t += 1024)
    if (sS[t] == n) atomicMin(&s_cnt, t);
  __syncthreads();
  const int count = (n == 0) ? 0 : s_cnt;
  int32_t* out = tile_start + (size_t)blockIdx.x * (CH + 1);
  for (int t = tid; t <= count; t += 1024) out[t] = c0 + sS[t];
  if (tid == 0) tile_count[blockIdx.x] = count;
  if (tid == 0 && blockIdx.x == 0) *ticket = 0;  // the fused kernel deals tiles through this counter
}

// ------------------------------------------------------------------------------------------------------------------
// fused layers
// ------------------------------------------------------------------------------------------------------------------
// phase timing (thread 0 of every CTA accumulates its own clock64 deltas; read with desco_shmp_fused_phase_cycles)
enum { PH_SETUP = 0, PH_POOL, PH_ISSUE, PH_CANON, PH_WAIT_MMA, PH_T2S, PH_GATHER, PH_COUNT };
__device__ unsigned long long g_phase_cycles[PH_COUNT];

struct FusedArgs {
  const int32_t* nbh_ptr; const int32_t* edge_ptr; const int32_t* edge_col; const uint8_t* edge_tri;
  const int32_t* tile_start; const int32_t* tile_count; int32_t* ticket;
  const int32_t* g_dev;     // device-resident neighborhood count (stream-ordered form) or NULL
  int G, num_chunks, pyg_batch_size, layers, passes, input_dim, emb_ld;
  const float* feat;        // [V][input_dim] or NULL (ZeroNodeFeat)
  const float* w_pre;       // per node type (count, canonical): W[input_dim][64], b[64]
  const uint8_t* w_layers;  // layers x LAYER_BYTES
  float* emb_a;             // [G][emb_ld] canonical rows, all layers
  uint8_t* emb_img;         // the same rows as ready-made tensor-core A operands of the anchor GEMM (or NULL): per block of
                            // 128 neighborhoods and layer one 64-wide K atom, bf16 hi | mid | lo SWIZZLE_128B images
  float* pool;              // [G][emb_ld] sum over count rows, all layers
  int32_t* status;
};

// D(16x8, fp32) += A(16x16, bf16, row) . B(16x8, bf16, col) on the warp-level tensor path
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// a += b as two packed fp32x2 adds (sm_100 FADD2: half the issue slots of four scalar adds; same rounding)
__device__ __forceinline__ void add4(float4& a, const float4 b) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%6, %7};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%2, %3}, ra;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
}
// four fp32 -> bf16 hi / lo pairs with the packed converts (x = hi + lo + O(2^-17 |x|), as tc05::split_bf16)
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
  hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

__global__ void __launch_bounds__(THREADS, 1) shmp_fused_kernel(const FusedArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment computed as an OFFSET from the __shared__ symbol, so the compiler keeps the shared address space
  // (LDS/STS); casting through uintptr_t would turn every access into a generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sBhi = smem + SM_BHI;
  uint8_t* sAhi = smem + SM_AHI;
  uint8_t* sAlo = smem + SM_ALO;
  float* sP = reinterpret_cast<float*>(smem + SM_P);
  float* sStage = reinterpret_cast<float*>(smem + SM_STAGE);
  __nv_bfloat16* sCinHi = reinterpret_cast<__nv_bfloat16*>(smem + SM_CIN);
  __nv_bfloat16* sCinLo = sCinHi + MAXC * CINP;
  float* sCvec = reinterpret_cast<float*>(smem + SM_CVEC);
  float* sCh = reinterpret_cast<float*>(smem + SM_CH);
  uint8_t* sEdge = smem + SM_EDGE;
  int* sEptr = reinterpret_cast<int*>(smem + SM_EPTR);
  uint8_t* sRowG = smem + SM_ROWG;
  uint8_t* sRowCode = smem + SM_ROWCODE;
  int* sNbhLo = reinterpret_cast<int*>(smem + SM_NBHLO);
  int* sQuirk = reinterpret_cast<int*>(smem + SM_QUIRK);
  float* sBiasC = reinterpret_cast<float*>(smem + SM_BIASC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BARS);  // [0] weights landed, [1] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // tiles are dealt dynamically (one atomic ticket per tile over the flattened (chunk, tile) list): tile costs vary by
  // ~2x with the number of neighborhoods / edges in the tile, and a static round-robin left the slowest CTA 1.5x behind
  __shared__ int s_ticket;
  int total_tiles = 0;
  for (int c = 0; c < p.num_chunks; ++c) total_tiles += p.tile_count[c];
  if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);
  __syncthreads();
  int cur_tile = s_ticket;
  if (cur_tile >= total_tiles) return;

  if (tid == 0) {
    tc05::mbar_init(&bars[0], 1);
    tc05::mbar_init(&bars[1], 1);
    tc05::fence_mbar_init();
  }
  if (warp == 0) tc05::tmem_alloc(tmem_slot, 256);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t layer_image_bytes = 2 * NB * 128;
  if (tid == ISSUER) {  // weights of layer 0 for the first tile
    tc05::mbar_arrive_expect_tx(&bars[0], layer_image_bytes);
    tc05::bulk_g2s(sBhi, p.w_layers + OFF_BHI, layer_image_bytes, &bars[0]);
  }
  bool copy_pending = true;  // meaningful in the ISSUER thread only
  uint32_t wphase = 0, mphase = 0;
  bool timed_out = false;
  const uint32_t idesc = tc05::make_idesc_bf16(TR, NB);
  const int hw = lane >> 4, hl = lane & 15;  // half-warp id / lane inside the half-warp (one half-warp per row)

  while (cur_tile < total_tiles) {
    {
      int c = 0, t = cur_tile;
      while (t >= p.tile_count[c]) t -= p.tile_count[c++];
      const int32_t* ts = p.tile_start + (size_t)c * (CH + 1);
      const int nb0 = ts[t], nb1 = ts[t + 1];
      __syncthreads();  // previous tile fully retired (s_ticket included)
      if (tid == 0) s_ticket = atomicAdd(p.ticket, 1);  // the ticket of the NEXT tile: known before the last layer's prefetch
      const int nc = nb1 - nb0;                    // neighborhoods in the tile
      const int row0 = p.nbh_ptr[nb0];
      const int R = p.nbh_ptr[nb1] - row0;         // rows in the tile
      const int e0 = p.edge_ptr[row0];
      const int Et = p.edge_ptr[row0 + R] - e0;
      const bool edges_staged = Et <= EDGE_CAP;
      if (R > TR || nc > MAXC) {  // the plan kernel has already raised DESCO_ERANGE; never overrun the tile buffers
        __syncthreads();
        cur_tile = s_ticket;
        continue;
      }

      // ---------------- tile setup ----------------
      long long tick = clock64();
      auto lap = [&](int phase) {
        if (tid == 0) {
          const long long now = clock64();
          atomicAdd(&g_phase_cycles[phase], (unsigned long long)(now - tick));
          tick = now;
        }
      };
      if (tid <= nc) sNbhLo[tid] = p.nbh_ptr[nb0 + tid] - row0;
      if (tid < nc) {
        // SAGEConv.forward runs remove_self_loops on the bipartite count<->canonical relations (gnn_model.py:389-390):
        // see shmp.cu shmp_plan_kernel / DESIGN.md "reference quirks"
        const int g = nb0 + tid;
        const int bs = p.pyg_batch_size > 0 ? p.pyg_batch_size : (p.g_dev ? *p.g_dev : p.G);
        const int g0 = (g / bs) * bs;
        const int lo = p.nbh_ptr[g];
        sQuirk[tid] = (p.pyg_batch_size >= 0 && lo - p.nbh_ptr[g0] == 2 * (g - g0)) ? (lo - row0) : -1;  // < 0: quirk off
      }
      for (int r = tid; r <= R; r += THREADS) sEptr[r] = p.edge_ptr[row0 + r] - e0;
      if (edges_staged)
        for (int e = tid; e < Et; e += THREADS)
          sEdge[e] = (uint8_t)((p.edge_col[e0 + e] - row0) | (p.edge_tri[e0 + e] ? 0x80 : 0));
      for (int i = tid; i < 2 * TR * 128 / 16; i += THREADS)  // zero both A images (rows >= R and canonical rows stay 0)
        reinterpret_cast<uint4*>(sAhi)[i] = make_uint4(0u, 0u, 0u, 0u);
      __syncthreads();
      auto edge_at = [&](int e) -> int {
        if (edges_staged) return sEdge[e];
        return (p.edge_col[e0 + e] - row0) | (p.edge_tri[e0 + e] ? 0x80 : 0);
      };
      if (tid < R) {
        int a = 0, b = nc;  // largest a with sNbhLo[a] <= tid
        while (b - a > 1) {
          const int mid = (a + b) >> 1;
          if (sNbhLo[mid] <= tid) a = mid; else b = mid;
        }
        const int canon = sNbhLo[a + 1] - 1;
        int code = 0;
        if (tid == canon) {
          code = 3;
        } else {
          const int eb = sEptr[tid], ee = sEptr[tid + 1];
          if (ee > eb && tid != sQuirk[a]) {
            const int last = edge_at(ee - 1);  // canonical = max row of its neighborhood = last entry of a sorted row
            if ((last & 127) == canon) code = (last & 0x80) ? 1 : 2;
          }
        }
        sRowG[tid] = (uint8_t)a;
        sRowCode[tid] = (uint8_t)code;
      }
      __syncthreads();

      // ---------------- layer-0 inputs: h0 = feat . Wpre + bpre per node type (gnn_model.py:231) ----------------
      {
        const float* Wc = p.w_pre;                                   // count:     [input_dim][64] then bias[64]
        const float* Wn = p.w_pre + (size_t)(p.input_dim + 1) * F;   // canonical
        for (int r = warp * 2 + hw; r < R; r += 2 * NWARPS) {
          const bool canon = sRowCode[r] == 3;
          const float* W = canon ? Wn : Wc;
          float4 v = *reinterpret_cast<const float4*>(W + (size_t)p.input_dim * F + 4 * hl);
          if (p.feat) {
            for (int d = 0; d < p.input_dim; ++d) {
              const float x = p.feat[(size_t)(row0 + r) * p.input_dim + d];
              const float4 w = *reinterpret_cast<const float4*>(W + (size_t)d * F + 4 * hl);
              v.x = fmaf(x, w.x, v.x); v.y = fmaf(x, w.y, v.y); v.z = fmaf(x, w.z, v.z); v.w = fmaf(x, w.w, v.w);
            }
          }
          if (canon) {
            *reinterpret_cast<float4*>(sCh + sRowG[r] * F + 4 * hl) = v;
          } else {
            *reinterpret_cast<float4*>(sStage + r * LDS_ + 4 * hl) = v;
            __align__(8) __nv_bfloat16 hi[4], lo[4];
            tc05::split_bf16(v.x, hi[0], lo[0]); tc05::split_bf16(v.y, hi[1], lo[1]);
            tc05::split_bf16(v.z, hi[2], lo[2]); tc05::split_bf16(v.w, hi[3], lo[3]);
            const uint32_t off = tc05::sw128_offset(r, 4 * hl);
            *reinterpret_cast<uint2*>(sAhi + off) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sAlo + off) = *reinterpret_cast<const uint2*>(lo);
          }
        }
      }
      tc05::fence_proxy_async_smem();  // A images written through the generic proxy -> visible to tcgen05.mma
      __syncthreads();
      lap(PH_SETUP);

      for (int l = 0; l <= p.layers; ++l) {
        // ------------ tensor pipe: P = h . [W_tri | W_tride | W_self], issued first so that it runs under the pool
        // and canonical phases (one thread of the last warp issues: that warp is the least loaded in the pool phase)
        if (l < p.layers && tid == ISSUER) {
          if (!tc05::mbar_wait(&bars[0], wphase)) timed_out = true;
          copy_pending = false;
          tc05::fence_after_sync();
          const uint64_t dAhi = tc05::make_smem_desc(sAhi), dAlo = tc05::make_smem_desc(sAlo);
          const uint64_t dBhi = tc05::make_smem_desc(sBhi), dBlo = tc05::make_smem_desc(smem + SM_BLO);
          bool acc = false;
          for (int pass = 0; pass < p.passes; ++pass) {
            const uint64_t da = (pass == 1) ? dAlo : dAhi;
            const uint64_t db = (pass == 2) ? dBlo : dBhi;
#pragma unroll
            for (int k = 0; k < F / 16; ++k) {
              tc05::mma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, acc);
              acc = true;
            }
          }
          tc05::mma_commit(&bars[1]);
        }
        if (l < p.layers) wphase ^= 1;
        // this warp's B fragments of the canonical-row weights of layer l: issued now, consumed after the pool phase.
        // Warps 0-7: z_a column tile `warp` (K = 192, 12 k-steps); warps 8-15: cvec column tiles 2(warp-8), 2(warp-8)+1
        // (K = 64, 4 k-steps each).
        const uint8_t* wl = p.w_layers + (size_t)(l < p.layers ? l : 0) * LAYER_BYTES;
        const int fg = lane >> 2, ft = lane & 3;  // mma fragment coordinates: group (row / column), thread in group
        uint4 wf[12];
        // biases of this layer, fetched here too so that their L2 latency hides under the pool phase
        float2 bias_a = make_float2(0.f, 0.f);
        if (l < p.layers) {
          if (tid < F) sBiasC[tid] = __ldg(reinterpret_cast<const float*>(wl + OFF_BIASC) + tid);  // read after two barriers
          if (warp < 8) bias_a = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(wl + OFF_BIASA) + 8 * warp + 2 * ft));
          if (warp < 8) {
            const uint4* W = reinterpret_cast<const uint4*>(wl + OFF_WAT) + (size_t)warp * 12 * 32 + lane;
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) wf[ks] = __ldg(W + 32 * ks);
          } else {
            const uint4* W = reinterpret_cast<const uint4*>(wl + OFF_CWT) + (size_t)(2 * (warp - 8)) * 4 * 32 + lane;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) wf[ks] = __ldg(W + 32 * ks);
          }
        }
        // ------------ pool + canonical inputs of layer l (from the fp32 rows of h^l in sStage, h_a^l in sCh) ------------
        // thread = (neighborhood, feature): 64 consecutive features per neighborhood, 8 neighborhoods at a time
        {
          const int f = tid & (F - 1);
          for (int i = tid >> 6; i < nc; i += THREADS / F) {
            const int lo = sNbhLo[i], canon = sNbhLo[i + 1] - 1;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            int r = lo;
            for (; r + 3 < canon; r += 4) {
              s0 += sStage[r * LDS_ + f];
              s1 += sStage[(r + 1) * LDS_ + f];
              s2 += sStage[(r + 2) * LDS_ + f];
              s3 += sStage[(r + 3) * LDS_ + f];
            }
            for (; r < canon; ++r) s0 += sStage[r * LDS_ + f];
            const size_t gofs = (size_t)(nb0 + i) * p.emb_ld + (size_t)l * F + f;
            p.pool[gofs] = (s0 + s1) + (s2 + s3);          // global_add_pool, count rows (gnn_model.py:107)
            const float ha = sCh[i * F + f];
            if (p.emb_img) {  // skip-concat of the canonical row (:275), split three ways for csrc/dense_tc.cu
              const int g = nb0 + i;
              uint8_t* img = p.emb_img + ((size_t)(g >> 7) * (p.layers + 1) + l) * (3 * TR * 128) + tc05::sw128_offset(g & 127, f);
              const __nv_bfloat16 hi = __float2bfloat16_rn(ha);
              const float r1 = ha - __bfloat162float(hi);
              const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
              *reinterpret_cast<__nv_bfloat16*>(img) = hi;
              *reinterpret_cast<__nv_bfloat16*>(img + TR * 128) = mid;
              *reinterpret_cast<__nv_bfloat16*>(img + 2 * TR * 128) = __float2bfloat16_rn(r1 - __bfloat162float(mid));
            } else {
              p.emb_a[gofs] = ha;
            }
            if (l < p.layers) {
              float vt = 0.f, vd = 0.f;
              const int quirk = sQuirk[i];
              for (int e = sEptr[canon], ee = sEptr[canon + 1]; e < ee; ++e) {
                const int b = edge_at(e);
                const int j = b & 127;
                const float v = (j == quirk) ? 0.f : sStage[j * LDS_ + f];
                if (b & 0x80) vt += v; else vd += v;
              }
              tc05::split_bf16(vt, sCinHi[i * CINP + f], sCinLo[i * CINP + f]);
              tc05::split_bf16(vd, sCinHi[i * CINP + F + f], sCinLo[i * CINP + F + f]);
              tc05::split_bf16(ha, sCinHi[i * CINP + 2 * F + f], sCinLo[i * CINP + 2 * F + f]);
            }
          }
        }
        if (l == p.layers) break;
        __syncthreads();
        lap(PH_POOL);

        // ------------ canonical rows on the warp-level tensor path (mma.sync m16n8k16, bf16 hi/lo split, 3 passes:
        // same fp32-grade products as the tile GEMM); 16 canonical rows per row block ------------
        for (int mb = 0; mb < nc; mb += 16) {
          // A fragment registers: rows fg / fg + 8 of the block, bf16 pairs at k = 16 ks + 2 ft (+ 8)
          const uint32_t* xh0 = reinterpret_cast<const uint32_t*>(sCinHi + (mb + fg) * CINP + 2 * ft);
          const uint32_t* xl0 = reinterpret_cast<const uint32_t*>(sCinLo + (mb + fg) * CINP + 2 * ft);
          constexpr int R8 = 8 * CINP / 2, K8 = 4;  // +8 rows / +8 columns in 32-bit words
          if (warp < 8) {  // z_a = [sum_tri | sum_tride | h_a] . Wa, columns 8 warp .. 8 warp + 7
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
              const uint32_t ahi[4] = {xh0[8 * ks], xh0[8 * ks + R8], xh0[8 * ks + K8], xh0[8 * ks + R8 + K8]};
              const uint32_t alo[4] = {xl0[8 * ks], xl0[8 * ks + R8], xl0[8 * ks + K8], xl0[8 * ks + R8 + K8]};
              mma16816(acc[ks & 1], ahi, wf[ks].x, wf[ks].y);
              mma16816(acc[ks & 1], alo, wf[ks].x, wf[ks].y);
              mma16816(acc[ks & 1], ahi, wf[ks].z, wf[ks].w);
            }
            const int n = 8 * warp + 2 * ft;
            const float2 b = bias_a;
            const int r0 = mb + fg, r1 = r0 + 8;  // h_a^{l+1}; read again only after the next barriers
            if (r0 < nc)
              *reinterpret_cast<float2*>(sCh + r0 * F + n) =
                  make_float2(fmaxf(acc[0][0] + acc[1][0] + b.x, 0.f), fmaxf(acc[0][1] + acc[1][1] + b.y, 0.f));
            if (r1 < nc)
              *reinterpret_cast<float2*>(sCh + r1 * F + n) =
                  make_float2(fmaxf(acc[0][2] + acc[1][2] + b.x, 0.f), fmaxf(acc[0][3] + acc[1][3] + b.y, 0.f));
          } else {  // cvec = h_a . [Cw_tri | Cw_tride], column tiles 2(warp-8), 2(warp-8)+1
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int w0 = F + 8 * ks;  // the h_a block starts at bf16 column 2F = word F
              const uint32_t ahi[4] = {xh0[w0], xh0[w0 + R8], xh0[w0 + K8], xh0[w0 + R8 + K8]};
              const uint32_t alo[4] = {xl0[w0], xl0[w0 + R8], xl0[w0 + K8], xl0[w0 + R8 + K8]};
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                mma16816(acc[j], ahi, wf[4 * j + ks].x, wf[4 * j + ks].y);
                mma16816(acc[j], alo, wf[4 * j + ks].x, wf[4 * j + ks].y);
                mma16816(acc[j], ahi, wf[4 * j + ks].z, wf[4 * j + ks].w);
              }
            }
            const int r0 = mb + fg, r1 = r0 + 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int n = 8 * (2 * (warp - 8) + j) + 2 * ft;
              if (r0 < nc) *reinterpret_cast<float2*>(sCvec + r0 * 2 * F + n) = make_float2(acc[j][0], acc[j][1]);
              if (r1 < nc) *reinterpret_cast<float2*>(sCvec + r1 * 2 * F + n) = make_float2(acc[j][2], acc[j][3]);
            }
          }
        }

        lap(PH_CANON);
        // ------------ accumulator: TMEM -> shared memory ------------
        if (!tc05::mbar_wait(&bars[1], mphase)) timed_out = true;
        lap(PH_WAIT_MMA);
        mphase ^= 1;
        tc05::fence_after_sync();
        if (tid == ISSUER) {  // the B images are free again: stream in the next layer's (or the next tile's layer-0) weights
          const bool last = (s_ticket >= total_tiles) && (l + 1 == p.layers);
          if (!last) {
            const int nl = (l + 1) % p.layers;
            tc05::mbar_arrive_expect_tx(&bars[0], layer_image_bytes);
            tc05::bulk_g2s(sBhi, p.w_layers + (size_t)nl * LAYER_BYTES + OFF_BHI, layer_image_bytes, &bars[0]);
            copy_pending = true;
          }
        }
        {
          const int q = warp & 3, cg = warp >> 2;  // TMEM lane quarter of this warp, column group of 48
          const int r = 32 * q + lane;
          uint32_t v[3][16];  // the three loads of this warp's 48 columns are in flight together
#pragma unroll
          for (int j = 0; j < 3; ++j) tc05::tmem_ld16_async(tmem + (static_cast<uint32_t>(32 * q) << 16) + cg * 48 + j * 16, v[j]);
          tc05::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const int c0 = cg * 48 + j * 16;
            if (r < R) {
              float* dst = (c0 < 2 * F) ? (sP + r * LDP + c0) : (sStage + r * LDS_ + (c0 - 2 * F));
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<uint4*>(dst + i) = make_uint4(v[j][i], v[j][i + 1], v[j][i + 2], v[j][i + 3]);
            }
          }
        }
        tc05::fence_before_sync();
        __syncthreads();
        lap(PH_T2S);

        // ------------ segmented, edge-type-split gather out of shared memory; one quarter-warp per row ------------
        // lane q of a quarter owns features 4q..4q+3 and 32+4q..32+4q+3 (two conflict-free 128-byte row halves per
        // quarter and two independent accumulators per lane); the edge records of a row are fetched eight at a time by
        // the lanes of its quarter and broadcast by shuffle, so the P-row loads of consecutive edges overlap
        {
          const int rw = lane >> 3, q = lane & 7;
          for (int rb = warp * 4; rb < R; rb += 4 * NWARPS) {
            const int r = rb + rw;
            const int code = (r < R) ? sRowCode[r] : 3;
            const bool active = code != 3;  // canonical rows were done on the warp-level tensor path above
            int eb = 0, ee = 0;
            if (active) { eb = sEptr[r]; ee = sEptr[r + 1]; }
            const int deg = ee - eb;
            int dmax = max(deg, __shfl_xor_sync(FULL_MASK, deg, 8));
            dmax = max(dmax, __shfl_xor_sync(FULL_MASK, dmax, 16));  // warp-uniform trip count
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (active) {
              const float* self = sStage + r * LDS_ + 4 * q;  // P_self
              a0 = *reinterpret_cast<const float4*>(self);
              a1 = *reinterpret_cast<const float4*>(self + 32);
              add4(a0, *reinterpret_cast<const float4*>(sBiasC + 4 * q));
              add4(a1, *reinterpret_cast<const float4*>(sBiasC + 32 + 4 * q));
            }
            for (int c0 = 0; c0 < dmax; c0 += 8) {
              const int myb = (c0 + q < deg) ? edge_at(eb + c0 + q) : -1;
              const int cnt = min(8, dmax - c0);
#pragma unroll 4
              for (int k = 0; k < cnt; ++k) {
                const int b = __shfl_sync(FULL_MASK, myb, k, 8);
                // (an edge to the canonical row adds its P row, which is exactly 0: canonical rows of A are zero and
                //  the canonical -> count message arrives through cvec instead)
                if (b >= 0) {
                  const float* src = sP + (b & 127) * LDP + ((b & 0x80) ? 0 : F) + 4 * q;
                  add4(a0, *reinterpret_cast<const float4*>(src));
                  add4(a1, *reinterpret_cast<const float4*>(src + 32));
                }
              }
            }
            if (!active) continue;
            if (code) {
              const float* cv = sCvec + sRowG[r] * 2 * F + (code - 1) * F + 4 * q;
              add4(a0, *reinterpret_cast<const float4*>(cv));
              add4(a1, *reinterpret_cast<const float4*>(cv + 32));
            }
            a0.x = fmaxf(a0.x, 0.f); a0.y = fmaxf(a0.y, 0.f); a0.z = fmaxf(a0.z, 0.f); a0.w = fmaxf(a0.w, 0.f);
            a1.x = fmaxf(a1.x, 0.f); a1.y = fmaxf(a1.y, 0.f); a1.z = fmaxf(a1.z, 0.f); a1.w = fmaxf(a1.w, 0.f);
            float* dst = sStage + r * LDS_ + 4 * q;  // h^{l+1}, fp32 (pooling, canonical inputs)
            *reinterpret_cast<float4*>(dst) = a0;
            *reinterpret_cast<float4*>(dst + 32) = a1;
            uint2 hi, lo;
            split4(a0, hi, lo);
            uint32_t off = tc05::sw128_offset(r, 4 * q);
            *reinterpret_cast<uint2*>(sAhi + off) = hi;  // next layer's A operand
            *reinterpret_cast<uint2*>(sAlo + off) = lo;
            split4(a1, hi, lo);
            off = tc05::sw128_offset(r, 32 + 4 * q);
            *reinterpret_cast<uint2*>(sAhi + off) = hi;
            *reinterpret_cast<uint2*>(sAlo + off) = lo;
          }
        }
        tc05::fence_proxy_async_smem();  // A images written through the generic proxy -> visible to tcgen05.mma
        __syncthreads();
        lap(PH_GATHER);
      }
      lap(PH_POOL);
      cur_tile = s_ticket;
    }
  }
  if (tid == ISSUER && copy_pending) tc05::mbar_wait(&bars[0], wphase);  // never exit with a bulk copy in flight
  if (timed_out) atomicExch(p.status, DESCO_ECUDA);
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int desco_shmp_fused_phase_cycles(uint64_t* out, int32_t reset) {
  if (!out) return DESCO_EINVAL;
  unsigned long long h[PH_COUNT];
  DESCO_CUDA_TRY(cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h)));
  for (int i = 0; i < PH_COUNT; ++i) out[i] = h[i];
  if (reset) {
    for (int i = 0; i < PH_COUNT; ++i) h[i] = 0;
    DESCO_CUDA_TRY(cudaMemcpyToSymbol(g_phase_cycles, h, sizeof(h)));
  }
  return DESCO_OK;
}

int64_t desco_internal_shmp_fused_workspace_bytes(int num_neighborhoods) {
  const int chunks = (num_neighborhoods + CH - 1) / CH;
  return (int64_t)(((size_t)chunks * (CH + 1) * 4 + 255) / 256 * 256 + ((size_t)chunks * 4 + 255) / 256 * 256 + 256);
}

int desco_internal_shmp_fused_layers(const int32_t* nbh_ptr, const int32_t* edge_ptr, const int32_t* edge_col,
                                     const uint8_t* edge_tri, int G, int pyg_batch_size, const float* feat, int input_dim,
                                     const float* w_pre, const void* w_layers_tc, int layers, int passes, float* emb_a,
                                     void* emb_img, float* pool, int emb_ld, void* workspace, int32_t* status,
                                     const int32_t* g_dev, cudaStream_t s) {
  if (G <= 0) return DESCO_OK;
  if (!status || !workspace || !w_layers_tc) return DESCO_EINVAL;
  const int chunks = (G + CH - 1) / CH;
  int32_t* tile_start = (int32_t*)workspace;
  int32_t* tile_count = (int32_t*)((char*)workspace + ((size_t)chunks * (CH + 1) * 4 + 255) / 256 * 256);
  int32_t* ticket = (int32_t*)((char*)tile_count + ((size_t)chunks * 4 + 255) / 256 * 256);
  {
    DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
    shmp_tile_plan_kernel<<<chunks, 1024, 0, s>>>(nbh_ptr, G, g_dev, tile_start, tile_count, ticket, status);
    DESCO_LAUNCH_CHECK();
  }
  static bool attr_set = false;
  if (!attr_set) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(shmp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  FusedArgs a;
  a.nbh_ptr = nbh_ptr; a.edge_ptr = edge_ptr; a.edge_col = edge_col; a.edge_tri = edge_tri;
  a.tile_start = tile_start; a.tile_count = tile_count; a.ticket = ticket; a.g_dev = g_dev;
  a.G = G; a.num_chunks = chunks; a.pyg_batch_size = pyg_batch_size; a.layers = layers; a.passes = passes;
  a.input_dim = input_dim; a.emb_ld = emb_ld;
  a.feat = feat; a.w_pre = w_pre; a.w_layers = (const uint8_t*)w_layers_tc; a.emb_a = emb_a; a.emb_img = (uint8_t*)emb_img; a.pool = pool; a.status = status;
  {
    DescoProfScope prof(DESCO_PROF_SHMP_LAYER, s);
    shmp_fused_kernel<<<desco_num_sms(), THREADS, SMEM_BYTES, s>>>(a);
    DESCO_LAUNCH_CHECK();
  }
  return DESCO_OK;
}
