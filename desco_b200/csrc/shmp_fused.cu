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
//      with one bulk async copy per image; the next layer's weights stream in while this layer's epilogue runs.  The
//      issuing lane is held by the tensor pipe for the length of the 12 MMAs, so its warp sits the pool phase out;
//   2. under the MMA, the other 15 warps pool: global_add_pool over the count rows and the canonical update's
//      messages (sum of the rows with a tri / tride edge into the canonical row) as a two-stage fixed-order reduction
//      (stage A: one quarter-warp per run of <= 4 rows; stage B: one thread per neighborhood and 4 features), h_a goes
//      out (skip-concat) and the canonical inputs [sum_tri | sum_tride | h_a] are parked as bf16 hi/lo rows;
//   3. the few canonical rows of the tile run on the warp-level tensor path (mma.sync m16n8k16, A fragments by
//      ldmatrix, weights as host-packed B fragments, same three-pass split):
//        h_a' = relu([sum_tri h_j | sum_tride h_j | h_a] . Wa + b_a),   cvec = h_a . [Cw_tri | Cw_tride];
//   4. TMEM -> shared memory, then the segmented, edge-type-split gather runs out of shared memory, rows taken in
//      descending degree order (counting sort per tile) so that the rows sharing a warp pad to similar lengths:
//        h_i' = relu(P_self[i] + sum_{j in N_tri(i)} P_tri[j] + sum_{j in N_tride(i)} P_tride[j] + cvec[type(i,a)] + b);
//   5. h' is written back as the next layer's bf16 hi/lo A operand (swizzled) and as fp32 rows for pooling.
// Only the pooled sums and the canonical rows ([G, (L+1)*64] each) ever leave the chip.
#include <stdlib.h>
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
constexpr int EDGE_CAP = 4096;          // staged edges per tile (2 bytes each); larger tiles read edges from global
constexpr int ZERO_OFF = TR * LDP;      // float offset (from sP) of an all-zero row: the padding edge of the gather
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
constexpr int SM_STAGE = SM_P + (TR + 1) * LDP * 4;      // + the all-zero row
constexpr int SM_CIN = SM_STAGE + TR * LDS_ * 4;        // bf16 hi then lo, each [MAXC][CINP]: sum_tri h_j | sum_tride h_j | h_a
constexpr int SM_CVEC = SM_CIN + 2 * MAXC * CINP * 2;   // [MAXC][128]
constexpr int SM_CH = SM_CVEC + MAXC * 2 * F * 4;       // [MAXC][64]  h_a of the next layer
constexpr int SM_EDGE = SM_CH + MAXC * F * 4;           // [EDGE_CAP] uint16: float offset of the P row the edge adds
constexpr int SM_EPTR = SM_EDGE + EDGE_CAP * 2;         // [TR + 1] int
constexpr int SM_ROWG = SM_EPTR + ((TR + 1) * 4 + 15) / 16 * 16;  // [TR] uint8 local neighborhood of the row
constexpr int SM_ROWCODE = SM_ROWG + TR;                // [TR] uint8: 0 none, 1 tri / 2 tride edge to canonical, 3 canonical
constexpr int SM_NBHLO = SM_ROWCODE + TR;               // [MAXC + 1] int local first row
constexpr int SM_QUIRK = SM_NBHLO + ((MAXC + 1) * 4 + 15) / 16 * 16;  // [MAXC] int local row or -1
constexpr int SM_BIASC = SM_QUIRK + (MAXC * 4 + 15) / 16 * 16;        // [layers][64] float: bias of the count rows
constexpr int MAX_LAYERS_SMEM = 8;      // count-row biases of up to this many layers stay resident in shared memory
constexpr int SM_ITEM = SM_BIASC + MAX_LAYERS_SMEM * F * 4;                             // [64] int: pooling work items (first row | rows << 8)
constexpr int SM_ITEMBASE = SM_ITEM + 64 * 4;                         // [32] int: first item of neighborhood i
constexpr int SM_HIST = SM_ITEMBASE + 32 * 4;                         // [66] int: rows per degree key, then their exclusive scan
constexpr int SM_ORDER = SM_HIST + 68 * 4;                            // [TR] uint8: rows by descending degree (gather order)
constexpr int SM_CANIN = SM_ORDER + TR;                               // [TR] uint8: 1 tri / 2 tride edge INTO the canonical row
constexpr int SM_BARS = SM_CANIN + TR;                                // 2 mbarriers + tmem slot
constexpr int SM_TOTAL = SM_BARS + 64;
constexpr int SMEM_BYTES = SM_TOTAL + 1024;  // slack for the manual 1024-B alignment
static_assert(MAXC <= 32 && MAXC * CINP * 2 + MAXC * 3 * F * 4 >= 32 * CINP * 2, "the second mma row block reads (and ignores) rows past MAXC");
static_assert(SMEM_BYTES <= 232448, "fused SHMP kernel exceeds the 227 KB shared-memory limit");
static_assert(SM_AHI % 1024 == 0 && SM_ALO % 1024 == 0 && SM_BLO % 1024 == 0, "UMMA tiles must be 1024-B aligned");
static_assert(SM_CIN % 16 == 0 && (CINP * 2) % 16 == 0 && SM_BARS % 8 == 0, "ldmatrix rows / mbarriers");
static_assert(ZERO_OFF + 2 * F <= 65535, "edge records are 16-bit float offsets");
// pooling work items (<= 4 count rows of one neighborhood each): sum ceil(n_i / 4) <= (TR - nc) / 4 + 3 nc / 4
static_assert(MAXC < 32 && TR / 4 + (3 * MAXC + 3) / 4 <= (NWARPS - 1) * 4 && 16 * MAXC + 8 * MAXC <= THREADS - 32,
              "pool phase thread / item budget (15 warps: the issuing warp does not take part)");
static_assert(64 * 3 * F * 4 <= TR * LDP * 4, "the pooling partials alias the (dead) P buffer");

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
enum { PH_SETUP = 0, PH_POOL, PH_ISSUE, PH_CANON, PH_WAIT_MMA, PH_T2S, PH_GATHER, PH_POOLB, PH_ISSUER_WAIT, PH_ISSUER_MMA, PH_COUNT };
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
// one 16 x 16 bf16 A fragment (four 8 x 8 matrices: rows 0-7 / 8-15 x k 0-7 / 8-15) in one instruction; lane i passes the
// shared address of row (i & 7) of matrix i >> 3
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// a += b as two packed fp32x2 adds (sm_100 FADD2: half the issue slots of four scalar adds; same rounding)
__device__ __forceinline__ void add4(float4& a, const float4 b) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%6, %7};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%2, %3}, ra;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w));
}
// a += b * m as two packed fp32x2 FMAs
__device__ __forceinline__ void fma4(float4& a, const float4 b, const float m) {
  asm("{\n\t.reg .b64 ra, rb, rm;\n\t"
      "mov.b64 rm, {%8, %8};\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%4, %5};\n\tfma.rn.f32x2 ra, rb, rm, ra;\n\tmov.b64 {%0, %1}, ra;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%6, %7};\n\tfma.rn.f32x2 ra, rb, rm, ra;\n\tmov.b64 {%2, %3}, ra;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w), "f"(m));
}
__device__ __forceinline__ uint32_t bf2_bits(const __nv_bfloat162 v) { return *reinterpret_cast<const uint32_t*>(&v); }
// four fp32 -> bf16 hi / lo pairs with the packed converts (x = hi + lo + O(2^-17 |x|), as tc05::split_bf16)
__device__ __forceinline__ void split4(const float4 v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - f01.x, v.y - f01.y), l23 = __floats2bfloat162_rn(v.z - f23.x, v.w - f23.y);
  hi = make_uint2(bf2_bits(h01), bf2_bits(h23));
  lo = make_uint2(bf2_bits(l01), bf2_bits(l23));
}
// two fp32 -> bf16 hi / mid / lo pairs: hi + mid + lo == x exactly (3 x 8 significand bits)
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t (&out)[3]) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const float2 hf = __bfloat1622float2(h);
  const float r0 = x0 - hf.x, r1 = x1 - hf.y;
  const __nv_bfloat162 m = __floats2bfloat162_rn(r0, r1);
  const float2 mf = __bfloat1622float2(m);
  const __nv_bfloat162 l = __floats2bfloat162_rn(r0 - mf.x, r1 - mf.y);
  out[0] = bf2_bits(h); out[1] = bf2_bits(m); out[2] = bf2_bits(l);
}
// four fp32 -> the hi | mid | lo operand images of csrc/dense_tc.cu (each `plane` bytes apart)
__device__ __forceinline__ void store_split3(uint8_t* dst, uint32_t plane, const float4 v) {
  uint32_t a[3], b[3];
  split3_pair(v.x, v.y, a);
  split3_pair(v.z, v.w, b);
#pragma unroll
  for (int i = 0; i < 3; ++i) *reinterpret_cast<uint2*>(dst + i * plane) = make_uint2(a[i], b[i]);
}
template <bool TIMED>
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
  uint16_t* sEdge = reinterpret_cast<uint16_t*>(smem + SM_EDGE);
  int* sEptr = reinterpret_cast<int*>(smem + SM_EPTR);
  uint8_t* sRowG = smem + SM_ROWG;
  uint8_t* sRowCode = smem + SM_ROWCODE;
  int* sNbhLo = reinterpret_cast<int*>(smem + SM_NBHLO);
  int* sQuirk = reinterpret_cast<int*>(smem + SM_QUIRK);
  float* sBiasC = reinterpret_cast<float*>(smem + SM_BIASC);
  int* sItem = reinterpret_cast<int*>(smem + SM_ITEM);
  int* sItemBase = reinterpret_cast<int*>(smem + SM_ITEMBASE);
  int* sHist = reinterpret_cast<int*>(smem + SM_HIST);
  uint8_t* sOrder = smem + SM_ORDER;
  uint8_t* sCanIn = smem + SM_CANIN;
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
  const bool bias_resident = p.layers <= MAX_LAYERS_SMEM;  // else: refetched per layer into slot 0
  if (bias_resident)
    for (int i = tid; i < p.layers * F; i += THREADS)
      sBiasC[i] = __ldg(reinterpret_cast<const float*>(p.w_layers + (size_t)(i / F) * LAYER_BYTES + OFF_BIASC) + (i % F));
  bool copy_pending = true;  // meaningful in the ISSUER thread only
  uint32_t wphase = 0, mphase = 0;
  bool timed_out = false;
  const uint32_t idesc = tc05::make_idesc_bf16(TR, NB);
  const int hw = lane >> 4, hl = lane & 15;  // half-warp id / lane inside the half-warp (layer-0 inputs)
  const int rw = lane >> 3, q = lane & 7;    // quarter-warp id / lane inside the quarter (gathers: one quarter per row)
  const int fg = lane >> 2, ft = lane & 3;   // mma fragment coordinates: group (row / column), thread in group

  // this warp's B fragments of the canonical-row weights of a layer (warps 0-7: z_a column tile `warp`, K = 192, 12
  // k-steps; warps 8-15: cvec column tiles 2(warp-8), 2(warp-8)+1, K = 64, 4 k-steps each) and its bias: requested at the
  // start of the layer, consumed after the pool phase.  (80 KB per tile and layer through the load/store path; when
  // they are requested makes no difference to the kernel time - measured at the layer start, in instalments over the
  // pool phase, and one layer ahead spread over the TMEM read-out and the gather: 0.240 ms each - because the L1
  // wavefront pipe, shared with all the shared-memory traffic, is the busiest unit either way.)
  uint4 wf[12];
  float2 bias_a = make_float2(0.f, 0.f);
  auto fetch_frags = [&](int layer) {
    const uint8_t* wl = p.w_layers + (size_t)layer * LAYER_BYTES;
    if (warp < 8) {
      bias_a = __ldg(reinterpret_cast<const float2*>(reinterpret_cast<const float*>(wl + OFF_BIASA) + 8 * warp + 2 * ft));
      const uint4* W = reinterpret_cast<const uint4*>(wl + OFF_WAT) + (size_t)warp * 12 * 32 + lane;
#pragma unroll
      for (int ks = 0; ks < 12; ++ks) wf[ks] = __ldg(W + 32 * ks);
    } else {
      const uint4* W = reinterpret_cast<const uint4*>(wl + OFF_CWT) + (size_t)(2 * (warp - 8)) * 4 * 32 + lane;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) wf[ks] = __ldg(W + 32 * ks);
    }
  };

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
      long long tick = TIMED ? clock64() : 0;
      auto lap = [&](int phase) {  // phase timing: compiled out of the production instantiation
        if (TIMED && tid == 0) {
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
      // an edge is staged as the float offset (from sP) of the P row it adds: P_tri[j] or P_tride[j]
      auto edge_off_global = [&](int e) -> int { return (p.edge_col[e0 + e] - row0) * LDP + (p.edge_tri[e0 + e] ? 0 : F); };
      if (edges_staged)
        for (int e = tid; e < Et; e += THREADS) sEdge[e] = (uint16_t)edge_off_global(e);
      for (int i = tid; i < 2 * TR * 128 / 16; i += THREADS)  // zero both A images (rows >= R and canonical rows stay 0)
        reinterpret_cast<uint4*>(sAhi)[i] = make_uint4(0u, 0u, 0u, 0u);
      // rows >= R of the fp32 tile are read (times 0) by the pooling GEMM: keep them finite; and the all-zero P row
      for (int i = R * LDS_ + tid; i < TR * LDS_; i += THREADS) sStage[i] = 0.f;
      if (tid < LDP) sP[ZERO_OFF + tid] = 0.f;
      __syncthreads();
      auto edge_off = [&](int e) -> int { return edges_staged ? (int)sEdge[e] : edge_off_global(e); };
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
            const int last = edge_off(ee - 1);  // canonical = max row of its neighborhood = last entry of a sorted row
            const int j = last / LDP;
            if (j == canon) code = (last - j * LDP < F) ? 1 : 2;
          }
        }
        sRowG[tid] = (uint8_t)a;
        sRowCode[tid] = (uint8_t)code;
      }
      if (warp == 4) {  // pooling work items: runs of <= 4 count rows of one neighborhood, in (neighborhood, row) order
        int n = 0, lo = 0;
        if (lane < nc) { lo = sNbhLo[lane]; n = sNbhLo[lane + 1] - 1 - lo; }
        const int it = (n + 3) >> 2, base = warp_incl_scan(it) - it;
        if (lane <= nc) sItemBase[lane] = base;  // lane == nc: the total
        for (int k = 0; k < it; ++k) sItem[base + k] = (lo + 4 * k) | (min(4, n - 4 * k) << 8);
      } else if (warp >= 5 && warp < 9) {
        sCanIn[tid - 160] = 0;
      } else if (warp >= 9 && warp < 12) {
        if (tid - 288 < 66) sHist[tid - 288] = 0;
      }
      __syncthreads();

      // edges INTO the canonical rows (the messages sum_tri h_j / sum_tride h_j of the canonical update), as a per-row code
      for (int i = warp; i < nc; i += NWARPS) {
        const int canon = sNbhLo[i + 1] - 1, quirk = sQuirk[i];
        for (int e = sEptr[canon] + lane, ee = sEptr[canon + 1]; e < ee; e += 32) {
          const int o = edge_off(e), j = o / LDP;
          if (j != quirk) sCanIn[j] = (o - j * LDP < F) ? 1 : 2;
        }
      }
      // gather order: rows by descending degree (counting sort), so that the rows sharing a warp pad to similar lengths
      int sort_key = 64, sort_pos = 0;  // key 64: not gathered (canonical rows, rows >= R)
      if (tid < TR) {
        if (tid < R && sRowCode[tid] != 3) sort_key = 63 - min(sEptr[tid + 1] - sEptr[tid], 63);
        sort_pos = atomicAdd(&sHist[sort_key], 1);
      }
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
            *reinterpret_cast<float4*>(sStage + r * LDS_ + 4 * hl) = make_float4(0.f, 0.f, 0.f, 0.f);  // masked out of the pool
          } else {
            *reinterpret_cast<float4*>(sStage + r * LDS_ + 4 * hl) = v;
            uint2 hi, lo;
            split4(v, hi, lo);
            const uint32_t off = tc05::sw128_offset(r, 4 * hl);
            *reinterpret_cast<uint2*>(sAhi + off) = hi;
            *reinterpret_cast<uint2*>(sAlo + off) = lo;
          }
        }
      }
      tc05::fence_proxy_async_smem();  // A images written through the generic proxy -> visible to tcgen05.mma
      __syncthreads();
      if (warp == 1) {  // exclusive scan of the degree histogram (65 keys); consumed after the first pool barrier
        const int a = sHist[2 * lane], b = sHist[2 * lane + 1];
        const int incl = warp_incl_scan(a + b);
        __syncwarp();
        sHist[2 * lane] = incl - a - b;
        sHist[2 * lane + 1] = incl - b;
        if (lane == 31) sHist[64] = incl;  // = number of gathered rows
      }
      lap(PH_SETUP);

      for (int l = 0; l <= p.layers; ++l) {
        // ------------ tensor pipe: P = h . [W_tri | W_tride | W_self], issued first so that it runs under the pool
        // and canonical phases ------------
        if (l < p.layers && tid == ISSUER) {
          const long long t_w0 = TIMED ? clock64() : 0;
          if (!tc05::mbar_wait(&bars[0], wphase)) timed_out = true;
          const long long t_w1 = TIMED ? clock64() : 0;
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
          if (TIMED) {
            atomicAdd(&g_phase_cycles[PH_ISSUER_WAIT], (unsigned long long)(t_w1 - t_w0));
            atomicAdd(&g_phase_cycles[PH_ISSUER_MMA], (unsigned long long)(clock64() - t_w1));
          }
        }
        if (l < p.layers) wphase ^= 1;
        if (l < p.layers) fetch_frags(l);
        if (!bias_resident && l < p.layers && tid < F)  // read after two barriers
          sBiasC[tid] = __ldg(reinterpret_cast<const float*>(p.w_layers + (size_t)l * LAYER_BYTES + OFF_BIASC) + tid);
        lap(PH_ISSUE);
        // ------------ pool + canonical inputs of layer l (from the fp32 rows of h^l in sStage, h_a^l in sCh) ------------
        // stage A: one quarter-warp per work item (<= 4 rows): partial sums of the rows (global_add_pool, gnn_model.py:107)
        // and of the rows with a tri / tride edge into the canonical row (the canonical update's messages); fixed order
        // (the last warp sits the pool phase out: its lane 0 is held by the tensor pipe for the length of the 12 MMAs it
        //  issues - tcgen05.mma issue blocks until the pipe accepts the instruction - and would make every stage-A
        //  barrier wait ~1.2 k cycles; stages A and B fit the other 15 warps and meet at a named barrier)
        float* sPart = sP;  // [item][pool 64 | tri 64 | tride 64]; the P buffer is dead between the gather and the next T2S
        if (warp < NWARPS - 1) {
          const int item = warp * 4 + rw;
          if (item < sItemBase[nc]) {
            const int it = sItem[item], r0 = it & 255, cnt = it >> 8;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 p0 = z4, p1 = z4, t0 = z4, t1 = z4, d0 = z4, d1 = z4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (j < cnt) {
                const float* src = sStage + (r0 + j) * LDS_ + 4 * q;
                const float4 v0 = *reinterpret_cast<const float4*>(src), v1 = *reinterpret_cast<const float4*>(src + 32);
                const int cin = sCanIn[r0 + j];
                add4(p0, v0); add4(p1, v1);
                // branch-free (the quarters of a warp disagree on cin, and a divergent branch per row cost ~200 cycles)
                const float mt = cin == 1 ? 1.f : 0.f, md = cin == 2 ? 1.f : 0.f;
                fma4(t0, v0, mt); fma4(t1, v1, mt);
                fma4(d0, v0, md); fma4(d1, v1, md);
              }
            }
            float* dst = sPart + item * (3 * F) + 4 * q;
            *reinterpret_cast<float4*>(dst) = p0; *reinterpret_cast<float4*>(dst + 32) = p1;
            *reinterpret_cast<float4*>(dst + F) = t0; *reinterpret_cast<float4*>(dst + F + 32) = t1;
            *reinterpret_cast<float4*>(dst + 2 * F) = d0; *reinterpret_cast<float4*>(dst + 2 * F + 32) = d1;
          }
        }
        if (warp < NWARPS - 1) asm volatile("bar.sync 1, %0;" ::"n"(THREADS - 32) : "memory");
        lap(PH_POOL);
        if (l == 0 && tid < TR) sOrder[sHist[sort_key] + sort_pos] = (uint8_t)tid;  // gather order (read two barriers on)
        // stage B: (neighborhood, 4 features) sums its items in order; h_a^l goes out (skip-concat, gnn_model.py:275) and
        // joins the canonical inputs [sum_tri h_j | sum_tride h_j | h_a] as bf16 hi/lo rows
        {
          auto put = [&](int i, int colbase, const float4 v) {
            uint2 hi, lo;
            split4(v, hi, lo);
            *reinterpret_cast<uint2*>(sCinHi + i * CINP + colbase) = hi;
            *reinterpret_cast<uint2*>(sCinLo + i * CINP + colbase) = lo;
          };
          if (tid < 16 * MAXC) {
            const int i = tid >> 4, f = (tid & 15) * 4;
            if (i < nc) {
              const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
              float4 ps = z4, tsum = z4, dsum = z4;
              for (int k = sItemBase[i], k1 = sItemBase[i + 1]; k < k1; ++k) {
                const float* src = sPart + k * (3 * F) + f;
                add4(ps, *reinterpret_cast<const float4*>(src));
                add4(tsum, *reinterpret_cast<const float4*>(src + F));
                add4(dsum, *reinterpret_cast<const float4*>(src + 2 * F));
              }
              *reinterpret_cast<float4*>(p.pool + (size_t)(nb0 + i) * p.emb_ld + (size_t)l * F + f) = ps;
              if (l < p.layers) { put(i, f, tsum); put(i, F + f, dsum); }
            }
          } else if (tid < 16 * MAXC + 8 * MAXC) {
            const int i = (tid - 16 * MAXC) >> 3;  // 16 MAXC is a multiple of 32: lane & 7 == q
            if (i < nc) {
              const float4 ha0 = *reinterpret_cast<const float4*>(sCh + i * F + 4 * q);
              const float4 ha1 = *reinterpret_cast<const float4*>(sCh + i * F + 32 + 4 * q);
              const int g = nb0 + i;
              if (p.emb_img) {  // split three ways for csrc/dense_tc.cu
                uint8_t* img = p.emb_img + ((size_t)(g >> 7) * (p.layers + 1) + l) * (3 * TR * 128);
                store_split3(img + tc05::sw128_offset(g & 127, 4 * q), TR * 128, ha0);
                store_split3(img + tc05::sw128_offset(g & 127, 32 + 4 * q), TR * 128, ha1);
              } else {
                float* dst = p.emb_a + (size_t)g * p.emb_ld + (size_t)l * F + 4 * q;
                *reinterpret_cast<float4*>(dst) = ha0;
                *reinterpret_cast<float4*>(dst + 32) = ha1;
              }
              if (l < p.layers) { put(i, 2 * F + 4 * q, ha0); put(i, 2 * F + 32 + 4 * q, ha1); }
            }
          }
        }
        if (l == p.layers) break;
        __syncthreads();
        lap(PH_POOLB);

        // ------------ canonical rows on the warp-level tensor path (mma.sync m16n8k16, bf16 hi/lo split, 3 passes:
        // same fp32-grade products as the tile GEMM); 16 canonical rows per row block, A fragments by ldmatrix ------------
        for (int mb = 0; mb < nc; mb += 16) {
          const int lrow = mb + (lane & 7) + ((lane >> 3) & 1) * 8, lk = (lane >> 4) * 8;
          const uint32_t ahi_addr = tc05::smem_u32(sCinHi + lrow * CINP + lk), alo_addr = tc05::smem_u32(sCinLo + lrow * CINP + lk);
          if (warp < 8) {  // z_a = [sum_tri | sum_tride | h_a] . Wa, columns 8 warp .. 8 warp + 7
            float acc[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int ks = 0; ks < 12; ++ks) {
              uint32_t ahi[4], alo[4];
              ldsm4(ahi, ahi_addr + 32 * ks);
              ldsm4(alo, alo_addr + 32 * ks);
              mma16816(acc[0], ahi, wf[ks].x, wf[ks].y);
              mma16816(acc[1], alo, wf[ks].x, wf[ks].y);
              mma16816(acc[2], ahi, wf[ks].z, wf[ks].w);
            }
            const int n = 8 * warp + 2 * ft;
            const float2 b = bias_a;
            const int r0 = mb + fg, r1 = r0 + 8;  // h_a^{l+1}; read again only after the next barriers
            if (r0 < nc)
              *reinterpret_cast<float2*>(sCh + r0 * F + n) =
                  make_float2(fmaxf((acc[0][0] + acc[1][0]) + acc[2][0] + b.x, 0.f), fmaxf((acc[0][1] + acc[1][1]) + acc[2][1] + b.y, 0.f));
            if (r1 < nc)
              *reinterpret_cast<float2*>(sCh + r1 * F + n) =
                  make_float2(fmaxf((acc[0][2] + acc[1][2]) + acc[2][2] + b.x, 0.f), fmaxf((acc[0][3] + acc[1][3]) + acc[2][3] + b.y, 0.f));
          } else {  // cvec = h_a . [Cw_tri | Cw_tride], column tiles 2(warp-8), 2(warp-8)+1
            float acc[2][3][4];
#pragma unroll
            for (int i = 0; i < 24; ++i) (&acc[0][0][0])[i] = 0.f;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              uint32_t ahi[4], alo[4];  // the h_a block starts at bf16 column 2F
              ldsm4(ahi, ahi_addr + 2 * (2 * F) + 32 * ks);
              ldsm4(alo, alo_addr + 2 * (2 * F) + 32 * ks);
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                mma16816(acc[j][0], ahi, wf[4 * j + ks].x, wf[4 * j + ks].y);
                mma16816(acc[j][1], alo, wf[4 * j + ks].x, wf[4 * j + ks].y);
                mma16816(acc[j][2], ahi, wf[4 * j + ks].z, wf[4 * j + ks].w);
              }
            }
            const int r0 = mb + fg, r1 = r0 + 8;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const int n = 8 * (2 * (warp - 8) + j) + 2 * ft;
              if (r0 < nc)
                *reinterpret_cast<float2*>(sCvec + r0 * 2 * F + n) =
                    make_float2((acc[j][0][0] + acc[j][1][0]) + acc[j][2][0], (acc[j][0][1] + acc[j][1][1]) + acc[j][2][1]);
              if (r1 < nc)
                *reinterpret_cast<float2*>(sCvec + r1 * 2 * F + n) =
                    make_float2((acc[j][0][2] + acc[j][1][2]) + acc[j][2][2], (acc[j][0][3] + acc[j][1][3]) + acc[j][2][3]);
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
          const int tq = warp & 3, cg = warp >> 2;  // TMEM lane quarter of this warp, column group of 48
          const int r = 32 * tq + lane;
          uint32_t v[3][16];  // the three loads of this warp's 48 columns are in flight together
#pragma unroll
          for (int j = 0; j < 3; ++j) tc05::tmem_ld16_async(tmem + (static_cast<uint32_t>(32 * tq) << 16) + cg * 48 + j * 16, v[j]);
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

        // ------------ segmented, edge-type-split gather out of shared memory; one quarter-warp per row, two rows
        // interleaved per quarter so that two independent chains of P-row loads are in flight; rows are taken in
        // descending degree order, so the eight rows of a warp pad to similar lengths ------------
        // lane q of a quarter owns features 4q..4q+3 and 32+4q..32+4q+3 (two conflict-free 128-byte row halves per
        // quarter); the edge records of a row (ready-made P-row offsets) are fetched eight at a time by the lanes of its
        // quarter and broadcast by shuffle; a missing edge adds the all-zero row, so the inner loop has no branch
        {
          const float* sPq = sP + 4 * q;
          const float* biasc = sBiasC + (bias_resident ? l * F : 0) + 4 * q;
          const int nact = sHist[64];
          int rr[2], dg[2], eb[2], code[2];
          float4 a0[2], a1[2];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int slot = warp * 8 + rw + 4 * h;  // canonical rows (done on the warp-level tensor path above) sort last
            rr[h] = (slot < nact) ? sOrder[slot] : 0;
            code[h] = (slot < nact) ? sRowCode[rr[h]] : 3;
            eb[h] = 0; dg[h] = 0;
            a0[h] = make_float4(0.f, 0.f, 0.f, 0.f); a1[h] = a0[h];
            if (code[h] != 3) {
              eb[h] = sEptr[rr[h]];
              dg[h] = sEptr[rr[h] + 1] - eb[h];
              const float* self = sStage + rr[h] * LDS_ + 4 * q;  // P_self
              a0[h] = *reinterpret_cast<const float4*>(self);
              a1[h] = *reinterpret_cast<const float4*>(self + 32);
              add4(a0[h], *reinterpret_cast<const float4*>(biasc));
              add4(a1[h], *reinterpret_cast<const float4*>(biasc + 32));
            }
          }
          int dmax = max(dg[0], dg[1]);
          dmax = max(dmax, __shfl_xor_sync(FULL_MASK, dmax, 8));
          dmax = max(dmax, __shfl_xor_sync(FULL_MASK, dmax, 16));  // warp-uniform trip count
          for (int c0 = 0; c0 < dmax; c0 += 8) {
            const int my0 = (c0 + q < dg[0]) ? edge_off(eb[0] + c0 + q) : ZERO_OFF;
            const int my1 = (c0 + q < dg[1]) ? edge_off(eb[1] + c0 + q) : ZERO_OFF;
            const int cnt = min(8, dmax - c0);
#pragma unroll 4
            for (int k = 0; k < cnt; ++k) {
              // (an edge to the canonical row adds its P row, which is exactly 0: canonical rows of A are zero and
              //  the canonical -> count message arrives through cvec instead)
              const float* s0 = sPq + __shfl_sync(FULL_MASK, my0, k, 8);
              const float* s1 = sPq + __shfl_sync(FULL_MASK, my1, k, 8);
              add4(a0[0], *reinterpret_cast<const float4*>(s0));
              add4(a1[0], *reinterpret_cast<const float4*>(s0 + 32));
              add4(a0[1], *reinterpret_cast<const float4*>(s1));
              add4(a1[1], *reinterpret_cast<const float4*>(s1 + 32));
            }
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (code[h] == 3) continue;
            const int r = rr[h];
            if (code[h]) {
              const float* cv = sCvec + sRowG[r] * 2 * F + (code[h] - 1) * F + 4 * q;
              add4(a0[h], *reinterpret_cast<const float4*>(cv));
              add4(a1[h], *reinterpret_cast<const float4*>(cv + 32));
            }
            float4 x0 = a0[h], x1 = a1[h];
            x0.x = fmaxf(x0.x, 0.f); x0.y = fmaxf(x0.y, 0.f); x0.z = fmaxf(x0.z, 0.f); x0.w = fmaxf(x0.w, 0.f);
            x1.x = fmaxf(x1.x, 0.f); x1.y = fmaxf(x1.y, 0.f); x1.z = fmaxf(x1.z, 0.f); x1.w = fmaxf(x1.w, 0.f);
            float* dst = sStage + r * LDS_ + 4 * q;  // h^{l+1}, fp32 (pooling, canonical inputs)
            *reinterpret_cast<float4*>(dst) = x0;
            *reinterpret_cast<float4*>(dst + 32) = x1;
            uint2 hi, lo;
            split4(x0, hi, lo);
            uint32_t off = tc05::sw128_offset(r, 4 * q);
            *reinterpret_cast<uint2*>(sAhi + off) = hi;  // next layer's A operand
            *reinterpret_cast<uint2*>(sAlo + off) = lo;
            split4(x1, hi, lo);
            off = tc05::sw128_offset(r, 32 + 4 * q);
            *reinterpret_cast<uint2*>(sAhi + off) = hi;
            *reinterpret_cast<uint2*>(sAlo + off) = lo;
          }
        }
        tc05::fence_proxy_async_smem();  // A images written through the generic proxy -> visible to tcgen05.mma
        __syncthreads();
        lap(PH_GATHER);
      }
      lap(PH_POOLB);
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
  // DESCO_FUSED_PHASE_TIMING=1 launches the instantiation with per-phase clock64 counters (profiles/tools)
  static const bool timed = [] { const char* e = getenv("DESCO_FUSED_PHASE_TIMING"); return e && e[0] == '1'; }();
  static bool attr_set = false;
  if (!attr_set) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(shmp_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    DESCO_CUDA_TRY(cudaFuncSetAttribute(shmp_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
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
    if (timed) shmp_fused_kernel<true><<<desco_num_sms(), THREADS, SMEM_BYTES, s>>>(a);
    else shmp_fused_kernel<false><<<desco_num_sms(), THREADS, SMEM_BYTES, s>>>(a);
    DESCO_LAUNCH_CHECK();
  }
  return DESCO_OK;
}
