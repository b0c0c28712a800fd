// Row-wise dense layer on the tensor pipe:  Y = act(X . W^T + b) (+ R),  X fp32 [M, K], W given as pre-swizzled bf16
// hi/mid/lo operand images (desco_b200.tcpack.pack_dense_tc), fp32 accumulation in TMEM.  Both operands are split
// THREE ways (x = hi + mid + lo exactly) and the six products above 2^-24 are accumulated
// (hi.hi, hi.mid, mid.hi, mid.mid, hi.lo, lo.hi): fp32-grade results - these readout GEMMs sum 576 terms with heavy
// cancellation, where the 2-way split of the layer kernel (2^-17) is not enough.  passes = 1 is the plain bf16 variant.
//
// Used for the readout MLPs of the SHMP model (anchor_mlp 576x576, post_mp 576-64-64-256-64: gnn_model.py:40-53) and the
// target / query halves of the count head's first Linear (lightning_model.py:127-131).
//
// One CTA = 128 rows x NBLK output columns.  K is walked in 64-wide atoms through a 2-stage ring: while the tensor
// pipe multiplies atom a, the threads convert atom a+1 of X to bf16 hi/lo (swizzled A images) and the TMA engine
// bulk-copies the weight images of atom a+1.
#include "common.cuh"
#include "tc05.cuh"
#include "shmp_internal.h"
#include "../../include/desco_b200.h"

namespace {

constexpr int TM = 128;
constexpr int THREADS = 256;
constexpr int A_BYTES = TM * 128;  // one bf16 image of a 128 x 64 atom

struct DenseArgs {
  const float* X; int ldx;
  const uint8_t* Ximg;   // [row block][k_atom][hi | mid | lo] A operand images, or NULL (then X is converted here)
  const uint8_t* Wimg;   // [n_block][k_atom][hi | mid | lo], each image nblk * 128 bytes
  const float* bias;     // [N] or NULL
  const float* R; int ldr;  // residual added AFTER the activation, or NULL
  float* Y; int ldy;
  int M, K, nblk, act, passes;
  float slope;
  int32_t* status;
  const int32_t* m_dev;  // device-resident row count (stream-ordered form; M is then the capacity) or NULL
};

__global__ void __launch_bounds__(THREADS, 1) dense_tc_kernel(const DenseArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int b_bytes = p.nblk * 128;
  const int stage_bytes = 3 * A_BYTES + 3 * b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);  // full[2], empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hw = lane >> 4, hl = lane & 15;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * p.nblk;
  const int M = p.m_dev ? min(p.M, *p.m_dev) : p.M;
  if (m0 >= M) return;  // (whole CTA, before any barrier / TMEM allocation)
  const int KA = p.K / 64;
  const uint32_t tmem_cols = p.nblk <= 32 ? 32 : p.nblk <= 64 ? 64 : p.nblk <= 128 ? 128 : 256;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) tc05::mbar_init(&bars[i], 1);
    tc05::fence_mbar_init();
  }
  if (warp == 0) tc05::tmem_alloc(tmem_slot, tmem_cols);
  tc05::fence_before_sync();
  __syncthreads();
  tc05::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = tc05::make_idesc_bf16(TM, p.nblk);
  const uint8_t* wbase = p.Wimg + (size_t)blockIdx.y * KA * 3 * b_bytes;
  bool timed_out = false;

  auto stage_ptr = [&](int s) { return smem + s * stage_bytes; };
  const bool a_img = p.Ximg != nullptr;
  auto issue_b = [&](int a) {  // thread 0: weight images (and, when given, the A images) of atom a -> stage a & 1
    uint8_t* st = stage_ptr(a & 1);
    tc05::mbar_arrive_expect_tx(&bars[a & 1], 3 * b_bytes + (a_img ? 3 * A_BYTES : 0));
    tc05::bulk_g2s(st + 3 * A_BYTES, wbase + (size_t)a * 3 * b_bytes, 3 * b_bytes, &bars[a & 1]);
    if (a_img) {
      const uint8_t* src = p.Ximg + ((size_t)blockIdx.x * KA + a) * (3 * A_BYTES);
      for (int part = 0; part < 3; ++part) tc05::bulk_g2s(st + part * A_BYTES, src + part * A_BYTES, A_BYTES, &bars[a & 1]);
    }
  };
  if (tid == 0) issue_b(0);

  constexpr int ROWS_PER_THREAD = TM / (THREADS / 16);  // 8 row slices (float4) per thread and atom
  float4 xin[ROWS_PER_THREAD], xnext[ROWS_PER_THREAD];
  auto load_x = [&](int a, float4 (&dst)[ROWS_PER_THREAD]) {
#pragma unroll
    for (int it = 0; it < ROWS_PER_THREAD; ++it) {
      const int r = warp * 2 + hw + it * (THREADS / 16);
      dst[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M) dst[it] = __ldg(reinterpret_cast<const float4*>(p.X + (size_t)(m0 + r) * p.ldx + 64 * a + 4 * hl));
    }
  };
  if (!a_img) load_x(0, xin);

  for (int a = 0; a < KA; ++a) {
    const int s = a & 1;
    if (a + 1 < KA) {  // free the other stage (read by the MMAs of atom a-1) and start fetching atom a+1's weights
      if (!a_img) load_x(a + 1, xnext);  // ... and atom a+1 of X: in flight while atom a is converted and multiplied
      if (a >= 1 && !tc05::mbar_wait(&bars[2 + (s ^ 1)], ((a - 1) >> 1) & 1)) timed_out = true;
      if (tid == 0) issue_b(a + 1);
    }
    // X[m0 .. m0+127][64 a .. 64 a + 63] -> bf16 hi / mid / lo swizzled images (one half-warp per row)
    uint8_t* sA = stage_ptr(s);  // hi | mid | lo images of the A atom, then hi | mid | lo of the B atom
    if (!a_img) {
#pragma unroll
    for (int it = 0; it < ROWS_PER_THREAD; ++it) {
      const int r = warp * 2 + hw + it * (THREADS / 16);
      const float4 v = xin[it];
      const float x[4] = {v.x, v.y, v.z, v.w};
      __align__(8) __nv_bfloat16 hi[4], mid[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __float2bfloat16_rn(x[j]);
        const float r1 = x[j] - __bfloat162float(hi[j]);
        mid[j] = __float2bfloat16_rn(r1);
        lo[j] = __float2bfloat16_rn(r1 - __bfloat162float(mid[j]));
      }
      const uint32_t off = tc05::sw128_offset(r, 4 * hl);
      *reinterpret_cast<uint2*>(sA + off) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(sA + A_BYTES + off) = *reinterpret_cast<const uint2*>(mid);
      *reinterpret_cast<uint2*>(sA + 2 * A_BYTES + off) = *reinterpret_cast<const uint2*>(lo);
    }
#pragma unroll
    for (int it = 0; it < ROWS_PER_THREAD; ++it) xin[it] = xnext[it];
    tc05::fence_proxy_async_smem();
    __syncthreads();
    }
    if (tid == 0) {
      if (!tc05::mbar_wait(&bars[s], (a >> 1) & 1)) timed_out = true;
      tc05::fence_after_sync();
      uint64_t dA[3], dB[3];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        dA[i] = tc05::make_smem_desc(sA + i * A_BYTES);
        dB[i] = tc05::make_smem_desc(sA + 3 * A_BYTES + i * b_bytes);
      }
      // products in decreasing magnitude: hi.hi | hi.mid, mid.hi | mid.mid, hi.lo, lo.hi   (a index, b index)
      const int ia[6] = {0, 0, 1, 1, 0, 2}, ib[6] = {0, 1, 0, 1, 2, 0};
      for (int pass = 0; pass < p.passes; ++pass) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc05::mma_bf16(tmem, dA[ia[pass]] + 2 * k, dB[ib[pass]] + 2 * k, idesc, a > 0 || pass > 0 || k > 0);
      }
      tc05::mma_commit(&bars[2 + s]);
    }
  }
  {  // the last commit covers every MMA issued before it
    const int a = KA - 1;
    if (!tc05::mbar_wait(&bars[2 + (a & 1)], (a >> 1) & 1)) timed_out = true;
    tc05::fence_after_sync();
  }
  // epilogue: TMEM -> registers -> bias / activation / residual -> global (thread = row, warp pairs split the columns)
  {
    const int q = warp & 3, half = warp >> 2;
    const int r = m0 + 32 * q + lane;
    const int cols_per_half = (p.nblk + 31) / 32 * 16;  // whole 16-column loads; the tail of the second half is masked
    for (int c = half * cols_per_half; c < (half + 1) * cols_per_half && c < p.nblk; c += 16) {
      float v[16];
      tc05::tmem_ld16(tmem + (static_cast<uint32_t>(32 * q) << 16) + c, v);
      if (r < M) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          if (c + i >= p.nblk) break;
          float o[4] = {v[i], v[i + 1], v[i + 2], v[i + 3]};
          if (p.bias) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + n0 + c + i);
            o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (p.act == 1) o[j] = fmaxf(o[j], 0.f);
            else if (p.act == 2) o[j] = o[j] > 0.f ? o[j] : o[j] * p.slope;
          }
          if (p.R) {
            const float4 rr = *reinterpret_cast<const float4*>(p.R + (size_t)r * p.ldr + n0 + c + i);
            o[0] += rr.x; o[1] += rr.y; o[2] += rr.z; o[3] += rr.w;
          }
          *reinterpret_cast<float4*>(p.Y + (size_t)r * p.ldy + n0 + c + i) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
  }
  if (timed_out) atomicExch(p.status, DESCO_ECUDA);
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem, tmem_cols);
}

}  // namespace

int desco_internal_dense_tc(const float* X, const void* Ximg, int ldx, const void* Wimg, const float* bias, const float* R,
                            int ldr, float* Y, int ldy, int M, int K, int N, int nblk, int act, float slope, int passes,
                            int32_t* status, const int32_t* m_dev, cudaStream_t s) {
  if (M == 0) return DESCO_OK;
  if (K % 64 || nblk % 16 || nblk > 160 || N % nblk || passes < 1 || passes > 6 || !status) return DESCO_EINVAL;
  const size_t smem = 1024 + 2 * (size_t)(3 * A_BYTES + 3 * nblk * 128) + 64;
  static size_t attr = 0;
  if (smem > attr) {
    DESCO_CUDA_TRY(cudaFuncSetAttribute(dense_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  DenseArgs a;
  a.X = X; a.Ximg = (const uint8_t*)Ximg; a.ldx = ldx; a.Wimg = (const uint8_t*)Wimg; a.bias = bias; a.R = R; a.ldr = ldr; a.Y = Y; a.ldy = ldy;
  a.M = M; a.K = K; a.nblk = nblk; a.act = act; a.passes = passes; a.slope = slope; a.status = status; a.m_dev = m_dev;
  dim3 grid((M + TM - 1) / TM, N / nblk);
  DescoProfScope prof(DESCO_PROF_SHMP_OTHER, s);
  dense_tc_kernel<<<grid, THREADS, smem, s>>>(a);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}
