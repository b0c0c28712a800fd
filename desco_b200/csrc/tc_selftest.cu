// Unit-test kernel for the tcgen05 plumbing of tc05.cuh (no reference counterpart; the SHMP kernels build on it):
//   D[128][N] = A[128][64] . B[N][64]^T  with the bf16 hi/lo operand split, fp32 accumulation in TMEM.
// A arrives as plain fp32 rows and is split + swizzled by the threads (what the SHMP gather epilogue does);
// B arrives as the pre-swizzled hi/lo image built on the host (desco_b200.tcpack.pack_b_operand) and is fetched with
// one bulk async copy per image (what the SHMP layer loop does with its weights).
#include "common.cuh"
#include "tc05.cuh"
#include "../../include/desco_b200.h"

namespace {

constexpr int M = 128;
constexpr int K = 64;

__global__ void __launch_bounds__(128) tc_selftest_kernel(const float* __restrict__ A, const uint8_t* __restrict__ b_image,
                                                          int N, int passes, float* __restrict__ D, int* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-B alignment computed as an OFFSET from the __shared__ symbol, so the compiler keeps the shared address space
  // (LDS/STS); casting through uintptr_t would turn every access into a generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (tc05::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sAhi = smem;
  uint8_t* sAlo = sAhi + M * 128;
  uint8_t* sBhi = sAlo + M * 128;
  uint8_t* sBlo = sBhi + N * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBlo + N * 128);  // [0] weights landed, [1] mma done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
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

  if (tid == 0) {
    const uint32_t bytes = static_cast<uint32_t>(N) * 128u;
    tc05::mbar_arrive_expect_tx(&bars[0], 2 * bytes);
    tc05::bulk_g2s(sBhi, b_image, bytes, &bars[0]);
    tc05::bulk_g2s(sBlo, b_image + bytes, bytes, &bars[0]);
  }
  {  // row `tid` of A: split into bf16 hi / lo, store 16-byte chunks at their swizzled positions
    const float* a = A + static_cast<size_t>(tid) * K;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) tc05::split_bf16(a[c * 8 + j], hi[j], lo[j]);
      const uint32_t off = tc05::sw128_offset(tid, c * 8);
      *reinterpret_cast<uint4*>(sAhi + off) = *reinterpret_cast<const uint4*>(hi);
      *reinterpret_cast<uint4*>(sAlo + off) = *reinterpret_cast<const uint4*>(lo);
    }
  }
  tc05::fence_proxy_async_smem();
  __syncthreads();

  bool ok = true;
  if (tid == 0) {
    ok = tc05::mbar_wait(&bars[0], 0);
    tc05::fence_after_sync();
    const uint32_t idesc = tc05::make_idesc_bf16(M, N);
    const uint64_t dAhi = tc05::make_smem_desc(sAhi), dAlo = tc05::make_smem_desc(sAlo);
    const uint64_t dBhi = tc05::make_smem_desc(sBhi), dBlo = tc05::make_smem_desc(sBlo);
    bool acc = false;
    for (int p = 0; p < passes; ++p) {
      const uint64_t da = (p == 1) ? dAlo : dAhi;
      const uint64_t db = (p == 2) ? dBlo : dBhi;
#pragma unroll
      for (int k = 0; k < K / 16; ++k) {  // UMMA_K = 16 bf16 = 32 bytes = 2 descriptor units
        tc05::mma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, acc);
        acc = true;
      }
    }
    tc05::mma_commit(&bars[1]);
  }
  __syncwarp();
  ok = tc05::mbar_wait(&bars[1], 0) && ok;
  tc05::fence_after_sync();
  if (!ok) {
    if ((tid & 31) == 0) atomicExch(status, DESCO_ECUDA);
  } else {
    for (int c0 = 0; c0 < N; c0 += 32) {
      float v[32];
      tc05::tmem_ld32(tmem + (static_cast<uint32_t>(32 * warp) << 16) + c0, v);
      float* d = D + static_cast<size_t>(tid) * N + c0;
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = v[j];
    }
  }
  tc05::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc05::tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int desco_tc_selftest(const float* a, const void* b_image, int32_t n, int32_t passes, float* d, int32_t* status,
                                 void* stream) {
  if (!a || !b_image || !d || !status) return DESCO_EINVAL;
  if (n < 16 || n > 256 || n % 32 != 0 || passes < 1 || passes > 3) return DESCO_EINVAL;
  const size_t smem = 1024 + 2 * M * 128 + 2 * static_cast<size_t>(n) * 128 + 64;
  DESCO_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  desco_count_launches(1);
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(a, static_cast<const uint8_t*>(b_image), n, passes, d, status);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}
