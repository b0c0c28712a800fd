// Shared helpers for the desco_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DESCO_OK 0
#define DESCO_EINVAL (-22)
#define DESCO_ENOMEM (-12)
#define DESCO_ECUDA (-5)
#define DESCO_ERANGE (-34)
#define DESCO_ENOBUFS (-105)

#define DESCO_CUDA_TRY(expr)                      \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) return DESCO_ECUDA;    \
  } while (0)

#define DESCO_LAUNCH_CHECK()                      \
  do {                                            \
    cudaError_t _e = cudaGetLastError();          \
    if (_e != cudaSuccess) return DESCO_ECUDA;    \
  } while (0)

// launch accounting + optional per-kernel-group device timing (runtime.cu); slots: include/desco_b200.h DESCO_PROF_*
void desco_count_launches(int n);
class DescoProfScope {
 public:
  DescoProfScope(int slot, cudaStream_t s, int launches = 1);
  ~DescoProfScope();

 private:
  int slot_, launches_;
  cudaStream_t s_;
  bool on_;
  cudaEvent_t a_, b_;
};

constexpr unsigned FULL_MASK = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_incl_scan(int v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(FULL_MASK, v, d);
    if (lane_id() >= d) v += t;
  }
  return v;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
  return v;
}

// FFMA tile GEMM shared by the SHMP / gossip / readout kernels: 256 threads, thread (ty,tx) owns rows ty*4..+3 and
// columns tx*4..+3 of a 64 x 64 output tile.
// acc[4][4] += sA[ty*4+i][0..K) . sW[0..K)[tx*4+j]
template <int K, int LD, int LDW = 64>
__device__ __forceinline__ void tile_gemm(const float* sA, const float* sW, int ty, int tx, float acc[4][4]) {
#pragma unroll 2
  for (int k4 = 0; k4 < K; k4 += 4) {
    float4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(sA + (ty * 4 + i) * LD + k4);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(sW + (k4 + kk) * LDW + tx * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
        acc[i][0] = fmaf(av, w.x, acc[i][0]);
        acc[i][1] = fmaf(av, w.y, acc[i][1]);
        acc[i][2] = fmaf(av, w.z, acc[i][2]);
        acc[i][3] = fmaf(av, w.w, acc[i][3]);
      }
    }
  }
}

static inline int desco_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
