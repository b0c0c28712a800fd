// Shared helpers for the desco_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define DESCO_OK 0
#define DESCO_EINVAL (-22)
#define DESCO_ENOMEM (-12)
#define DESCO_ECUDA (-5)
#define DESCO_ERANGE (-34)

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
