// Warp-level MMA issue rate on sm_100a: independent accumulator chains of mma.sync m16n8k8 tf32 and m16n8k16 bf16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int KIND, int CHAINS>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, uint32_t seed) {
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) d[c][e] = 0.f;
  uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int CHAINS>
void run(const char* name, int ctas_per_sm, float* out) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  rate_kernel<KIND, CHAINS><<<sms * ctas_per_sm, 256>>>(out, 100, 1);
  cudaEventRecord(a);
  rate_kernel<KIND, CHAINS><<<sms * ctas_per_sm, 256>>>(out, iters, 1);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mmas_per_smsp = (double)iters * CHAINS * 2 * ctas_per_sm;  // 8 warps per CTA over 4 sub-partitions
  const double cycles = ms * 1e-3 * khz * 1e3;
  printf("%-28s chains %d warps/SM %2d: %.3f ms, %.2f cycles per MMA per sub-partition (at %d MHz)\n", name, CHAINS,
         8 * ctas_per_sm, ms, cycles / mmas_per_smsp, khz / 1000);
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
  run<0, 8>("m16n8k8 tf32", 1, out);
  run<0, 8>("m16n8k8 tf32", 2, out);
  run<0, 2>("m16n8k8 tf32", 2, out);
  run<0, 1>("m16n8k8 tf32 (latency)", 1, out);
  run<1, 8>("m16n8k16 bf16", 1, out);
  run<1, 8>("m16n8k16 bf16", 2, out);
  run<1, 1>("m16n8k16 bf16 (latency)", 1, out);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
