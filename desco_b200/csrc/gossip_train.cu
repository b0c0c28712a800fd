// Training step of the gossip model - the small kernels around the dense / aggregation primitives (sm_100a, fp32).
//
// Replaces (reference paths relative to fuvty/DeSCo @ 4508f7a):
//   subgraph_counting/lightning_model.py:585-608  GossipCountingModel.train_forward (loop over queries)
//   subgraph_counting/lightning_model.py:630-635  criterion: log2(|count - truth| + 1), summed over nodes and queries
//   subgraph_counting/gnn_model.py:326-348        GossipConv message / update, through autograd
//   subgraph_counting/gnn_model.py:294-301        lin_gate and its gradient
// The matrix work of a step runs on desco_train_dense / desco_train_wgrad (csrc/train.cu) and desco_spmm_sum
// (csrc/conv.cu, whose adjoint on a symmetric edge set is the same launch with the two gate weights swapped); the
// kernels here are the glue that has no torch equivalent on raw device pointers: the gated mix with a device-resident
// gate, the gate's gradient, dropout masks, and the loss.
#include <math.h>

#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

// out = g * a + (1 - g) * b, g = *gate (device scalar)
__global__ void gated_mix_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gate,
                                 float* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = *gate;
  out[i] = g * a[i] + (1.f - g) * b[i];
}

// *dgate += sum_i d[i] * (a[i] - b[i])   (fixed block order within a launch is not needed: one fp32 atomic per block)
__global__ void gate_grad_kernel(const float* __restrict__ d, const float* __restrict__ a, const float* __restrict__ b,
                                 long long n, float* __restrict__ dgate) {
  __shared__ float s[32];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc = fmaf(d[i], a[i] - b[i], acc);
  acc = warp_sum(acc);
  if (lane_id() == 0) s[warp_id()] = acc;
  __syncthreads();
  if (warp_id() == 0) {
    float v = lane_id() < (blockDim.x >> 5) ? s[lane_id()] : 0.f;
    v = warp_sum(v);
    if (lane_id() == 0) atomicAdd(dgate, v);
  }
}

// gradient of gate = leaky(sigmoid(w2 . sigmoid(W1 q + b1) + b2)) w.r.t. the lin_gate parameters (+=); one CTA, H threads
__global__ void gate_backward_kernel(const float* __restrict__ q, int E, const float* __restrict__ W1,
                                     const float* __restrict__ b1, int H, const float* __restrict__ w2,
                                     const float* __restrict__ b2, const float* __restrict__ dgate,
                                     float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dw2,
                                     float* __restrict__ db2) {
  __shared__ float s_red[32];
  const int h = threadIdx.x;
  float hid = 0.f;
  if (h < H) {
    float acc = b1[h];
    for (int k = 0; k < E; ++k) acc = fmaf(q[k], W1[(size_t)h * E + k], acc);
    hid = sigmoidf_(acc);
  }
  float part = h < H ? hid * w2[h] : 0.f;
  part = warp_sum(part);
  if (lane_id() == 0) s_red[warp_id()] = part;
  __syncthreads();
  float s = b2[0];
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_red[w];
  const float g = sigmoidf_(s);
  const float ds = (*dgate) * g * (1.f - g) * (g > 0.f ? 1.f : 0.01f);
  if (h < H) {
    dw2[h] += ds * hid;
    const float dz = ds * w2[h] * hid * (1.f - hid);
    db1[h] += dz;
    for (int k = 0; k < E; ++k) dW1[(size_t)h * E + k] += dz * q[k];
  }
  if (h == 0) db2[0] += ds;
}

// x[i] *= mask[i] ? scale : 0
__global__ void dropout_kernel(float* __restrict__ x, const uint8_t* __restrict__ mask, float scale, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = mask[i] ? x[i] * scale : 0.f;
}

// pred = c + out; loss += sum_i log2(|pred_i - y_i| + 1); dout_i = sign(pred_i - y_i) / ((|pred_i - y_i| + 1) ln 2)
__global__ void gossip_loss_kernel(const float* __restrict__ c, int ldc, const float* __restrict__ out, int ldo,
                                   const float* __restrict__ y, int ldy, int n, float* __restrict__ pred,
                                   float* __restrict__ dout, int ldd, float* __restrict__ loss) {
  __shared__ float s[32];
  float acc = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float p = c[(size_t)i * ldc] + out[(size_t)i * ldo];
    const float d = p - y[(size_t)i * ldy];
    if (pred) pred[i] = p;
    acc += log2f(fabsf(d) + 1.f);
    if (dout) dout[(size_t)i * ldd] = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) / ((fabsf(d) + 1.f) * 0.6931471805599453f);
  }
  acc = warp_sum(acc);
  if (lane_id() == 0) s[warp_id()] = acc;
  __syncthreads();
  if (warp_id() == 0) {
    float v = lane_id() < (blockDim.x >> 5) ? s[lane_id()] : 0.f;
    v = warp_sum(v);
    if (lane_id() == 0) atomicAdd(loss, v);
  }
}

}  // namespace

extern "C" {

int desco_gossip_gated_mix(const float* a, const float* b, const float* gate, float* out, int64_t n, void* stream) {
  if (n < 0) return DESCO_EINVAL;
  if (n == 0) return DESCO_OK;
  if (!a || !b || !gate || !out) return DESCO_EINVAL;
  desco_count_launches(1);
  gated_mix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, b, gate, out, (long long)n);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_gate_grad(const float* d, const float* a, const float* b, int64_t n, float* dgate, void* stream) {
  if (n < 0) return DESCO_EINVAL;
  if (n == 0) return DESCO_OK;
  if (!d || !a || !b || !dgate) return DESCO_EINVAL;
  desco_count_launches(1);
  long long blocks = (n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  gate_grad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d, a, b, (long long)n, dgate);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_gate_backward(const float* query_emb, int32_t emb_channels, const float* w1, const float* b1, int32_t hidden,
                               const float* w2, const float* b2, const float* dgate, float* dw1, float* db1, float* dw2,
                               float* db2, void* stream) {
  if (emb_channels < 1 || hidden < 1 || hidden > 1024) return DESCO_EINVAL;
  if (!query_emb || !w1 || !b1 || !w2 || !b2 || !dgate || !dw1 || !db1 || !dw2 || !db2) return DESCO_EINVAL;
  desco_count_launches(1);
  gate_backward_kernel<<<1, ((hidden + 31) / 32) * 32, 0, (cudaStream_t)stream>>>(query_emb, emb_channels, w1, b1, hidden, w2,
                                                                                  b2, dgate, dw1, db1, dw2, db2);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_train_dropout(float* x, const uint8_t* mask, float scale, int64_t n, void* stream) {
  if (n < 0) return DESCO_EINVAL;
  if (n == 0) return DESCO_OK;
  if (!x || !mask) return DESCO_EINVAL;
  desco_count_launches(1);
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, mask, scale, (long long)n);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

int desco_gossip_loss(const float* c, int32_t ldc, const float* out, int32_t ldo, const float* y, int32_t ldy, int32_t n,
                      float* pred, float* dout, int32_t ldd, float* loss, void* stream) {
  if (n < 0) return DESCO_EINVAL;
  if (n == 0) return DESCO_OK;
  if (!c || !out || !y || !loss) return DESCO_EINVAL;
  desco_count_launches(1);
  int blocks = (n + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  gossip_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(c, ldc, out, ldo, y, ldy, n, pred, dout, ldd, loss);
  DESCO_LAUNCH_CHECK();
  return DESCO_OK;
}

}  // extern "C"
