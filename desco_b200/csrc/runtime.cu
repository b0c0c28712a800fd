// Library-wide runtime services: launch accounting and per-kernel-group device timing (CUDA events recorded on the
// launching stream), used by bench.py for `gpu_launches` and the live roofline numbers.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "../../include/desco_b200.h"

namespace {
std::atomic<long long> g_launches{0};
std::atomic<bool> g_prof_on{false};
std::mutex g_mu;
struct Rec {
  cudaEvent_t a, b;
  int slot;
  int launches;
};
std::vector<Rec> g_recs;
}  // namespace

void desco_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

DescoProfScope::DescoProfScope(int slot, cudaStream_t s, int launches) : slot_(slot), launches_(launches), s_(s), on_(false) {
  desco_count_launches(launches);
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  on_ = true;
  cudaEventCreate(&a_);
  cudaEventCreate(&b_);
  cudaEventRecord(a_, s_);
}

DescoProfScope::~DescoProfScope() {
  if (!on_) return;
  cudaEventRecord(b_, s_);
  std::lock_guard<std::mutex> lk(g_mu);
  g_recs.push_back(Rec{a_, b_, slot_, launches_});
}

extern "C" {

int64_t desco_kernel_launches(void) { return g_launches.load(); }

int desco_profile_enable(int32_t on) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  g_prof_on.store(on != 0);
  return DESCO_OK;
}

int desco_profile_read(double* ms, int64_t* launches) {
  if (!ms || !launches) return DESCO_EINVAL;
  std::lock_guard<std::mutex> lk(g_mu);
  for (int i = 0; i < DESCO_PROF_SLOTS; ++i) {
    ms[i] = 0.0;
    launches[i] = 0;
  }
  for (auto& r : g_recs) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return DESCO_ECUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return DESCO_ECUDA;
    ms[r.slot] += t;
    launches[r.slot] += r.launches;
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_recs.clear();
  return DESCO_OK;
}

}  // extern "C"
