// Optional per-kernel-class device timing (CUDA events on the launching stream), used by bench.py to report
// the roofline of each kernel class from inside the timed region.  Off by default: zero overhead.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <mutex>
#include <vector>

#include "zv_common.h"

namespace zv {
namespace {
struct Rec { int cls; cudaEvent_t a, b; };
std::mutex g_mu;
bool g_on = false;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}
int num_sms() {
  static std::atomic<int> cache[64];
  const int dev = current_device() & 63;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (!n) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }

KernelTimer::KernelTimer(int cls, void* stream) : stream_(stream), idx_(-1) {
  if (!g_on) return;
  std::lock_guard<std::mutex> l(g_mu);
  if (g_recs.size() >= 65536) return;
  Rec r{cls, get_event(), get_event()};
  cudaEventRecord(r.a, static_cast<cudaStream_t>(stream));
  idx_ = (int)g_recs.size();
  g_recs.push_back(r);
}
KernelTimer::~KernelTimer() {
  if (idx_ < 0) return;
  std::lock_guard<std::mutex> l(g_mu);
  cudaEventRecord(g_recs[idx_].b, static_cast<cudaStream_t>(stream_));
}
}  // namespace zv

extern "C" {
void zv_timing_enable(int on) {
  std::lock_guard<std::mutex> l(zv::g_mu);
  zv::g_on = on != 0;
}
// Drops all records (events go back to the pool).
void zv_timing_reset(void) {
  std::lock_guard<std::mutex> l(zv::g_mu);
  for (auto& r : zv::g_recs) { zv::g_pool.push_back(r.a); zv::g_pool.push_back(r.b); }
  zv::g_recs.clear();
}
// Sum of elapsed milliseconds and launch count of one kernel class (waits for the recorded events).
int zv_timing_read(int cls, double* ms_total, int64_t* count) {
  std::lock_guard<std::mutex> l(zv::g_mu);
  double ms = 0;
  int64_t n = 0;
  for (auto& r : zv::g_recs) {
    if (r.cls != cls) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return zv::fail(ZV_ECUDA, "zv_timing_read: event sync failed");
    float t = 0;
    cudaEventElapsedTime(&t, r.a, r.b);
    ms += t;
    ++n;
  }
  if (ms_total) *ms_total = ms;
  if (count) *count = n;
  return ZV_OK;
}
}
