#include "host_util.h"
#include <stdlib.h>
#include "../../include/ctunet_b200.h"

namespace ctu {

static std::atomic<int64_t> g_launches{0};

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CTU_PDL"); return e ? atoi(e) != 0 : false; }();
  return on;
}

static std::atomic<int> g_sm_limit{0};
int persistent_sms(int device_sms) {
  const int lim = g_sm_limit.load(std::memory_order_relaxed);
  return (lim > 0 && lim < device_sms) ? lim : device_sms;
}
void set_sm_limit(int sms) { g_sm_limit.store(sms, std::memory_order_relaxed); }

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

tma_encode_fn tma_encoder() {
  static tma_encode_fn fn = []() -> tma_encode_fn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<tma_encode_fn>(p);
  }();
  return fn;
}

}  // namespace ctu

namespace ctu { void set_sm_limit(int); }
extern "C" void ctu_set_persistent_sm_limit(int sms) { ctu::set_sm_limit(sms); }

extern "C" int64_t ctu_launch_count(void) { return ctu::g_launches.load(std::memory_order_relaxed); }

extern "C" int ctu_device_ok(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  if (prop.major != 10) return 0;
  return ctu::tma_encoder() != nullptr ? 1 : 0;
}

extern "C" const char* ctu_version(void) { return "ctunet_b200 0.1 (sm_100a)"; }
