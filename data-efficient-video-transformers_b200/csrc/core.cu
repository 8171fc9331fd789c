// Library-wide host helpers: error string, device queries, launch checking.
#include <stdarg.h>
#include <stdio.h>

#include <mutex>

#include "tvt_common.cuh"

namespace tvt {

static thread_local char g_err[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int n = 0;
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  });
  return n;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return TVT_ECUDA;
  }
  return TVT_OK;
}

int require_sm100() {
  static int status = 1;  // 1 = not probed yet
  static std::once_flag once;
  std::call_once(once, [] {
    int dev = 0, major = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
      set_last_error("no usable CUDA device: %s", cudaGetErrorString(e));
      cudaGetLastError();
      status = TVT_ECUDA;
    } else {
      status = major == 10 ? TVT_OK : TVT_EARCH;
    }
  });
  if (status == TVT_EARCH) set_last_error("device is not sm_100 (compute capability 10.x required)");
  if (status == TVT_ECUDA && g_err[0] == 0) set_last_error("no usable CUDA device");
  return status;
}

// Device-resident step counter folded into every dropout seed (tvt_common.cuh mix_seed); nullptr = none.
static const unsigned long long* g_seed_source = nullptr;
const unsigned long long* seed_source() { return g_seed_source; }

__global__ void step_counter_kernel(unsigned long long* ctr, int n) {
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = 0; i < n; ++i) ctr[i] += 1ull;
}

}  // namespace tvt

extern "C" int tvt_set_seed_source(const void* device_counter) {
  tvt::g_seed_source = static_cast<const unsigned long long*>(device_counter);
  return TVT_OK;
}

extern "C" int tvt_step_counter_advance(void* device_counters, int count, void* stream) {
  using namespace tvt;
  TVT_REQUIRE(device_counters != nullptr && count >= 1 && count <= 16, "tvt_step_counter_advance: bad arguments");
  int rc = require_sm100();
  if (rc != TVT_OK) return rc;
  step_counter_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<unsigned long long*>(device_counters), count);
  return check_launch("tvt_step_counter_advance");
}

extern "C" const char* tvt_last_error(void) { return tvt::g_err; }
extern "C" int tvt_version(void) { return 100; }
extern "C" int tvt_device_check(void) { return tvt::require_sm100(); }
