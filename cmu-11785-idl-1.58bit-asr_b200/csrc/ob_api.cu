// ob_api.cu - library-level entry points: version, per-thread error string, device check.
#include <stdarg.h>

#include <atomic>

#include "ob_common.cuh"

namespace ob {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int g_pdl = 1;
int pdl_enabled() { return g_pdl; }
void set_pdl(int on) { g_pdl = on != 0; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_device() {
  static thread_local int cached_dev = -1;
  static thread_local int cached_rc = OB_OK;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("cudaGetDevice failed: no CUDA device");
    return OB_ERR_CUDA;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cached_dev = dev;
  cached_rc = OB_OK;
  if (major != 10) {
    set_error("libonebit is built for sm_100a (B200); device %d has compute capability major %d", dev, major);
    cached_rc = OB_ERR_ARCH;
  }
  return cached_rc;
}

}  // namespace ob

extern "C" int ob_version(void) { return OB_VERSION; }
extern "C" int64_t ob_launch_count(void) { return static_cast<int64_t>(ob::g_launches.load()); }
extern "C" const char* ob_last_error_string(void) { return ob::g_err; }
