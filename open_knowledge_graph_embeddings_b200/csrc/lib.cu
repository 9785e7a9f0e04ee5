// Library-level glue of libokge_b200.so: ABI version, per-thread error message, device check.
#include "okge_common.cuh"

#include <stdlib.h>
#include <string.h>

namespace okge {

namespace {
thread_local char g_last_error[512] = "";
}

void set_last_error(const char* file, int line, const char* msg) {
  const char* base = strrchr(file, '/');
  snprintf(g_last_error, sizeof(g_last_error), "%s:%d: %s", base ? base + 1 : file, line, msg);
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200
  }
  return cached;
}

bool pdl_enabled() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("OKGE_PDL");
    cached = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return cached != 0;
}

}  // namespace okge

extern "C" int okge_abi_version(void) { return OKGE_ABI_VERSION; }

extern "C" const char* okge_last_error(void) { return okge::g_last_error; }

extern "C" int okge_device_check(void) {
  // The tensor-map encoder is a DRIVER entry point and needs the primary context bound to the calling thread. Host
  // frameworks call us from worker threads (autograd) whose first CUDA call may be ours: bind it once per thread with a
  // runtime call (cudaFree(0) initialises / attaches the primary context of the current device).
  static thread_local bool context_bound = false;
  if (!context_bound) {
    cudaFree(nullptr);
    context_bound = true;
  }
  static int cached = -1;
  if (cached == OKGE_OK) return cached;
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    okge::set_last_error(__FILE__, __LINE__, "no CUDA device available");
    return OKGE_ERR_CUDA;
  }
  if (major != 10) {
    okge::set_last_error(__FILE__, __LINE__, "libokge_b200 requires a compute-capability 10.x GPU (B200, sm_100a)");
    return OKGE_ERR_UNSUPPORTED;
  }
  cached = OKGE_OK;
  return cached;
}
