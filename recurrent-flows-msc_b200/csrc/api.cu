// librfk: error reporting and device queries.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace rfk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;  // B200
  }
  return n;
}

// Programmatic dependent launch.  RFK_PDL=1 / 0 forces it on / off for every launch; otherwise it follows rfk_set_pdl(),
// which the host side switches on around the sampling direction (latency-bound chains of small launches: -6 % per frame)
// and leaves off elsewhere (the training step loses 4 % with it).
static int g_pdl_mode = 0;
bool pdl_enabled() {
  static int env = -2;
  if (env == -2) {
    const char* e = getenv("RFK_PDL");
    env = e ? (atoi(e) != 0 ? 1 : 0) : -1;
  }
  return env >= 0 ? env == 1 : g_pdl_mode == 1;
}

}  // namespace rfk

extern "C" int rfk_version(void) { return RFK_VERSION; }

extern "C" int rfk_set_pdl(int on) {
  const int prev = rfk::g_pdl_mode;
  rfk::g_pdl_mode = on ? 1 : 0;
  return prev;
}

extern "C" const char* rfk_last_error(void) { return rfk::g_err; }

extern "C" int rfk_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  cudaDeviceProp prop;
  if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) {
    rfk::set_error("rfk_device_info: %s", cudaGetErrorString(e));
    return RFK_ECUDA;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return RFK_OK;
}
