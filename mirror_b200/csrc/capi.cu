// Library-level C ABI: error string, version, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace mb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static const unsigned long long* g_drop_epoch = nullptr;  // process-wide (one process per GPU), set around graph capture
const unsigned long long* drop_epoch_ptr() { return g_drop_epoch; }
}  // namespace mb

extern "C" const char* mirror_last_error(void) { return mb::g_err; }
extern "C" int mirror_abi_version(void) { return 1; }
extern "C" int mirror_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}
extern "C" int mirror_set_dropout_epoch(const void* device_counter) {
  mb::g_drop_epoch = reinterpret_cast<const unsigned long long*>(device_counter);
  return 0;
}
