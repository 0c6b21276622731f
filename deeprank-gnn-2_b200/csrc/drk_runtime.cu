// Library-level state: thread-local error text, ABI version, launch counter.
#include "drk_common.cuh"

namespace drk {

static thread_local char t_error[768] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

}  // namespace drk

extern "C" {

int drk_abi_version(void) { return DRK_ABI_VERSION; }
const char* drk_last_error(void) { return drk::t_error; }
int64_t drk_launch_count(void) { return drk::g_launches.load(std::memory_order_relaxed); }

// Every entry point reads cudaGetLastError() after its launch, so a non-sticky error that ANOTHER user of the runtime in this process
// left pending (torch's own start-up probes) would be reported against our first kernel.  Called once when the library is loaded.
int drk_runtime_init(void) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) drk::set_error("pending CUDA error cleared at load: %s", cudaGetErrorString(e));
  return (int)e;
}

}  // extern "C"
