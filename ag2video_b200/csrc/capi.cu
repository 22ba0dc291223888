// C-ABI plumbing shared by all entry points of libag2v_sm100a.so (include/ag2v.h).
#include "common.cuh"

namespace ag2v {

char* err_buf() {
  static thread_local char buf[1024] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 1024, fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int check_arch() {
  int dev = 0, major = 0;
  AG2V_CUDA(cudaGetDevice(&dev));
  AG2V_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(AG2V_ERR_ARCH, "libag2v_sm100a needs a compute-capability 10.x device (found %d.x)", major);
  return AG2V_OK;
}

}  // namespace ag2v

extern "C" const char* ag2v_last_error_string(void) { return ag2v::err_buf(); }
extern "C" int ag2v_version(void) { return 100; }
extern "C" int ag2v_check_device(void) { return ag2v::check_arch(); }
extern "C" int ag2v_sm_count(void) { return ag2v::sm_count(); }
extern "C" unsigned long long ag2v_launch_count(void) { return __atomic_load_n(&ag2v::g_launches, __ATOMIC_RELAXED); }
