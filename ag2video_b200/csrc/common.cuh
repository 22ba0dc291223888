// Shared helpers for the sm_100a kernels behind the C ABI (include/ag2v.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

namespace ag2v {

// Thread-local last-error text returned by ag2v_last_error_string().
char* err_buf();
int fail(int code, const char* fmt, ...);

enum : int {
  AG2V_OK = 0,
  AG2V_ERR_ARG = -1,      // bad shape / null pointer / alignment
  AG2V_ERR_CUDA = -2,     // CUDA runtime error (text in last error string)
  AG2V_ERR_ARCH = -3,     // not an sm_100 device
  AG2V_ERR_UNSUPPORTED = -4,
};

#define AG2V_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) return ::ag2v::fail(::ag2v::AG2V_ERR_ARG, __VA_ARGS__); \
  } while (0)

#define AG2V_CUDA(call)                                                         \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess)                                                     \
      return ::ag2v::fail(::ag2v::AG2V_ERR_CUDA, "%s:%d %s -> %s", __FILE__,    \
                          __LINE__, #call, cudaGetErrorString(e__));            \
  } while (0)

// Every kernel launch goes through one of these two, so ag2v_launch_count() is
// the number of kernels this library has enqueued (bench.py reports it).
void count_launch();
#define AG2V_LAUNCH_CHECK()            \
  do {                                 \
    ::ag2v::count_launch();            \
    AG2V_CUDA(cudaGetLastError());     \
  } while (0)
#define AG2V_COOP_LAUNCH(kernel, grid, block, args, smem, stream)                                    \
  do {                                                                                               \
    ::ag2v::count_launch();                                                                          \
    AG2V_CUDA(cudaLaunchCooperativeKernel((void*)(kernel), (grid), (block), (args), (smem), (stream))); \
  } while (0)

int sm_count();                 // SMs of the current device (cached per device)
int check_arch();               // AG2V_OK iff the current device is sm_100

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ag2v
