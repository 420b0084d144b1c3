// Shared helpers for libdf3d_b200.so (error convention, launch checks, warp reductions).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/df3d_b200.h"

namespace df3d {

// thread-local message behind df3d_last_error()
void set_error(const char* fmt, ...);

#define DF3D_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::df3d::set_error(__VA_ARGS__);      \
      return (code);                       \
    }                                      \
  } while (0)

// map a CUDA error to DF3D_ECUDA without leaving it sticky
#define DF3D_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ::df3d::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),     \
                        __FILE__, __LINE__);                                         \
      (void)cudaGetLastError();                                                      \
      return DF3D_ECUDA;                                                             \
    }                                                                                \
  } while (0)

#define DF3D_LAUNCH_CHECK(name)                                                      \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      ::df3d::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));   \
      return DF3D_ECUDA;                                                             \
    }                                                                                \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace df3d
