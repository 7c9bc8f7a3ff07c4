// Shared helpers for libviai_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include "../../include/viai_b200.h"

namespace viai {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
// Where the most recent launch's sweep over its tensors ended: 0 = at the tail (a front-to-back sweep: every kernel unless it
// says otherwise), 1 = at the front.  Read by the normalisation passes (norm_act.cu), which start where the previous kernel
// ended -- on the lines that are still in L2.  A performance hint only (results do not depend on it).
extern std::atomic<int> g_sweep_end;

inline cudaStream_t STR(viai_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define VIAI_REQUIRE(cond, ...)                                   \
  do {                                                            \
    if (!(cond)) {                                                \
      viai::set_error(__VA_ARGS__);                               \
      return VIAI_ERR_ARG;                                        \
    }                                                             \
  } while (0)

#define VIAI_CUDA(expr)                                                            \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      viai::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return VIAI_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

// after every <<< >>>
#define VIAI_LAUNCHED()                                                            \
  do {                                                                             \
    viai::g_launches.fetch_add(1, std::memory_order_relaxed);                      \
    viai::g_sweep_end.store(0, std::memory_order_relaxed);                         \
    cudaError_t _e = cudaPeekAtLastError();                                        \
    if (_e != cudaSuccess) {                                                       \
      (void)cudaGetLastError();                                                    \
      viai::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return VIAI_ERR_CUDA;                                                        \
    }                                                                              \
  } while (0)

constexpr int kNumSMs = 148;

__host__ __device__ inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  switch (act) {
    case VIAI_ACT_RELU: return v > 0.f ? v : 0.f;
    case VIAI_ACT_LRELU: return v > 0.f ? v : v * slope;
    case VIAI_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    default: return v;
  }
}
// derivative wrt the pre-activation value `v`
__device__ __forceinline__ float act_grad(float v, int act, float slope) {
  switch (act) {
    case VIAI_ACT_RELU: return v > 0.f ? 1.f : 0.f;
    case VIAI_ACT_LRELU: return v > 0.f ? 1.f : slope;
    case VIAI_ACT_SIGMOID: { float s = 1.f / (1.f + expf(-v)); return s * (1.f - s); }
    default: return 1.f;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace viai
