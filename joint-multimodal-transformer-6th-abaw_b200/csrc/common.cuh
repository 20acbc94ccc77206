// Shared device/host helpers for the jmt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

#include "../../include/jmt_b200.h"

namespace jmt {

extern std::atomic<int64_t> g_launch_count;
void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    cudaGetLastError();
    return JMT_ERR_CUDA;
  }
  return JMT_OK;
}

#define JMT_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      jmt::set_error(__VA_ARGS__);        \
      return JMT_ERR_INVALID;             \
    }                                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---- dtype helpers -----------------------------------------------------------------------
template <typename T> struct Vec8;   // 8 consecutive elements
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 r = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = r;
  }
};

__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  if (act == JMT_ACT_RELU) return fmaxf(x, 0.f);
  if (act == JMT_ACT_LEAKY_RELU) return x >= 0.f ? x : x * slope;
  return x;
}

// dispatch on a runtime dtype code
#define JMT_DISPATCH_DTYPE(code, T, ...)                                  \
  do {                                                                    \
    if ((code) == JMT_F32) { using T = float; __VA_ARGS__; }              \
    else if ((code) == JMT_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }\
    else { jmt::set_error("bad dtype code %d", (int)(code)); return JMT_ERR_INVALID; } \
  } while (0)

// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl() with JMT_PDL_ALL=1 may be scheduled while its
// predecessor in the stream is still draining: its CTAs become resident early and MUST call pdl_wait() before they touch global
// memory -- it blocks until the predecessor grid has completed and its writes are visible (a no-op without the attribute).
// Measured on B200 (round 2, profiles/ab_same_box_r2.txt, r2w): giving the memory-bound kernels the attribute -- even without
// an early trigger of their own -- costs 1.3 % of the step (14.69 k vs 14.89 k windows/s): their early-resident CTAs take
// issue slots and registers from the tail of the tensor-core kernel in front of them.  OFF by default; the tensor-core kernels
// keep PDL among themselves (prologue under the predecessor's tail).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_all_enabled() {
  static const bool on = []() {
    const char* a = getenv("JMT_PDL"); const char* b = getenv("JMT_PDL_ALL");
    return (a ? atoi(a) != 0 : true) && (b ? atoi(b) != 0 : false);
  }();
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_all_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks = kNumSMs * 16) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace jmt
