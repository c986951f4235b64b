// Shared device/host helpers for the vinet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "vinet_b200.h"

namespace vinet {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
extern std::atomic<long long> g_up2_launches;   // tensor-core launches whose producer warps interpolated a VINET_XF_UP2 source
void note_kernel(const char* name);   // which CUDA kernel the last conv / wgrad call of this thread launched (vinet_last_kernel)

#define VINET_CHECK(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vinet::set_error(__VA_ARGS__);  \
      return -1;                        \
    }                                   \
  } while (0)

// call after every launch: counts it and converts launch-configuration errors into return codes
#define VINET_LAUNCH_OK(name)                                                        \
  do {                                                                               \
    ::vinet::g_launches.fetch_add(1, std::memory_order_relaxed);                     \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      ::vinet::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return -2;                                                                     \
    }                                                                                \
  } while (0)

__host__ __device__ static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ------------------------------------------------------------------ typed vector load/store
template <typename T>
struct Store;
template <>
struct Store<float> {
  static constexpr int dtype = VINET_F32;
};
template <>
struct Store<__nv_bfloat16> {
  static constexpr int dtype = VINET_BF16;
};

__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// load 8 consecutive elements (16B-aligned for bf16, 32B for fp32) as floats
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
  v[4] = bf16_lo(u.z); v[5] = bf16_hi(u.z); v[6] = bf16_lo(u.w); v[7] = bf16_hi(u.w);
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = bf16_lo(u.x); v[1] = bf16_hi(u.x); v[2] = bf16_lo(u.y); v[3] = bf16_hi(u.y);
}
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 u;
  u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
  u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ float load1(const float* p) { return *p; }
__device__ __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
__device__ __forceinline__ void store1(float* p, float v) { *p = v; }

// pending transform of a producer layer, applied on read
template <int V>
__device__ __forceinline__ void apply_xform(float (&v)[V], int xform, const float* __restrict__ scale,
                                            const float* __restrict__ shift, int c) {
  if (xform & 2) {
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      float4 s = __ldg(reinterpret_cast<const float4*>(scale + c + i));
      float4 h = __ldg(reinterpret_cast<const float4*>(shift + c + i));
      v[i + 0] = fmaf(v[i + 0], s.x, h.x); v[i + 1] = fmaf(v[i + 1], s.y, h.y);
      v[i + 2] = fmaf(v[i + 2], s.z, h.z); v[i + 3] = fmaf(v[i + 3], s.w, h.w);
    }
  }
  if (xform & 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], 0.f);
  }
}

// dispatch a storage dtype to a template argument
#define VINET_DISPATCH_DTYPE(dt, T, ...)                  \
  do {                                                    \
    if ((dt) == VINET_BF16) {                             \
      using T = __nv_bfloat16;                            \
      __VA_ARGS__;                                        \
    } else {                                              \
      using T = float;                                    \
      __VA_ARGS__;                                        \
    }                                                     \
  } while (0)

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

}  // namespace vinet
