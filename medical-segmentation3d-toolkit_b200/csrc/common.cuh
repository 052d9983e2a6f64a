// common.cuh - shared helpers for the seg3d_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/seg3d_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "seg3d_b200 is written for sm_100a (B200) only"
#endif

void seg3d_set_error(const char* fmt, ...);

#define SEG3D_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) { seg3d_set_error(__VA_ARGS__); return SEG3D_EINVAL; } \
  } while (0)

#define SEG3D_CHECK_LAUNCH(name)                                                        \
  do {                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) {                                                           \
      seg3d_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));     \
      return SEG3D_ECUDA;                                                               \
    }                                                                                   \
  } while (0)

static inline int seg3d_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- storage-type helpers: activations are float, __half or __nv_bfloat16; math is fp32 ----
template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = a; *reinterpret_cast<float4*>(p + 4) = b; }
  __device__ __forceinline__ void get(float* f) const { f[0]=a.x; f[1]=a.y; f[2]=a.z; f[3]=a.w; f[4]=b.x; f[5]=b.y; f[6]=b.z; f[7]=b.w; }
  __device__ __forceinline__ void set(const float* f) { a = make_float4(f[0],f[1],f[2],f[3]); b = make_float4(f[4],f[5],f[6],f[7]); }
  __device__ __forceinline__ void zero() { a = make_float4(0,0,0,0); b = a; }
};
template <> struct Vec8<__half> {
  uint4 v;
  __device__ __forceinline__ void load(const __half* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__half* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const {
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __half22float2(h[i]); f[2*i] = t.x; f[2*i+1] = t.y; }
  }
  __device__ __forceinline__ void set(const float* f) {
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2*i], f[2*i+1]);
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0,0,0,0); }
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
  __device__ __forceinline__ void get(float* f) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2*i] = t.x; f[2*i+1] = t.y; }
  }
  __device__ __forceinline__ void set(const float* f) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2*i], f[2*i+1]);
  }
  __device__ __forceinline__ void zero() { v = make_uint4(0,0,0,0); }
};

// 16 consecutive channels from fp32 registers.  `wide` (pointer 32-byte aligned): one 256-bit store for the 2-byte
// types, so a 32-byte sector is written by a single request instead of two half-sector ones.
template <typename T>
__device__ __forceinline__ void store16(T* dst, const float* f, bool wide) {
  Vec8<T> a, b;
  a.set(f); b.set(f + 8);
  if constexpr (sizeof(T) == 2) {
    if (wide) {
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a.v.x), "r"(a.v.y), "r"(a.v.z), "r"(a.v.w),
                   "r"(b.v.x), "r"(b.v.y), "r"(b.v.z), "r"(b.v.w) : "memory");
      return;
    }
  }
  a.store(dst); b.store(dst + 8);
}
__host__ __device__ __forceinline__ bool wide_ok(const void* base, int ld_elems, int elem_bytes) {
  return (((uintptr_t)base) % 32 == 0) && ((ld_elems * elem_bytes) % 32 == 0);
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

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

// Block-wide sum of (s, ss) followed by one double atomicAdd pair into stats[0], stats[1].
// Must be called by every thread of the block; red is >= 2*32 floats of shared memory.
__device__ __forceinline__ void block_stats_atomic(float s, float ss, double* stats, float* red) {
  s = warp_sum(s); ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { red[warp] = s; red[32 + warp] = ss; }
  __syncthreads();
  if (warp == 0) {
    double a = lane < nw ? (double)red[lane] : 0.0, b = lane < nw ? (double)red[32 + lane] : 0.0;
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { atomicAdd(stats, a); atomicAdd(stats + 1, b); }
  }
}

// mean / rstd of GroupNorm(1,C) from the per-sample double sums (biased variance).
__device__ __forceinline__ void gn_mean_rstd(const double* stats, double count, float eps, float& mean, float& rstd) {
  const double m = stats[0] / count;
  double var = stats[1] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)eps));
}

#define SEG3D_DISPATCH_DTYPE(dtype, T, ...)                                    \
  switch (dtype) {                                                             \
    case SEG3D_F32: { using T = float; __VA_ARGS__; break; }                   \
    case SEG3D_F16: { using T = __half; __VA_ARGS__; break; }                  \
    case SEG3D_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }          \
    default: seg3d_set_error("unknown dtype %d", (int)dtype); return SEG3D_EINVAL; \
  }
