// Shared helpers for libagnn (sm_100a).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "agnn.h"

namespace agnn {

constexpr int kNumSM = 148;  // B200

// thread-local error text behind agnn_last_error()
char* error_buffer();
int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(AGNN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return AGNN_OK;
}

// AGNN_SYNC_CHECK=1 (debugging): wait for the stream after a launch and name the kernel that faulted.  Never set
// inside a CUDA-graph capture.
inline int sync_check(cudaStream_t st, const char* what) {
  static const bool on = [] { const char* e = getenv("AGNN_SYNC_CHECK"); return e && *e && *e != '0'; }();
  if (!on) return AGNN_OK;
  cudaError_t e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(AGNN_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return AGNN_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ __device__ inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- 16-byte vector access: T -> float[E] ---------------------------------
template <typename T>
struct Vec16;

template <>
struct Vec16<float> {
  static constexpr int E = 4;
  __device__ static __forceinline__ void load_nc(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  // the 16 bytes as loaded: what a kernel keeps in registers while many loads are in flight (see unpack)
  __device__ static __forceinline__ uint4 load_raw_nc(const float* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[4]) {
    v[0] = __uint_as_float(t.x); v[1] = __uint_as_float(t.y); v[2] = __uint_as_float(t.z); v[3] = __uint_as_float(t.w);
  }
  __device__ static __forceinline__ void load(const float* p, float (&v)[4]) {   // coherent (data of this launch's peers)
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int E = 8;
  __device__ static __forceinline__ void load_nc(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {  // bf16 -> f32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ uint4 load_raw_nc(const __nv_bfloat16* p) {
    return __ldg(reinterpret_cast<const uint4*>(p));
  }
  __device__ static __forceinline__ void unpack(const uint4& t, float (&v)[8]) {   // 4 registers in flight, 8 at use
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// ---- fp16 hi / lo operand pairs of agnn_gemm's F16X3 mode ---------------------------------------------
// One power-of-two scale per tensor, derived from its amax so that max |s x| lies in [2^13, 2^14) -- two
// binades below the fp16 maximum, which leaves room for producers that can exceed the amax they were given by
// up to 4x (a mean with a self term: 2x).  amax == 0, denormal, inf or nan: s = 1.  |log2 s| <= 100.
__host__ __device__ inline float f16_scale_of(float amax) {
#ifdef __CUDA_ARCH__
  const uint32_t bits = __float_as_uint(amax) & 0x7fffffffu;
#else
  uint32_t bits;
  memcpy(&bits, &amax, 4);
  bits &= 0x7fffffffu;
#endif
  const int e = (int)(bits >> 23);            // biased exponent: floor(log2 amax) + 127
  if (e == 0 || e == 255) return 1.f;
  int k = 13 - (e - 127);
  k = k > 100 ? 100 : (k < -100 ? -100 : k);
  const uint32_t sb = (uint32_t)(k + 127) << 23;
#ifdef __CUDA_ARCH__
  return __uint_as_float(sb);
#else
  float s;
  memcpy(&s, &sb, 4);
  return s;
#endif
}
// 4 values -> 4 fp16 hi + 4 fp16 lo (8 bytes each) of s * v
__device__ __forceinline__ void f16_pair4(const float (&v)[4], float s, uint2& hi, uint2& lo) {
  __half h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x = v[e] * s;
    h[e] = __float2half_rn(x);
    l[e] = __float2half_rn(x - __half2float(h[e]));
  }
  __half2 h01 = __halves2half2(h[0], h[1]), h23 = __halves2half2(h[2], h[3]);
  __half2 l01 = __halves2half2(l[0], l[1]), l23 = __halves2half2(l[2], l[3]);
  hi = make_uint2(*reinterpret_cast<uint32_t*>(&h01), *reinterpret_cast<uint32_t*>(&h23));
  lo = make_uint2(*reinterpret_cast<uint32_t*>(&l01), *reinterpret_cast<uint32_t*>(&l23));
}

// ---- counter-based dropout mask --------------------------------------------------------------------
// keep(element) is a pure function of (seed, step, stream id, element index): the backward recomputes the mask
// instead of reading one, and a CUDA-graph replay draws new masks because `step` is read from device memory.
// One splitmix64 draw decides four consecutive elements (16 bits each): keep iff bits >= p * 2^16.
struct DropoutState {           // device memory: {seed, step}; agnn_dropout_advance increments step
  uint64_t seed, step;
};
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ uint64_t dropout_key(const uint64_t* state, uint32_t stream_id) {
  return mix64(mix64(state[0] + 0x9E3779B97F4A7C15ull) ^ (state[1] * 0xD1B54A32D192ED03ull + stream_id));
}
// bits for the 4 elements starting at linear index `idx4 * 4`
__device__ __forceinline__ uint64_t dropout_bits(uint64_t key, uint64_t idx4) { return mix64(key ^ (idx4 * 0x9E3779B97F4A7C15ull)); }
__device__ __forceinline__ bool dropout_keep(uint64_t bits, int e, uint32_t threshold) {
  return ((uint32_t)(bits >> (16 * e)) & 0xffffu) >= threshold;
}
__host__ __device__ inline uint32_t dropout_threshold(float p) {
  const float t = p * 65536.f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : (uint32_t)(t + 0.5f));
}

}  // namespace agnn
