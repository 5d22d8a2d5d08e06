// Microbenchmark: issue rate of FFMA (3-register) vs FFMA2 (fma.rn.f32x2) and of LDS.128 broadcast patterns on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate ffma_rate.cu && ./ffma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

template <int CH>
__global__ void k_ffma(float* out, int iters, float b, float c) {
  float a[CH];
  for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = fma1(a[i], b, c);
  float s = 0;
  for (int i = 0; i < CH; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_ffma2(float* out, int iters, uint64_t b, uint64_t c) {
  uint64_t a[CH];
  for (int i = 0; i < CH; ++i) a[i] = (uint64_t)__float_as_uint(threadIdx.x * 1e-3f + i) * 0x100000001ull;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = fma2(a[i], b, c);
  uint64_t s = 0;
  for (int i = 0; i < CH; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((uint32_t)s ^ (uint32_t)(s >> 32));
}
// LDS.128 where the 32 lanes read `distinct` different 16-byte words (stride 16 B) -- broadcast within groups
__global__ void k_lds(float* out, int iters, int distinct) {
  __shared__ float4 s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = lane % distinct;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 v = s[(idx + 32 * i) & 1023];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    idx = (idx + (int)acc.w) & 1023;   // acc.w stays a multiple of 3*8: keeps the pattern, defeats hoisting
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  cudaEventRecord(e0);
  f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  const int ctas = 148, iters = 8192;
  float* out;
  cudaMalloc(&out, 148 * 1024 * 4 * sizeof(float));
  const uint64_t b2 = 0x3f8000013f800001ull, c2 = 0x3400000034000000ull;
  for (int threads : {128, 256, 512, 1024}) {
    float t1 = time_ms([&] { k_ffma<16><<<ctas, threads>>>(out, iters, 1.0000001f, 1e-7f); });
    float t2 = time_ms([&] { k_ffma2<16><<<ctas, threads>>>(out, iters, b2, c2); });
    double cyc1 = t1 * 1e-3 * clk_khz * 1e3, cyc2 = t2 * 1e-3 * clk_khz * 1e3;
    double f1 = (double)threads * 16 * iters / cyc1, f2 = (double)threads * 32 * iters / cyc2;
    printf("threads/SM %4d: FFMA %.3f ms = %.1f FMA/clk/SM;  FFMA2 %.3f ms = %.1f FMA/clk/SM (nominal clock %d MHz)\n",
           threads, t1, f1, t2, f2, clk_khz / 1000);
  }
  for (int distinct : {1, 4, 8, 16, 32}) {
    float t = time_ms([&] { k_lds<<<ctas, 256>>>(out, 2048, distinct); });
    double cyc = t * 1e-3 * clk_khz * 1e3;
    printf("LDS.128, %2d distinct 16-byte words per warp: %.2f cycles per warp-level load per SM\n", distinct,
           cyc / (8.0 * 2048 * 8));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
