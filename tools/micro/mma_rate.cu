// Microbenchmark (sm_100a): issue rate of the legacy warp-level mma.sync (TF32 m16n8k8, F16 m16n8k16) and the time a
// CTA needs to stream a tile from L2 with cp.async when 128 CTAs do the same (shared or private source).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mma_tf32(float* out, int iters) {
  float d[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3e800000u}, b0 = 0x3f800000u, b1 = 0x3f000000u;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  float s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mma_f16(float* out, int iters) {
  float d[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x38003800u, 0x34003400u}, b0 = 0x3c003c00u, b1 = 0x38003800u;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  float s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every CTA copies `bytes` (multiple of 16 KB) from src (+ blockIdx.x * bytes if !shared) into a 4-stage ring of 16 KB
__global__ void k_stream(const float4* src, float* out, int bytes, int shared_src, int reps) {
  extern __shared__ float4 ring[];   // 4 x 1024 float4
  const float4* s = src + (shared_src ? 0 : (size_t)blockIdx.x * (bytes / 16));
  const int chunks = bytes / 16384;
  float acc = 0.f;
  for (int r = 0; r < reps; ++r) {
    for (int q = 0; q < 3; ++q) {
      if (q < chunks)
        for (int e = threadIdx.x; e < 1024; e += blockDim.x)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&ring[q * 1024 + e])), "l"(s + q * 1024 + e) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int c = 0; c < chunks; ++c) {
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __syncthreads();
      const int n = c + 3;
      if (n < chunks)
        for (int e = threadIdx.x; e < 1024; e += blockDim.x)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(&ring[(n & 3) * 1024 + e])), "l"(s + n * 1024 + e) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      acc += ring[(c & 3) * 1024 + threadIdx.x].x;
    }
    __syncthreads();
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
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
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  const int iters = 4096;
  for (int threads : {128, 256, 512}) {
    float t1 = time_ms([&] { k_mma_tf32<<<148, threads>>>(out, iters); });
    float t2 = time_ms([&] { k_mma_f16<<<148, threads>>>(out, iters); });
    double c1 = t1 * 1e-3 * clk_khz * 1e3, c2 = t2 * 1e-3 * clk_khz * 1e3;
    double n = (double)(threads / 32) * 4 * iters;
    printf("warps/SM %2d: mma.sync tf32 m16n8k8 %.2f cycles per MMA per SM (%.0f FMA/clk/SM); f16 m16n8k16 %.2f cycles (%.0f FMA/clk/SM)\n",
           threads / 32, c1 / n, n * 1024 / c1, c2 / n, n * 2048 / c2);
  }
  float4* src;
  const int bytes = 176 * 1024 / 16384 * 16384;
  cudaMalloc(&src, (size_t)128 * bytes);
  cudaMemset(src, 0, (size_t)128 * bytes);
  cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int shared_src : {1, 0}) {
    const int reps = 50;
    float t = time_ms([&] { k_stream<<<128, 256, 65536>>>(src, out, bytes, shared_src, reps); });
    printf("128 CTAs x %d KB via cp.async (4 x 16 KB ring), %s source: %.2f us per pass = %.1f B/clk/SM, %.2f TB/s total\n",
           bytes / 1024, shared_src ? "one shared" : "private", t * 1e3 / reps,
           bytes / (t * 1e-3 / reps * clk_khz * 1e3), 128.0 * bytes / (t * 1e-3 / reps) / 1e12);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
