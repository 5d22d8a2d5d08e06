#include "scan.cuh"

namespace agnn {
namespace {

__device__ __forceinline__ int block_excl_scan(int v, int& total, int* warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int wprefix = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    int sw = warp_sums[w];
    if (w < warp) wprefix += sw;
    tot += sw;
  }
  __syncthreads();
  total = tot;
  return wprefix + inc - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tiles(int32_t* data, int64_t n, int32_t* tile_sums) {
  __shared__ int warp_sums[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int carry = 0;
  for (int pass = 0; pass < kScanTile / kScanThreads; ++pass) {
    const int64_t k = base + pass * kScanThreads + threadIdx.x;
    const int v = k < n ? data[k] : 0;
    int total;
    const int pre = block_excl_scan(v, total, warp_sums);
    if (k < n) data[k] = carry + pre;
    carry += total;
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_sums(int32_t* tile_sums, int n_tiles) {
  __shared__ int warp_sums[kScanThreads / 32];
  int carry = 0;
  for (int t0 = 0; t0 < n_tiles; t0 += kScanThreads) {
    const int t = t0 + threadIdx.x;
    const int v = t < n_tiles ? tile_sums[t] : 0;
    int total;
    const int pre = block_excl_scan(v, total, warp_sums);
    if (t < n_tiles) tile_sums[t] = carry + pre;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_add(int32_t* data, int64_t n, const int32_t* tile_sums) {
  if (blockIdx.x == 0) return;
  const int add = tile_sums[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  for (int i = threadIdx.x; i < kScanTile; i += kScanThreads)
    if (base + i < n) data[base + i] += add;
}

}  // namespace

size_t scan_workspace_bytes(int64_t n) { return (size_t)(ceil_div(n, kScanTile) + 1) * sizeof(int32_t); }

int exclusive_scan_i32(int32_t* data, int64_t n, int32_t* tile_sums, cudaStream_t stream) {
  if (n <= 0) return AGNN_OK;
  const int tiles = (int)ceil_div(n, kScanTile);
  scan_tiles<<<tiles, kScanThreads, 0, stream>>>(data, n, tile_sums);
  if (tiles > 1) {
    scan_sums<<<1, kScanThreads, 0, stream>>>(tile_sums, tiles);
    scan_add<<<tiles, kScanThreads, 0, stream>>>(data, n, tile_sums);
  }
  return check_launch("exclusive_scan");
}

}  // namespace agnn
