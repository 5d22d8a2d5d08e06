// Onset-wise decode: what the reference does with the softmaxed task logits right after the heads
// (onsetwise_logit_aggregation, analysisgnn/models/analysis.py:44-101; called from predict, :1588).
//
//   1. per Roman-numeral task: mean over the onset edges with the self term (agnn_gather_reduce, as onset pooling),
//      softmax, selection of the rows with a valid label, softmax again        -> agnn_softmax2_rows
//   2. one representative row per run of equal onsets                          -> agnn_run_heads (flag, scan, scatter)
//   3. arg-max per representative, segments where it changes                   -> agnn_row_argmax, agnn_run_heads again
//   4. every note whose onset lies in segment i (all but the last segment) takes the distribution of the
//      segment's first onset                                                   -> agnn_decode_assign
//
// The reference runs step 4 as a Python loop over the change points with one boolean mask over all notes per
// segment (O(segments x notes)); here each note finds its segment by binary search over the change-point onsets.
// Integer work (2, 3, 4's search) is exact; rows are copied bit for bit.  HBM / latency bound; no atomics.
#include "common.cuh"
#include "scan.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// y[i] = softmax(softmax(x[rows ? rows[i] : i])); a lane keeps up to kPerLane columns in registers
constexpr int kPerLane = 8;   // cols <= 256
__global__ void __launch_bounds__(kThreads) softmax2_kernel(const float* __restrict__ x, int64_t ld_x,
                                                             const int32_t* __restrict__ rows, int64_t n_out, int cols,
                                                             float* __restrict__ y, int64_t ld_y) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + warp; i < n_out; i += (int64_t)gridDim.x * kWarps) {
    const float* xr = x + (rows ? (int64_t)rows[i] : i) * ld_x;
    float v[kPerLane];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      float m = -INFINITY;
#pragma unroll
      for (int q = 0; q < kPerLane; ++q) {
        const int c = q * 32 + lane;
        if (pass == 0) v[q] = c < cols ? xr[c] : -INFINITY;
        m = fmaxf(m, v[q]);
      }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < kPerLane; ++q) {
        const int c = q * 32 + lane;
        v[q] = c < cols ? expf(v[q] - m) : 0.f;
        s += v[q];
      }
      s = warp_sum(s);
#pragma unroll
      for (int q = 0; q < kPerLane; ++q) {
        const int c = q * 32 + lane;
        v[q] = c < cols ? v[q] / s : -INFINITY;
      }
    }
    float* yr = y + i * ld_y;
#pragma unroll
    for (int q = 0; q < kPerLane; ++q) {
      const int c = q * 32 + lane;
      if (c < cols) yr[c] = v[q];
    }
  }
}

// flag[i] = 1 where a run of equal keys starts; *unsorted |= keys decrease somewhere
__global__ void __launch_bounds__(kThreads) run_flag_kernel(const int64_t* __restrict__ keys, int64_t n,
                                                             const int32_t* __restrict__ n_dev,
                                                             int32_t* __restrict__ flag, int32_t* __restrict__ unsorted) {
  const int64_t live = n_dev ? min((int64_t)*n_dev, n) : n;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i <= n; i += (int64_t)gridDim.x * kThreads) {
    int f = 0;
    if (i < live) {
      f = i == 0 || keys[i] != keys[i - 1];
      if (unsorted && i > 0 && keys[i] < keys[i - 1]) *unsorted = 1;   // every writer stores the same value
    }
    flag[i] = f;
  }
}
// after the exclusive scan: heads[flag_scan[i]] = i for run starts, *n_runs = flag_scan[n]
__global__ void __launch_bounds__(kThreads) run_scatter_kernel(const int64_t* __restrict__ keys, int64_t n,
                                                                const int32_t* __restrict__ n_dev,
                                                                const int32_t* __restrict__ scan,
                                                                int32_t* __restrict__ heads, int32_t* __restrict__ n_runs) {
  const int64_t live = n_dev ? min((int64_t)*n_dev, n) : n;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < live; i += (int64_t)gridDim.x * kThreads)
    if (i == 0 || keys[i] != keys[i - 1]) heads[scan[i]] = (int32_t)i;
  if (blockIdx.x == 0 && threadIdx.x == 0) *n_runs = scan[n];
}

// out[u] = first arg-max of row rowmap(heads[u]) (torch.argmax), u < *n_dev
__global__ void __launch_bounds__(kThreads) row_argmax_kernel(const float* __restrict__ x, int64_t ld,
                                                               const int32_t* __restrict__ heads,
                                                               const int32_t* __restrict__ rowmap,
                                                               const int32_t* __restrict__ n_dev, int64_t n_max, int cols,
                                                               int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t live = n_dev ? min((int64_t)*n_dev, n_max) : n_max;
  for (int64_t u = (int64_t)blockIdx.x * kWarps + warp; u < live; u += (int64_t)gridDim.x * kWarps) {
    int32_t r = heads ? heads[u] : (int32_t)u;
    if (rowmap) r = rowmap[r];
    const float* xr = x + (int64_t)r * ld;
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int c = lane; c < cols; c += 32) {
      const float v = xr[c];
      if (v > best || (v == best && c < arg)) { best = v; arg = c; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, d);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, d);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) out[u] = arg == 0x7fffffff ? 0 : arg;
  }
}

// Note j with onset o: i = last change point whose onset is <= o; if it is not the last change point, row j takes the
// row of that change point's first note.  A representative row only ever receives itself, so reading and writing
// the same matrix is safe.
__global__ void __launch_bounds__(kThreads) decode_assign_kernel(float* __restrict__ y, int64_t ld, int cols,
                                                                  const int64_t* __restrict__ onsets, int64_t n_rows,
                                                                  const int64_t* __restrict__ onsets_f,
                                                                  const int32_t* __restrict__ onset_heads,
                                                                  const int32_t* __restrict__ rowmap,
                                                                  const int32_t* __restrict__ cp_heads,
                                                                  const int32_t* __restrict__ n_cp_dev) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_cp = *n_cp_dev;
  if (n_cp < 2) return;
  for (int64_t j = (int64_t)blockIdx.x * kWarps + warp; j < n_rows; j += (int64_t)gridDim.x * kWarps) {
    const int64_t o = onsets[j];
    int lo = 0, hi = n_cp;                          // first change point with onset > o
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (onsets_f[onset_heads[cp_heads[mid]]] <= o) lo = mid + 1;
      else hi = mid;
    }
    const int seg = lo - 1;
    if (seg < 0 || seg >= n_cp - 1) continue;
    int32_t src = onset_heads[cp_heads[seg]];
    if (rowmap) src = rowmap[src];
    if (src == j) continue;
    const float* s = y + (int64_t)src * ld;
    float* d = y + j * ld;
    for (int c = lane; c < cols; c += 32) d[c] = s[c];
  }
}

unsigned grid_for(int64_t items, int per_block) {
  int64_t b = ceil_div(items > 0 ? items : 1, per_block);
  if (b > kNumSM * 8) b = kNumSM * 8;
  return (unsigned)b;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_softmax2_rows(const float* x, int64_t ld_x, const int32_t* rows, int64_t n_out, int cols, float* y,
                                  int64_t ld_y, agnn_stream_t stream) {
  if (n_out < 0 || cols < 1 || cols > 32 * kPerLane || ld_x < cols || ld_y < cols)
    return fail(AGNN_ERR_UNSUPPORTED, "softmax2_rows: 1 <= cols <= %d and row strides >= cols (cols = %d)", 32 * kPerLane, cols);
  if (n_out == 0) return AGNN_OK;
  if (!x || !y) return fail(AGNN_ERR_ARG, "softmax2_rows: null pointer");
  softmax2_kernel<<<grid_for(n_out, kWarps), kThreads, 0, (cudaStream_t)stream>>>(x, ld_x, rows, n_out, cols, y, ld_y);
  return check_launch("softmax2_rows");
}

extern "C" size_t agnn_run_heads_workspace(int64_t n) {
  const size_t flags = ((size_t)(n + 1) * sizeof(int32_t) + 255) / 256 * 256;
  return flags + scan_workspace_bytes(n + 1) + 256;
}

extern "C" int agnn_run_heads(const int64_t* keys, int64_t n, const int32_t* n_dev, int32_t* heads, int32_t* n_runs,
                              int32_t* unsorted, void* workspace, size_t workspace_bytes, agnn_stream_t stream) {
  if (n < 0 || n >= (1ll << 31) - 1) return fail(AGNN_ERR_ARG, "run_heads: bad size");
  if (!n_runs || !workspace || workspace_bytes < agnn_run_heads_workspace(n) || (n > 0 && (!keys || !heads)))
    return fail(AGNN_ERR_WORKSPACE, "run_heads: null pointer or workspace smaller than agnn_run_heads_workspace(n)");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* flag = static_cast<int32_t*>(workspace);
  const size_t flags = ((size_t)(n + 1) * sizeof(int32_t) + 255) / 256 * 256;
  int32_t* tile_sums = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + flags);
  run_flag_kernel<<<grid_for(n + 1, kThreads), kThreads, 0, st>>>(keys, n, n_dev, flag, unsorted);
  int rc = exclusive_scan_i32(flag, n + 1, tile_sums, st);
  if (rc) return rc;
  run_scatter_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(keys, n, n_dev, flag, heads, n_runs);
  return check_launch("run_heads");
}

extern "C" int agnn_row_argmax(const float* x, int64_t ld, const int32_t* heads, const int32_t* rowmap,
                               const int32_t* n_dev, int64_t n_max, int cols, int64_t* out, agnn_stream_t stream) {
  if (n_max < 0 || cols < 1 || ld < cols) return fail(AGNN_ERR_ARG, "row_argmax: bad sizes");
  if (n_max == 0) return AGNN_OK;
  if (!x || !out) return fail(AGNN_ERR_ARG, "row_argmax: null pointer");
  row_argmax_kernel<<<grid_for(n_max, kWarps), kThreads, 0, (cudaStream_t)stream>>>(x, ld, heads, rowmap, n_dev, n_max,
                                                                                    cols, out);
  return check_launch("row_argmax");
}

extern "C" int agnn_decode_assign(float* y, int64_t ld, int cols, const int64_t* onsets, int64_t n_rows,
                                  const int64_t* onsets_f, const int32_t* onset_heads, const int32_t* rowmap,
                                  const int32_t* cp_heads, const int32_t* n_cp, agnn_stream_t stream) {
  if (n_rows < 0 || cols < 1 || ld < cols) return fail(AGNN_ERR_ARG, "decode_assign: bad sizes");
  if (n_rows == 0) return AGNN_OK;
  if (!y || !onsets || !onsets_f || !onset_heads || !cp_heads || !n_cp) return fail(AGNN_ERR_ARG, "decode_assign: null pointer");
  decode_assign_kernel<<<grid_for(n_rows, kWarps), kThreads, 0, (cudaStream_t)stream>>>(y, ld, cols, onsets, n_rows, onsets_f,
                                                                                        onset_heads, rowmap, cp_heads, n_cp);
  return check_launch("decode_assign");
}
