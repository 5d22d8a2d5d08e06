// The objective at the end of the hot path and the gradient of the two small lookup tables at its start.
//
//   agnn_softmax_ce_fwd / _bwd   CrossEntropyLoss(ignore_index, label_smoothing), mean over the rows that are
//                                not ignored -- one per task head, summed by MultiTaskLoss
//                                (analysisgnn/models/analysis.py:881-908, 1035-1037; label_smoothing 0.1 at :893)
//   agnn_embedding_bwd           gradient of nn.Embedding(35, 64) / nn.Embedding(15, 64) (pitch spelling and key
//                                signature, analysisgnn/models/analysis.py:427-428, 572-574): 50 000 rows fall on
//                                a few dozen table rows, so the sort-based ATen kernel is replaced by per-block
//                                tables in shared memory, summed in a fixed order (no atomics).
//
// Roofline: HBM.  CE fwd reads the logits once (online softmax), bwd reads them once and writes the gradient;
// the embedding gradient reads dY once.
#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

int ce_blocks(int64_t rows) {
  int64_t b = ceil_div(rows, kWarps * 4);
  if (b > kNumSM * 4) b = kNumSM * 4;
  return b < 1 ? 1 : (int)b;
}

// loss_i = (1 - eps) * (lse - x_y) + eps * (lse - mean_c x_c)   for rows whose label is not ignore_index
__global__ void __launch_bounds__(kThreads) ce_fwd_kernel(const float* __restrict__ x, int64_t ld,
                                                           const int64_t* __restrict__ labels, int64_t rows, int cols,
                                                           float smoothing, int64_t ignore_index,
                                                           float* __restrict__ lse_out, float* __restrict__ partials) {
  __shared__ float red[2][kWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float loss_acc = 0.f, cnt_acc = 0.f;             // identical in every lane of the warp
  for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < rows; r += (int64_t)gridDim.x * kWarps) {
    const float* xr = x + r * ld;
    float m = -INFINITY, s = 0.f, sum = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float v = __ldg(xr + c);
      sum += v;
      if (v > m) { s = s * expf(m - v) + 1.f; m = v; }
      else s += expf(v - m);
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, m, d), s2 = __shfl_xor_sync(0xffffffffu, s, d);
      const float mm = fmaxf(m, m2);
      s = (m == -INFINITY ? 0.f : s * expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * expf(m2 - mm));
      m = mm;
    }
    sum = warp_sum(sum);
    const float lse = m + logf(s);
    if (lane == 0) lse_out[r] = lse;
    const int64_t y = labels[r];
    if (y != ignore_index && y >= 0 && y < cols) {
      const float xy = __ldg(xr + y);
      loss_acc += (1.f - smoothing) * (lse - xy) + smoothing * (lse - sum / (float)cols);
      cnt_acc += 1.f;
    }
  }
  if (lane == 0) { red[0][warp] = loss_acc; red[1][warp] = cnt_acc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { a += red[0][w]; b += red[1][w]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

// out[0] = sum loss / valid rows, out[1] = valid rows (fixed summation order)
__global__ void __launch_bounds__(kThreads) ce_finish_kernel(const float* __restrict__ partials, int blocks,
                                                              float* __restrict__ out) {
  __shared__ float red[2][kThreads];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < blocks; i += kThreads) { a += partials[2 * i]; b += partials[2 * i + 1]; }
  red[0][threadIdx.x] = a; red[1][threadIdx.x] = b;
  __syncthreads();
  for (int d = kThreads / 2; d >= 1; d >>= 1) {
    if (threadIdx.x < d) { red[0][threadIdx.x] += red[0][threadIdx.x + d]; red[1][threadIdx.x] += red[1][threadIdx.x + d]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = red[0][0] / red[1][0]; out[1] = red[1][0]; }
}

// dx_ic = g / valid * (softmax_ic - (1 - eps) [c == y_i] - eps / C), zero rows where the label is ignored
__global__ void __launch_bounds__(kThreads) ce_bwd_kernel(const float* __restrict__ x, int64_t ld,
                                                           const int64_t* __restrict__ labels,
                                                           const float* __restrict__ lse, int64_t rows, int cols,
                                                           float smoothing, int64_t ignore_index,
                                                           const float* __restrict__ out, const float* __restrict__ gout,
                                                           float* __restrict__ dx, int64_t ld_d, int cols_pad,
                                                           float* __restrict__ amax_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float coef = __ldg(gout) / __ldg(out + 1);
  const float uniform = smoothing / (float)cols;
  // |softmax - target| <= 1, so |coef| bounds every entry: the fp16 operand scale of dx for the head's backward GEMMs
  if (amax_out && blockIdx.x == 0 && threadIdx.x == 0)
    atomicMax(reinterpret_cast<unsigned int*>(amax_out), __float_as_uint(coef) & 0x7fffffffu);
  for (int64_t r = (int64_t)blockIdx.x * kWarps + warp; r < rows; r += (int64_t)gridDim.x * kWarps) {
    const int64_t y = labels[r];
    const bool valid = y != ignore_index && y >= 0 && y < cols;
    const float* xr = x + r * ld;
    float* dr = dx + r * ld_d;
    const float l = lse[r];
    for (int c = lane; c < cols; c += 32) {
      float v = 0.f;
      if (valid) v = coef * (expf(__ldg(xr + c) - l) - (c == y ? 1.f - smoothing : 0.f) - uniform);
      dr[c] = v;
    }
    for (int c = cols + lane; c < cols_pad; c += 32) dr[c] = 0.f;     // padding columns of a 16-byte aligned row
  }
}

// ---------------------------------------------------------------- embedding gradient
// Block = 4 row groups x 64 column lanes; every (group, column) thread walks its rows in order and adds into its own
// column of the group's shared-memory table, the 4 tables are summed in group order and written as the block's partial.
constexpr int kEmbGroups = 4;
constexpr int kEmbTableMax = 3072;                 // n_emb * dim floats per group table (48 KB / 4 groups)

__global__ void __launch_bounds__(kThreads) embedding_bwd_kernel(const float* __restrict__ g, int64_t ld_g,
                                                                  const int64_t* __restrict__ idx, int64_t rows, int dim,
                                                                  int n_emb, int64_t rows_per_block,
                                                                  float* __restrict__ partials) {
  extern __shared__ float table[];                 // [kEmbGroups][n_emb * dim]
  const int tsize = n_emb * dim;
  for (int i = threadIdx.x; i < kEmbGroups * tsize; i += kThreads) table[i] = 0.f;
  __syncthreads();
  const int lanes = kThreads / kEmbGroups;         // 64 column lanes per group
  const int grp = threadIdx.x / lanes, col0 = threadIdx.x % lanes;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  const int64_t span = ceil_div(r1 - r0 > 0 ? r1 - r0 : 0, kEmbGroups);
  const int64_t a = r0 + grp * span, b = a + span < r1 ? a + span : r1;
  float* mine = table + grp * tsize;
  for (int64_t r = a; r < b; ++r) {
    const int64_t e = idx[r];
    if (e < 0 || e >= n_emb) continue;
    for (int c = col0; c < dim; c += lanes) mine[e * dim + c] += __ldg(g + r * ld_g + c);
  }
  __syncthreads();
  float* out = partials + (int64_t)blockIdx.x * tsize;
  for (int i = threadIdx.x; i < tsize; i += kThreads)
    out[i] = (table[i] + table[tsize + i]) + (table[2 * tsize + i] + table[3 * tsize + i]);
}

// dweight[i] = sum_b partials[b][i] in block order
__global__ void __launch_bounds__(kThreads) embedding_finish_kernel(const float* __restrict__ partials, int blocks,
                                                                     int tsize, float* __restrict__ dweight) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= tsize) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int b = 0;
  for (; b + 3 < blocks; b += 4) {
    a0 += partials[(int64_t)b * tsize + i];
    a1 += partials[(int64_t)(b + 1) * tsize + i];
    a2 += partials[(int64_t)(b + 2) * tsize + i];
    a3 += partials[(int64_t)(b + 3) * tsize + i];
  }
  for (; b < blocks; ++b) a0 += partials[(int64_t)b * tsize + i];
  dweight[i] = (a0 + a1) + (a2 + a3);
}

int emb_blocks(int64_t rows) {
  int64_t b = ceil_div(rows, 256);                 // >= 64 rows per row group
  if (b > kNumSM * 2) b = kNumSM * 2;
  return b < 1 ? 1 : (int)b;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_ce_blocks(int64_t rows) { return ce_blocks(rows); }

extern "C" int agnn_softmax_ce_fwd(const float* logits, int64_t ld, const int64_t* labels, int64_t rows, int cols,
                                   float smoothing, int64_t ignore_index, float* lse, float* partials, float* out,
                                   agnn_stream_t stream) {
  if (rows < 0 || cols < 1 || ld < cols) return fail(AGNN_ERR_ARG, "softmax_ce_fwd: bad sizes");
  if (!out || !partials || (rows > 0 && (!logits || !labels || !lse))) return fail(AGNN_ERR_ARG, "softmax_ce_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = ce_blocks(rows);
  ce_fwd_kernel<<<blocks, kThreads, 0, st>>>(logits, ld, labels, rows, cols, smoothing, ignore_index, lse, partials);
  ce_finish_kernel<<<1, kThreads, 0, st>>>(partials, blocks, out);
  return check_launch("softmax_ce_fwd");
}

extern "C" int agnn_softmax_ce_bwd(const float* logits, int64_t ld, const int64_t* labels, const float* lse,
                                   int64_t rows, int cols, float smoothing, int64_t ignore_index, const float* out,
                                   const float* grad_out, float* dlogits, int64_t ld_d, agnn_stream_t stream) {
  return agnn_softmax_ce_bwd_padded(logits, ld, labels, lse, rows, cols, smoothing, ignore_index, out, grad_out, dlogits,
                                    ld_d, cols, nullptr, stream);
}

extern "C" int agnn_softmax_ce_bwd_padded(const float* logits, int64_t ld, const int64_t* labels, const float* lse,
                                          int64_t rows, int cols, float smoothing, int64_t ignore_index, const float* out,
                                          const float* grad_out, float* dlogits, int64_t ld_d, int cols_padded,
                                          float* amax_out, agnn_stream_t stream) {
  if (rows < 0 || cols < 1 || ld < cols || cols_padded < cols || ld_d < cols_padded)
    return fail(AGNN_ERR_ARG, "softmax_ce_bwd: bad sizes");
  if (rows == 0) return AGNN_OK;
  if (!logits || !labels || !lse || !out || !grad_out || !dlogits) return fail(AGNN_ERR_ARG, "softmax_ce_bwd: null pointer");
  int64_t blocks = ceil_div(rows, kWarps);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  ce_bwd_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(logits, ld, labels, lse, rows, cols, smoothing,
                                                                        ignore_index, out, grad_out, dlogits, ld_d,
                                                                        cols_padded, amax_out);
  return check_launch("softmax_ce_bwd");
}

extern "C" int agnn_embedding_bwd_blocks(int64_t rows) { return emb_blocks(rows); }

extern "C" int agnn_embedding_bwd(const float* g, int64_t ld_g, const int64_t* idx, int64_t rows, int dim, int n_emb,
                                  float* partials, float* dweight, agnn_stream_t stream) {
  if (rows < 0 || dim < 1 || n_emb < 1 || ld_g < dim) return fail(AGNN_ERR_ARG, "embedding_bwd: bad sizes");
  if ((int64_t)dim * n_emb > kEmbTableMax)
    return fail(AGNN_ERR_UNSUPPORTED, "embedding_bwd: table of %d x %d exceeds %d floats (small tables only)", n_emb, dim,
                kEmbTableMax);
  if (!partials || !dweight || (rows > 0 && (!g || !idx))) return fail(AGNN_ERR_ARG, "embedding_bwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = emb_blocks(rows);
  const int tsize = dim * n_emb;
  const int64_t rows_per_block = ceil_div(rows > 0 ? rows : 1, blocks);
  embedding_bwd_kernel<<<blocks, kThreads, kEmbGroups * tsize * sizeof(float), st>>>(g, ld_g, idx, rows, dim, n_emb,
                                                                                  rows_per_block, partials);
  embedding_finish_kernel<<<(unsigned)ceil_div(tsize, kThreads), kThreads, 0, st>>>(partials, blocks, tsize, dweight);
  return check_launch("embedding_bwd");
}
