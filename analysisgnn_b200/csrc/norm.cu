// Row-wise normalisation kernels of the hot path and the column sums of its backward.
//
//   agnn_layernorm_fwd / _bwd   nn.LayerNorm inside project_dict / project_enc / the sequence branch
//                               (analysisgnn/models/analysis.py:429-443, 474-485; models/cadence.py:249-260)
//   agnn_l2norm_relu_fwd / _bwd F.normalize(p=2) + ReLU between the MetricalGNN layers
//                               (analysisgnn/models/core/hgnn.py:415, 421-422, 431)
//   agnn_colsum_partials        bias gradients (column sums of dY)
//
// One warp per row (rows are 64..1024 wide), 128-bit loads, fp32 statistics, two-pass variance as in
// ATen.  Column reductions (d gamma, d beta, bias gradients) are accumulated per lane over a
// grid-stride loop, combined across the warps of a block through shared memory and written as
// per-block partials that the caller sums -- fixed order, no atomics.
// Roofline: HBM.  fwd 2 rows moved per row (x in, y out); bwd 3 (dy, x in; dx out).
#include <initializer_list>

#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxV = 8;  // float4 vectors per lane: up to 1024 columns

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

template <int V>
__device__ __forceinline__ void load_row(const float* p, int cols, int lane, float (&v)[V][4]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + c));
      v[i][0] = t.x; v[i][1] = t.y; v[i][2] = t.z; v[i][3] = t.w;
    } else {
      v[i][0] = v[i][1] = v[i][2] = v[i][3] = 0.f;
    }
  }
}

template <int V>
__device__ __forceinline__ void store_row(float* p, int cols, int lane, const float (&v)[V][4]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (i * 32 + lane) * 4;
    if (c < cols) *reinterpret_cast<float4*>(p + c) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) layernorm_fwd_kernel(const float* __restrict__ x, int64_t ld_x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta, float* __restrict__ y,
                                                                  int64_t ld_y, float* __restrict__ mean,
                                                                  float* __restrict__ rstd, int64_t rows, int cols,
                                                                  float eps, __half* __restrict__ pair_hi,
                                                                  __half* __restrict__ pair_lo, int64_t ld_pair,
                                                                  float* __restrict__ pair_amax, float drop_p,
                                                                  const uint64_t* __restrict__ rng, uint32_t rng_stream) {
  const int lane = threadIdx.x & 31;
  float g[V][4], b[V][4];
  load_row<V>(gamma, cols, lane, g);
  load_row<V>(beta, cols, lane, b);
  const float inv = 1.f / (float)cols;
  // fp16 operand pair of the (dropped-out) output for the projection behind the norm.  Its scale must be known before
  // the first element is written, so it comes from a BOUND instead of a measured amax: a row with zero mean and unit
  // (biased) variance has |xhat| <= sqrt(cols - 1), hence |y| <= max|gamma| sqrt(cols - 1) + max|beta|, times 1/(1-p).
  float f16s = 0.f;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thr = dropout_threshold(drop_p);
  const uint64_t key = (drop_p > 0.f && rng) ? dropout_key(rng, rng_stream) : 0ull;
  if (pair_hi) {
    float mg = 0.f, mb = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        mg = fmaxf(mg, fabsf(g[i][e]));
        mb = fmaxf(mb, fabsf(b[i][e]));
      }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, d));
      mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, d));
    }
    const float bound = (mg * sqrtf((float)(cols - 1)) + mb) * keep_scale;
    f16s = f16_scale_of(bound);
    if (blockIdx.x == 0 && threadIdx.x == 0) *pair_amax = bound;
  }
  for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarps) {
    float v[V][4];
    load_row<V>(x + row * ld_x, cols, lane, v);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) s += (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
    const float mu = warp_sum(s) * inv;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = ((i * 32 + lane) * 4 + e < cols) ? v[i][e] - mu : 0.f;
        q = fmaf(d, d, q);
      }
    const float rs = rsqrtf(warp_sum(q) * inv + eps);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] = fmaf((v[i][e] - mu) * rs, g[i][e], b[i][e]);
    if (drop_p > 0.f) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
          const uint64_t bits = dropout_bits(key, ((uint64_t)row * cols + c) >> 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[i][e] = dropout_keep(bits, e, thr) ? v[i][e] * keep_scale : 0.f;
        }
      }
    }
    if (y) store_row<V>(y + row * ld_y, cols, lane, v);
    if (pair_hi) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
          uint2 h, l;
          f16_pair4(v[i], f16s, h, l);
          *reinterpret_cast<uint2*>(pair_hi + row * ld_pair + c) = h;
          *reinterpret_cast<uint2*>(pair_lo + row * ld_pair + c) = l;
        }
      }
    }
    if (lane == 0) {
      mean[row] = mu;
      rstd[row] = rs;
    }
  }
}

// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat)); partials: d gamma = sum dy*xhat, d beta = sum dy
template <int V>
__global__ void __launch_bounds__(kThreads) layernorm_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                                                  const float* __restrict__ x, int64_t ld_x,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd, float* __restrict__ dx,
                                                                  int64_t ld_dx, float* __restrict__ dgamma_part,
                                                                  float* __restrict__ dbeta_part, int64_t rows,
                                                                  int cols, float drop_p,
                                                                  const uint64_t* __restrict__ rng, uint32_t rng_stream,
                                                                  float* __restrict__ amax_out) {
  __shared__ float red[kWarps][32 * 4 + 4];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the column reduction behind this kernel may line up
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thr = dropout_threshold(drop_p);
  const uint64_t key = (drop_p > 0.f && rng) ? dropout_key(rng, rng_stream) : 0ull;
  uint32_t mx = 0;
  float g[V][4], ag[V][4], ab[V][4];
  load_row<V>(gamma, cols, lane, g);
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) ag[i][e] = ab[i][e] = 0.f;
  const float inv = 1.f / (float)cols;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < rows; row += (int64_t)gridDim.x * kWarps) {
    float d[V][4], v[V][4];
    load_row<V>(dy + row * ld_dy, cols, lane, d);
    load_row<V>(x + row * ld_x, cols, lane, v);
    if (drop_p > 0.f) {                                  // the forward's mask, recomputed: d(y') / d y = keep / (1 - p)
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
          const uint64_t bits = dropout_bits(key, ((uint64_t)row * cols + c) >> 2);
#pragma unroll
          for (int e = 0; e < 4; ++e) d[i][e] = dropout_keep(bits, e, thr) ? d[i][e] * keep_scale : 0.f;
        }
      }
    }
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool on = (i * 32 + lane) * 4 + e < cols;
        const float xh = on ? (v[i][e] - mu) * rs : 0.f;
        const float dg = d[i][e] * g[i][e];
        s1 += dg;
        s2 = fmaf(dg, xh, s2);
        ag[i][e] = fmaf(d[i][e], xh, ag[i][e]);
        ab[i][e] += d[i][e];
        v[i][e] = xh;
        d[i][e] = dg;
      }
    s1 = warp_sum(s1) * inv;
    s2 = warp_sum(s2) * inv;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        d[i][e] = rs * (d[i][e] - s1 - v[i][e] * s2);
        mx = max(mx, __float_as_uint(d[i][e]) & 0x7fffffffu);
      }
    store_row<V>(dx + row * ld_dx, cols, lane, d);
  }
  if (amax_out) {
#pragma unroll
    for (int dd = 16; dd >= 1; dd >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, dd));
    if (lane == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(amax_out), mx);
  }
  // combine the warps of the block, one 128-column slab at a time
#pragma unroll
  for (int i = 0; i < V; ++i) {
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      __syncthreads();
#pragma unroll
      for (int e = 0; e < 4; ++e) red[warp][lane * 4 + e] = pass == 0 ? ag[i][e] : ab[i][e];
      __syncthreads();
      if (threadIdx.x < 128) {
        const int c = i * 128 + threadIdx.x;
        if (c < cols) {
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
          (pass == 0 ? dgamma_part : dbeta_part)[(int64_t)blockIdx.x * cols + c] = t;
        }
      }
    }
  }
}

// y = relu?(x) / max(||relu?(x)||, eps) (relu_first)   or   relu(x / max(||x||, eps))
template <int V>
__global__ void __launch_bounds__(kThreads) l2norm_relu_fwd_kernel(const float* __restrict__ x, int64_t ld_x,
                                                                    float* __restrict__ y, int64_t ld_y,
                                                                    float* __restrict__ inv_norm, int64_t rows, int cols,
                                                                    int relu_first, float eps) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarps) {
    float v[V][4];
    load_row<V>(x + row * ld_x, cols, lane, v);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (relu_first) v[i][e] = fmaxf(v[i][e], 0.f);
        q = fmaf(v[i][e], v[i][e], q);
      }
    const float inv = 1.f / fmaxf(sqrtf(warp_sum(q)), eps);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) v[i][e] = relu_first ? v[i][e] * inv : fmaxf(v[i][e] * inv, 0.f);
    store_row<V>(y + row * ld_y, cols, lane, v);
    if (lane == 0) inv_norm[row] = inv;
  }
}

// with u = relu?(x), n = u * inv:  d u = inv * (g' - n * (g' . n)),  g' = dy masked by the outer relu;
// relu_first additionally masks d u by x > 0.  y (= saved output) provides n and both masks.
template <int V>
__global__ void __launch_bounds__(kThreads) l2norm_relu_bwd_kernel(const float* __restrict__ dy, int64_t ld_dy,
                                                                    const float* __restrict__ x, int64_t ld_x,
                                                                    const float* __restrict__ inv_norm,
                                                                    float* __restrict__ dx, int64_t ld_dx, int64_t rows,
                                                                    int cols, int relu_first) {
  const int lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); row < rows; row += (int64_t)gridDim.x * kWarps) {
    float d[V][4], v[V][4];
    load_row<V>(dy + row * ld_dy, cols, lane, d);
    load_row<V>(x + row * ld_x, cols, lane, v);
    const float inv = __ldg(inv_norm + row);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (relu_first) v[i][e] = fmaxf(v[i][e], 0.f);
        const float n = v[i][e] * inv;                     // normalised value before the outer relu
        if (!relu_first && n <= 0.f) d[i][e] = 0.f;        // relu(normalize(x)) backward mask
        dot = fmaf(d[i][e], n, dot);
        v[i][e] = n;
      }
    dot = warp_sum(dot);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float t = inv * (d[i][e] - v[i][e] * dot);
        if (relu_first && v[i][e] <= 0.f) t = 0.f;         // normalize(relu(x)): x <= 0 gets no gradient
        d[i][e] = t;
      }
    store_row<V>(dx + row * ld_dx, cols, lane, d);
  }
}

template <int V>
__global__ void __launch_bounds__(kThreads) colsum_kernel(const float* __restrict__ x, int64_t ld_x,
                                                           float* __restrict__ part, int64_t rows, int cols) {
  __shared__ float red[kWarps][32 * 4 + 4];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the column reduction behind this kernel may line up
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < rows; row += (int64_t)gridDim.x * kWarps) {
    float v[V][4];
    load_row<V>(x + row * ld_x, cols, lane, v);
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][e] += v[i][e];
  }
#pragma unroll
  for (int i = 0; i < V; ++i) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[warp][lane * 4 + e] = acc[i][e];
    __syncthreads();
    if (threadIdx.x < 128) {
      const int c = i * 128 + threadIdx.x;
      if (c < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
        part[(int64_t)blockIdx.x * cols + c] = t;
      }
    }
  }
}

// Everything the backward of a projection needs from its incoming gradient in one pass over it:
// g' = g * [relu_out > 0] (optional), the TF32 hi / lo operand pair of g' for the grad-input and grad-weight
// GEMMs, and the per-block column sums of g' (bias gradient).  Replaces where() + colsum + split_tf32:
// 7 rows of traffic per row become 4.
template <int V>
__global__ void __launch_bounds__(kThreads) grad_prepare_kernel(const float* __restrict__ g, int64_t ld_g,
                                                                 const float* __restrict__ relu_out, int64_t ld_o,
                                                                 float* __restrict__ hi, float* __restrict__ lo,
                                                                 int64_t ld_s, float* __restrict__ part, int64_t rows,
                                                                 int cols, const float* __restrict__ amax) {
  __shared__ float red[kWarps][32 * 4 + 4];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the column reduction behind this kernel may line up
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // amax given: hi / lo are fp16 matrices (ld_s in fp16 elements) receiving the F16X3 operand pair of s g'
  const float f16s = amax ? f16_scale_of(__ldg(amax)) : 0.f;
  float acc[V][4];
#pragma unroll
  for (int i = 0; i < V; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[i][e] = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * kWarps + warp; row < rows; row += (int64_t)gridDim.x * kWarps) {
    float v[V][4], h[V][4], l[V][4];
    load_row<V>(g + row * ld_g, cols, lane, v);
    if (relu_out) {
      float o[V][4];
      load_row<V>(relu_out + row * ld_o, cols, lane, o);
#pragma unroll
      for (int i = 0; i < V; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) v[i][e] = o[i][e] > 0.f ? v[i][e] : 0.f;
    }
    if (f16s != 0.f) {
      __half* hh = reinterpret_cast<__half*>(hi) + row * ld_s;
      __half* ll = reinterpret_cast<__half*>(lo) + row * ld_s;
#pragma unroll
      for (int i = 0; i < V; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[i][e] += v[i][e];
        const int c = (i * 32 + lane) * 4;
        if (c < cols) {
          uint2 a, b;
          f16_pair4(v[i], f16s, a, b);
          *reinterpret_cast<uint2*>(hh + c) = a;
          *reinterpret_cast<uint2*>(ll + c) = b;
        }
      }
      continue;
    }
#pragma unroll
    for (int i = 0; i < V; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[i][e] += v[i][e];
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[i][e]));
        h[i][e] = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v[i][e] - h[i][e]));
        l[i][e] = __uint_as_float(lb);
      }
    store_row<V>(hi + row * ld_s, cols, lane, h);
    store_row<V>(lo + row * ld_s, cols, lane, l);
  }
  if (!part) return;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    __syncthreads();
#pragma unroll
    for (int e = 0; e < 4; ++e) red[warp][lane * 4 + e] = acc[i][e];
    __syncthreads();
    if (threadIdx.x < 128) {
      const int c = i * 128 + threadIdx.x;
      if (c < cols) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) t += red[w][threadIdx.x];
        part[(int64_t)blockIdx.x * cols + c] = t;
      }
    }
  }
}

// out[c] = sum_b part[b][c] (fixed order => deterministic).  One CTA per 32 columns: lane = column (128-byte
// coalesced rows), the 8 warps stride over the partial rows with 4 loads in flight each, then combine in
// shared memory -- a single-thread-per-column loop over ~600 rows is a 100 us latency chain.
__global__ void __launch_bounds__(kThreads) reduce_partials_kernel(const float* __restrict__ part, int blocks, int cols,
                                                                    float* __restrict__ out, int64_t set_stride_part,
                                                                    int64_t set_stride_out) {
  __shared__ float red[kWarps][32];
  // launched as a programmatic dependent of the kernel that writes `part` (launch_reduce_partials): this grid is
  // already resident when the producer ends, and starts reading once its writes are visible
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const float* p = part + blockIdx.y * set_stride_part;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < cols) {
    int b = warp;
    for (; b + 3 * kWarps < blocks; b += 4 * kWarps) {
      a0 += p[(int64_t)b * cols + c];
      a1 += p[(int64_t)(b + kWarps) * cols + c];
      a2 += p[(int64_t)(b + 2 * kWarps) * cols + c];
      a3 += p[(int64_t)(b + 3 * kWarps) * cols + c];
    }
    for (; b < blocks; b += kWarps) a0 += p[(int64_t)b * cols + c];
  }
  red[warp][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (warp == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += red[w][lane];
    out[blockIdx.y * set_stride_out + c] = t;
  }
}

// The second stage of a column reduction is 8 - 16 CTAs and 4 - 9 us, 44 times per training step, most of it the gap
// between two dependent launches: the producers announce their dependents at once (griddepcontrol.launch_dependents)
// and this launch carries the programmatic-serialization attribute, so the gap overlaps the producer's run.
void launch_reduce_partials(dim3 grid, const float* part, int blocks, int cols, float* out, int64_t set_stride_part,
                            int64_t set_stride_out, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, reduce_partials_kernel, part, blocks, cols, out, set_stride_part, set_stride_out);
}

__global__ void dropout_advance_kernel(uint64_t* state) { state[1] += 1; }

// fused-layer weight of one destination type: wcat[n, :] = [sum_r Wr_r[n] | Wl_1[n] | .. | Wl_k[n]] * scale,
// bias[n] = sum_r b_r[n] * scale (sums in relation order)
struct SageWeightParams {
  int k, n, f4;
  float scale;
  const float* wr[AGNN_MAX_REL];
  const float* wl[AGNN_MAX_REL];
  const float* bl[AGNN_MAX_REL];
  float* wcat;
  float* bias;
};

__global__ void __launch_bounds__(kThreads) sage_weights_kernel(const __grid_constant__ SageWeightParams p) {
  const int row4 = (p.k + 1) * p.f4;
  const int64_t total = (int64_t)p.n * row4;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total + p.n; i += (int64_t)gridDim.x * kThreads) {
    if (i >= total) {                               // the bias tail
      const int r = (int)(i - total);
      float b = 0.f;
      for (int j = 0; j < p.k; ++j) b += __ldg(p.bl[j] + r);
      p.bias[r] = b * p.scale;
      continue;
    }
    const int r = (int)(i / row4), c4 = (int)(i - (int64_t)r * row4);
    const int blk = c4 / p.f4, c = (c4 - blk * p.f4) * 4;
    float4 v;
    if (blk == 0) {
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < p.k; ++j) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.wr[j] + (int64_t)r * p.f4 * 4 + c));
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
      }
    } else {
      v = __ldg(reinterpret_cast<const float4*>(p.wl[blk - 1] + (int64_t)r * p.f4 * 4 + c));
    }
    *reinterpret_cast<float4*>(p.wcat + (int64_t)r * row4 * 4 + (int64_t)c4 * 4) =
        make_float4(v.x * p.scale, v.y * p.scale, v.z * p.scale, v.w * p.scale);
  }
}

// y = keep ? x / (1 - p) : 0, mask of (rng, stream, row * cols + col); cols % 4 == 0
__global__ void __launch_bounds__(kThreads) dropout_apply_kernel(const float* __restrict__ x, int64_t ld_x,
                                                                  float* __restrict__ y, int64_t ld_y, int64_t rows,
                                                                  int cols4, float drop_p,
                                                                  const uint64_t* __restrict__ rng, uint32_t rng_stream,
                                                                  float* __restrict__ amax_out) {
  const float keep_scale = 1.f / (1.f - drop_p);
  const uint32_t thr = dropout_threshold(drop_p);
  const uint64_t key = dropout_key(rng, rng_stream);
  const int64_t total = rows * cols4;
  uint32_t mx = 0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    const float4 t = __ldg(reinterpret_cast<const float4*>(x + r * ld_x + c));
    float v[4] = {t.x, t.y, t.z, t.w};
    const uint64_t bits = dropout_bits(key, (uint64_t)i);          // (r * cols + c) / 4 == i
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[e] = dropout_keep(bits, e, thr) ? v[e] * keep_scale : 0.f;
      mx = max(mx, __float_as_uint(v[e]) & 0x7fffffffu);
    }
    *reinterpret_cast<float4*>(y + r * ld_y + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
  if (amax_out) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(reinterpret_cast<unsigned int*>(amax_out), mx);
  }
}

int row_blocks(int64_t rows) {
  int64_t b = ceil_div(rows, kWarps * 4);   // >= 4 rows per warp so the column partials amortise
  if (b > kNumSM * 2) b = kNumSM * 2;
  return b < 1 ? 1 : (int)b;
}

int check(const char* what, int64_t rows, int cols, std::initializer_list<const void*> ptrs,
          std::initializer_list<int64_t> lds) {
  if (rows < 0 || cols < 4 || (cols % 4) || cols > kMaxV * 128)
    return fail(AGNN_ERR_UNSUPPORTED, "%s: columns must be a multiple of 4 in [4, %d] (got %d)", what, kMaxV * 128, cols);
  for (const void* p : ptrs)
    if (!p || !aligned16(p)) return fail(AGNN_ERR_ARG, "%s: null or unaligned pointer", what);
  for (int64_t ld : lds)
    if (ld % 4) return fail(AGNN_ERR_ARG, "%s: row strides must be multiples of 4 elements", what);
  return AGNN_OK;
}

#define AGNN_DISPATCH_V(cols, CALL)                 \
  do {                                              \
    const int v_ = (int)ceil_div((cols), 128);      \
    if (v_ <= 1) { CALL(1); }                       \
    else if (v_ <= 2) { CALL(2); }                  \
    else if (v_ <= 4) { CALL(4); }                  \
    else { CALL(8); }                               \
  } while (0)

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_row_blocks(int64_t rows) { return row_blocks(rows); }

extern "C" int agnn_layernorm_fwd(const float* x, int64_t ld_x, const float* gamma, const float* beta, float* y,
                                  int64_t ld_y, float* mean, float* rstd, int64_t rows, int cols, float eps,
                                  agnn_stream_t stream) {
  if (!y) return fail(AGNN_ERR_ARG, "layernorm_fwd: null output");
  return agnn_layernorm_fwd_pair(x, ld_x, gamma, beta, y, ld_y, mean, rstd, rows, cols, eps, nullptr, nullptr, 0, nullptr,
                                 0.f, nullptr, 0, stream);
}

extern "C" int agnn_layernorm_fwd_pair(const float* x, int64_t ld_x, const float* gamma, const float* beta, float* y,
                                       int64_t ld_y, float* mean, float* rstd, int64_t rows, int cols, float eps,
                                       void* pair_hi, void* pair_lo, int64_t ld_pair, float* pair_amax, float dropout_p,
                                       const uint64_t* rng_state, uint32_t rng_stream, agnn_stream_t stream) {
  int rc = y ? check("layernorm_fwd", rows, cols, {x, gamma, beta, y}, {ld_x, ld_y})
             : check("layernorm_fwd", rows, cols, {x, gamma, beta}, {ld_x});
  if (rc) return rc;
  if (!mean || !rstd) return fail(AGNN_ERR_ARG, "layernorm_fwd: null statistics pointer");
  if (!y && !pair_hi) return fail(AGNN_ERR_ARG, "layernorm_fwd: no output requested");
  if (pair_hi && (!pair_lo || !pair_amax || ld_pair % 8 || cols % 8 || !aligned16(pair_hi) || !aligned16(pair_lo)))
    return fail(AGNN_ERR_ARG, "layernorm_fwd: the fp16 pair needs hi, lo, the amax scalar, 16-byte aligned rows and a "
                              "column count that is a multiple of 8");
  if (dropout_p < 0.f || dropout_p >= 1.f || (dropout_p > 0.f && !rng_state))
    return fail(AGNN_ERR_ARG, "layernorm_fwd: dropout needs 0 <= p < 1 and the device RNG state");
  if (rows == 0) return AGNN_OK;
  int64_t blocks = ceil_div(rows, kWarps);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) layernorm_fwd_kernel<V><<<(unsigned)blocks, kThreads, 0, st>>>(x, ld_x, gamma, beta, y, ld_y, mean, rstd, rows, cols, eps, static_cast<__half*>(pair_hi), static_cast<__half*>(pair_lo), ld_pair, pair_amax, dropout_p, rng_state, rng_stream)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  return check_launch("layernorm_fwd");
}

extern "C" int agnn_layernorm_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* gamma,
                                  const float* mean, const float* rstd, float* dx, int64_t ld_dx, float* dgamma_part,
                                  float* dbeta_part, float* dgamma, float* dbeta, int64_t rows, int cols,
                                  agnn_stream_t stream) {
  return agnn_layernorm_bwd_dropout(dy, ld_dy, x, ld_x, gamma, mean, rstd, dx, ld_dx, dgamma_part, dbeta_part, dgamma,
                                    dbeta, rows, cols, 0.f, nullptr, 0, nullptr, stream);
}

extern "C" int agnn_layernorm_bwd_dropout(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x,
                                          const float* gamma, const float* mean, const float* rstd, float* dx,
                                          int64_t ld_dx, float* dgamma_part, float* dbeta_part, float* dgamma,
                                          float* dbeta, int64_t rows, int cols, float dropout_p,
                                          const uint64_t* rng_state, uint32_t rng_stream, float* amax_out,
                                          agnn_stream_t stream) {
  int rc = check("layernorm_bwd", rows, cols, {dy, x, gamma, dx}, {ld_dy, ld_x, ld_dx});
  if (rc) return rc;
  if (!mean || !rstd || !dgamma_part || !dbeta_part) return fail(AGNN_ERR_ARG, "layernorm_bwd: null pointer");
  if (dropout_p < 0.f || dropout_p >= 1.f || (dropout_p > 0.f && !rng_state))
    return fail(AGNN_ERR_ARG, "layernorm_bwd: dropout needs 0 <= p < 1 and the device RNG state");
  const int blocks = row_blocks(rows);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) layernorm_bwd_kernel<V><<<blocks, kThreads, 0, st>>>(dy, ld_dy, x, ld_x, gamma, mean, rstd, dx, ld_dx, dgamma_part, dbeta_part, rows, cols, dropout_p, rng_state, rng_stream, amax_out)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  if (dgamma && dbeta) {
    if (dbeta_part != dgamma_part + (int64_t)blocks * cols || dbeta != dgamma + cols)
      return fail(AGNN_ERR_ARG, "layernorm_bwd: the fused final reduction needs dbeta directly behind dgamma "
                                "(partials [2][blocks][cols], outputs [2][cols])");
    launch_reduce_partials(dim3((unsigned)ceil_div(cols, 32), 2), dgamma_part, blocks, cols, dgamma,
                           (int64_t)blocks * cols, cols, st);
  }
  return check_launch("layernorm_bwd");
}

extern "C" int agnn_sage_weights(int k, int32_t n, int32_t f, const float* const* lin_r, const float* const* lin_l,
                                 const float* const* bias_l, float scale, float* wcat, float* bias, agnn_stream_t stream) {
  if (k < 1 || k > AGNN_MAX_REL || n < 1 || f < 4 || f % 4 || !lin_r || !lin_l || !bias_l || !wcat || !bias)
    return fail(AGNN_ERR_ARG, "sage_weights: 1..%d relations, feature count a multiple of 4", AGNN_MAX_REL);
  SageWeightParams p;
  p.k = k; p.n = n; p.f4 = f / 4; p.scale = scale; p.wcat = wcat; p.bias = bias;
  for (int j = 0; j < k; ++j) {
    if (!lin_r[j] || !lin_l[j] || !bias_l[j] || !aligned16(lin_r[j]) || !aligned16(lin_l[j]))
      return fail(AGNN_ERR_ARG, "sage_weights: null or unaligned parameter %d", j);
    p.wr[j] = lin_r[j]; p.wl[j] = lin_l[j]; p.bl[j] = bias_l[j];
  }
  if (!aligned16(wcat)) return fail(AGNN_ERR_ARG, "sage_weights: unaligned output");
  int64_t blocks = ceil_div((int64_t)n * (k + 1) * (f / 4) + n, kThreads);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  sage_weights_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("sage_weights");
}

extern "C" int agnn_dropout_advance(uint64_t* rng_state, agnn_stream_t stream) {
  if (!rng_state) return fail(AGNN_ERR_ARG, "dropout_advance: null state");
  dropout_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng_state);
  return check_launch("dropout_advance");
}

extern "C" int agnn_dropout_apply(const float* x, int64_t ld_x, float* y, int64_t ld_y, int64_t rows, int cols,
                                  float dropout_p, const uint64_t* rng_state, uint32_t rng_stream, float* amax_out,
                                  agnn_stream_t stream) {
  if (rows < 0 || cols < 4 || cols % 4 || ld_x % 4 || ld_y % 4 || !x || !y || !aligned16(x) || !aligned16(y))
    return fail(AGNN_ERR_ARG, "dropout_apply: needs 16-byte aligned rows and a column count multiple of 4");
  if (dropout_p <= 0.f || dropout_p >= 1.f || !rng_state)
    return fail(AGNN_ERR_ARG, "dropout_apply: needs 0 < p < 1 and the device RNG state");
  if (rows == 0) return AGNN_OK;
  int64_t blocks = ceil_div(rows * (cols / 4), kThreads);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  dropout_apply_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(x, ld_x, y, ld_y, rows, cols / 4, dropout_p,
                                                                                rng_state, rng_stream, amax_out);
  return check_launch("dropout_apply");
}

extern "C" int agnn_l2norm_relu_fwd(const float* x, int64_t ld_x, float* y, int64_t ld_y, float* inv_norm, int64_t rows,
                                    int cols, int relu_first, float eps, agnn_stream_t stream) {
  int rc = check("l2norm_relu_fwd", rows, cols, {x, y}, {ld_x, ld_y});
  if (rc) return rc;
  if (!inv_norm) return fail(AGNN_ERR_ARG, "l2norm_relu_fwd: null inv_norm");
  if (rows == 0) return AGNN_OK;
  int64_t blocks = ceil_div(rows, kWarps);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) l2norm_relu_fwd_kernel<V><<<(unsigned)blocks, kThreads, 0, st>>>(x, ld_x, y, ld_y, inv_norm, rows, cols, relu_first, eps)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  return check_launch("l2norm_relu_fwd");
}

extern "C" int agnn_l2norm_relu_bwd(const float* dy, int64_t ld_dy, const float* x, int64_t ld_x, const float* inv_norm,
                                    float* dx, int64_t ld_dx, int64_t rows, int cols, int relu_first,
                                    agnn_stream_t stream) {
  int rc = check("l2norm_relu_bwd", rows, cols, {dy, x, dx}, {ld_dy, ld_x, ld_dx});
  if (rc) return rc;
  if (!inv_norm) return fail(AGNN_ERR_ARG, "l2norm_relu_bwd: null inv_norm");
  if (rows == 0) return AGNN_OK;
  int64_t blocks = ceil_div(rows, kWarps);
  if (blocks > kNumSM * 8) blocks = kNumSM * 8;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) l2norm_relu_bwd_kernel<V><<<(unsigned)blocks, kThreads, 0, st>>>(dy, ld_dy, x, ld_x, inv_norm, dx, ld_dx, rows, cols, relu_first)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  return check_launch("l2norm_relu_bwd");
}

extern "C" int agnn_grad_prepare(const float* g, int64_t ld_g, const float* relu_out, int64_t ld_o, float* hi, float* lo,
                                 int64_t ld_s, float* partials, float* colsum, int64_t rows, int cols,
                                 agnn_stream_t stream) {
  return agnn_grad_prepare_f16(g, ld_g, relu_out, ld_o, nullptr, hi, lo, ld_s, partials, colsum, rows, cols, stream);
}

extern "C" int agnn_grad_prepare_f16(const float* g, int64_t ld_g, const float* relu_out, int64_t ld_o, const float* amax,
                                     void* hi_v, void* lo_v, int64_t ld_s, float* partials, float* colsum, int64_t rows,
                                     int cols, agnn_stream_t stream) {
  float* hi = static_cast<float*>(hi_v);
  float* lo = static_cast<float*>(lo_v);
  // fp16 outputs: a row stride of ld_s halves must be a 16-byte multiple, i.e. ld_s / 2 floats a multiple of 4
  if (amax && ld_s % 8) return fail(AGNN_ERR_ARG, "grad_prepare: fp16 row stride must be a multiple of 8 elements");
  int rc = check("grad_prepare", rows, cols, {g, hi, lo}, {ld_g, amax ? ld_s / 2 : ld_s, relu_out ? ld_o : 0});
  if (rc) return rc;
  if (relu_out && !aligned16(relu_out)) return fail(AGNN_ERR_ARG, "grad_prepare: unaligned relu_out");
  if (colsum && !partials) return fail(AGNN_ERR_ARG, "grad_prepare: colsum needs the partials buffer");
  if (rows == 0) {
    if (colsum) return fail(AGNN_ERR_ARG, "grad_prepare: no rows to sum");
    return AGNN_OK;
  }
  const int blocks = row_blocks(rows);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) grad_prepare_kernel<V><<<blocks, kThreads, 0, st>>>(g, ld_g, relu_out, ld_o, hi, lo, ld_s, partials, rows, cols, amax)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  if (colsum) launch_reduce_partials(dim3((unsigned)ceil_div(cols, 32), 1), partials, blocks, cols, colsum, 0, 0, st);
  return check_launch("grad_prepare");
}

extern "C" int agnn_colsum_partials(const float* x, int64_t ld_x, float* partials, float* out, int64_t rows, int cols,
                                    agnn_stream_t stream) {
  int rc = check("colsum_partials", rows, cols, {x, partials}, {ld_x});
  if (rc) return rc;
  const int blocks = row_blocks(rows);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(V) colsum_kernel<V><<<blocks, kThreads, 0, st>>>(x, ld_x, partials, rows, cols)
  AGNN_DISPATCH_V(cols, CALL);
#undef CALL
  if (out) launch_reduce_partials(dim3((unsigned)ceil_div(cols, 32), 1), partials, blocks, cols, out, 0, 0, st);
  return check_launch("colsum_partials");
}
