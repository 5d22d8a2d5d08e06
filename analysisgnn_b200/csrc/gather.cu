// Relation-fused CSR gather-reduce (forward and, on the transposed CSR, backward)
// for the message-passing layers.  See include/agnn.h for the exact semantics and
// the reference lines replaced (gnn.py:70-74, :511, :539; hgnn.py:406-407;
// analysis.py:586; PyG SAGEConv mean aggregation).
//
// Mapping: a group of LANES threads owns one output row and walks the row's CSR
// segment of every relation in turn; each lane keeps V 16-byte column vectors in
// fp32 registers.  Neighbour rows are fetched with 128-bit read-only loads, four
// edges in flight per group.  No atomics: the sum over a row runs in CSR (= input
// edge) order, so results are deterministic run to run.
//
// Roofline: HBM.  Algorithmic bytes per launch:
//   sum_r E_r * (F*b + 4)  gathered rows + col ids
//   + n_rows * n_out_slices * F*b  written  (+ self / copy rows read)
//   + 4 * n_rel * (n_rows + 1)     rowptr
#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

struct GatherParams {
  int n_rows, n_feat, n_rel, scale, combine;
  int copy_col;
  agnn_rel_t rel[AGNN_MAX_REL];
  const void* self_add;
  int64_t ld_self;
  const void* copy;
  int64_t ld_copy;
  void* out;
  int64_t ld_out;
  void* out_lo;  // optional (fp32 only): out receives rna_tf32(y), out_lo rna_tf32(y - out) -- agnn_gemm operands
  const float* pair_amax;  // optional (fp32, CONCAT, with out_lo): out / out_lo are fp16 matrices receiving the F16X3
                           // operand pair of s y, s = f16_scale_of(*pair_amax); ld_out counts fp16 elements
  float* amax_out;      // optional (plain fp32 / bf16 output): *amax_out = max(*amax_out, max |out|) over what is written
  float* heavy_ws;      // optional: partial rows of the heavy-row path, [max_chunks][n_feat]
  int64_t max_chunks;
};

constexpr int kHeavyRow = AGNN_HEAVY_ROW;
constexpr int kHeavyChunk = AGNN_HEAVY_CHUNK;

// is (relation, row) handled by the heavy-row kernels instead of the row's own warp?
__device__ __forceinline__ bool is_heavy(const GatherParams& p, const agnn_rel_t& R, int deg) {
  return p.heavy_ws && R.heavy_rows && deg >= kHeavyRow;
}

// store one 16-byte vector at element `idx` of the output, optionally as the TF32 hi / lo pair the tensor-core GEMM
// consumes, or (f16s != 0) as the fp16 hi / lo pair of f16s * v
template <typename T>
__device__ __forceinline__ void store_split(T* out, T* out_lo, int64_t idx, const float (&v)[Vec16<T>::E],
                                            float f16s = 0.f) {
  if constexpr (sizeof(T) == 4) {
    if (f16s != 0.f) {
      uint2 h, l;
      f16_pair4(v, f16s, h, l);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out) + idx) = h;
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(out_lo) + idx) = l;
      return;
    }
    if (out_lo) {
      float h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        uint32_t hb, lb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v[e]));
        h[e] = __uint_as_float(hb);
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(v[e] - h[e]));
        l[e] = __uint_as_float(lb);
      }
      Vec16<T>::store(out + idx, h);
      Vec16<T>::store(out_lo + idx, l);
      return;
    }
  }
  Vec16<T>::store(out + idx, v);
}

template <int E>
__device__ __forceinline__ uint32_t absmax_bits(uint32_t m, const float (&v)[E]) {
#pragma unroll
  for (int e = 0; e < E; ++e) m = max(m, __float_as_uint(v[e]) & 0x7fffffffu);
  return m;
}
__device__ __forceinline__ void publish_amax(float* amax_out, uint32_t m) {
  if (!amax_out) return;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(reinterpret_cast<unsigned int*>(amax_out), m);
}

template <typename T, int LANES, int V>
__global__ void __launch_bounds__(kThreads, V <= 2 ? 4 : 1) gather_reduce_kernel(const __grid_constant__ GatherParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  constexpr int kRowsPerBlock = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int F = p.n_feat;
  T* const out = static_cast<T*>(p.out);
  T* const out_lo = static_cast<T*>(p.out_lo);
  const float f16s = p.pair_amax ? f16_scale_of(__ldg(p.pair_amax)) : 0.f;
  uint32_t mx = 0;

  for (int row = blockIdx.x * kRowsPerBlock + threadIdx.x / LANES; row < p.n_rows;
       row += gridDim.x * kRowsPerBlock) {
    float selfv[V][E];
    if (p.self_add) {
      const T* sp = static_cast<const T*>(p.self_add) + (int64_t)row * p.ld_self;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = (v * LANES + lane) * E;
        if (c < F) VT::load_nc(sp + c, selfv[v]);
      }
    }
    float tot[V][E];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int e = 0; e < E; ++e) tot[v][e] = 0.f;

    for (int r = 0; r < p.n_rel; ++r) {
      const agnn_rel_t& R = p.rel[r];
      const T* src = static_cast<const T*>(R.src);
      float acc[V][E];
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[v][e] = 0.f;
      float s = 1.f;

      if ((R.flags & AGNN_REL_IDENTITY_IF_EMPTY) && __ldg(R.rowptr + p.n_rows) == __ldg(R.rowptr)) {
        const T* rp = src + (int64_t)row * R.ld_src;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int c = (v * LANES + lane) * E;
          if (c < F) VT::load_nc(rp + c, acc[v]);
        }
      } else {
        const int beg = __ldg(R.rowptr + row), end = __ldg(R.rowptr + row + 1);
        if (is_heavy(p, R, end - beg)) continue;   // this slice / contribution comes from the heavy-row kernels
        for (int k = beg; k < end; k += kUnroll) {
          int idx[kUnroll];
          float w[kUnroll];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            idx[u] = (k + u < end) ? __ldg(R.col + k + u) : -1;
            w[u] = 1.f;
          }
          if (R.nbr_deg_rowptr) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
              if (idx[u] >= 0) {
                const int d = __ldg(R.nbr_deg_rowptr + idx[u] + 1) - __ldg(R.nbr_deg_rowptr + idx[u]);
                w[u] = 1.f / (float)max(d, 1);
              }
          }
          float x[kUnroll][V][E];
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
            if (idx[u] >= 0) {
              const T* rp = src + (int64_t)idx[u] * R.ld_src;
#pragma unroll
              for (int v = 0; v < V; ++v) {
                const int c = (v * LANES + lane) * E;
                if (c < F) VT::load_nc(rp + c, x[u][v]);
              }
            }
#pragma unroll
          for (int u = 0; u < kUnroll; ++u)
            if (idx[u] >= 0) {
#pragma unroll
              for (int v = 0; v < V; ++v)
#pragma unroll
                for (int e = 0; e < E; ++e) acc[v][e] = fmaf(w[u], x[u][v][e], acc[v][e]);
            }
        }
        if (p.combine == AGNN_COMBINE_CONCAT && p.self_add) {
#pragma unroll
          for (int v = 0; v < V; ++v)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[v][e] += selfv[v][e];
        }
        if (p.scale == AGNN_SCALE_MEAN) s = 1.f / (float)max(end - beg, 1);
      }

      if (p.combine == AGNN_COMBINE_CONCAT) {
        const int64_t off = (int64_t)row * p.ld_out + R.out_col;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const int c = (v * LANES + lane) * E;
          if (c < F) {
            float o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = acc[v][e] * s;
            mx = absmax_bits<E>(mx, o);
            store_split<T>(out, out_lo, off + c, o, f16s);
          }
        }
      } else {
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
          for (int e = 0; e < E; ++e) tot[v][e] = fmaf(acc[v][e], s, tot[v][e]);
      }
    }

    if (p.combine == AGNN_COMBINE_SUM) {
      const int64_t off = (int64_t)row * p.ld_out + p.rel[0].out_col;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = (v * LANES + lane) * E;
        if (c < F) {
          float o[E];
#pragma unroll
          for (int e = 0; e < E; ++e) o[e] = tot[v][e] + (p.self_add ? selfv[v][e] : 0.f);
          mx = absmax_bits<E>(mx, o);
          store_split<T>(out, out_lo, off + c, o, f16s);
        }
      }
    }
    if (p.copy) {
      const T* cp = static_cast<const T*>(p.copy) + (int64_t)row * p.ld_copy;
      const int64_t off = (int64_t)row * p.ld_out + p.copy_col;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = (v * LANES + lane) * E;
        if (c < F) {
          float t[E];
          VT::load_nc(cp + c, t);
          mx = absmax_bits<E>(mx, t);
          store_split<T>(out, out_lo, off + c, t, f16s);
        }
      }
    }
  }
  if (p.amax_out) {
    // groups narrower than a warp: every lane holds its own maximum, the warp publishes one
    publish_amax(p.amax_out, mx);
  }
}


// ---- low-degree single-relation launches -----------------------------------------------------------------
// A row with one or two entries is a chain of dependent round trips (rowptr -> col -> source row) for 1 - 2 KB: with one
// row per warp an SM has 32 KB in flight where 7 TB/s x ~1 us of latency asks for ~47 KB, and uniform degree 1 sat at
// 0.52 - 0.54 of the HBM peak.  Here a warp owns kRows CONSECUTIVE rows at a time: one lane-parallel rowptr load covers
// all of them, then round k gathers the k-th neighbour of every row that has one -- kRows source rows in flight per warp
// instead of min(degree, 4).  Taken for one relation flagged AGNN_REL_LOW_DEGREE (the caller knows edges / rows < 1.5)
// without neighbour weights, copy column or the E == 0 flag; hub rows are left to the hub-row kernels as usual.
constexpr int kRows = 4;

template <typename T, int V>
__global__ void __launch_bounds__(kThreads, 2) gather_lowdeg_kernel(const __grid_constant__ GatherParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int F = p.n_feat;
  const agnn_rel_t& R = p.rel[0];
  const T* src = static_cast<const T*>(R.src);
  T* const out = static_cast<T*>(p.out);
  T* const out_lo = static_cast<T*>(p.out_lo);
  const float f16s = p.pair_amax ? f16_scale_of(__ldg(p.pair_amax)) : 0.f;
  const bool concat = p.combine == AGNN_COMBINE_CONCAT;
  uint32_t mx = 0;
  const int warps = gridDim.x * (kThreads / 32);
  for (int64_t base = (int64_t)(blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5)) * kRows; base < p.n_rows;
       base += (int64_t)warps * kRows) {
    const int rp = __ldg(R.rowptr + min(base + lane, (int64_t)p.n_rows));     // lanes 0 .. kRows matter
    int beg[kRows], deg[kRows];
    bool skip[kRows];
    int rounds = 0;
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      beg[i] = __shfl_sync(kFull, rp, i);
      deg[i] = base + i < p.n_rows ? __shfl_sync(kFull, rp, i + 1) - beg[i] : 0;
      skip[i] = base + i >= p.n_rows || is_heavy(p, R, deg[i]);
      if (!skip[i]) rounds = max(rounds, deg[i]);
    }
    float acc[kRows][V][E];
#pragma unroll
    for (int i = 0; i < kRows; ++i)
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[i][v][e] = 0.f;
    // round k: the k-th neighbour of every row that has one -- up to kRows source rows in flight.  (A variant that
    // switched to the general kernel's row-by-row loop when one of the four rows is long cost the uniform case all it had
    // gained: 128 registers and spills; the hint is therefore only given below 1.5 entries per row.)
    for (int k = 0; k < rounds; ++k) {
      int idx[kRows];
#pragma unroll
      for (int i = 0; i < kRows; ++i) idx[i] = (!skip[i] && k < deg[i]) ? __ldg(R.col + beg[i] + k) : -1;
      float x[kRows][V][E];
#pragma unroll
      for (int i = 0; i < kRows; ++i)
        if (idx[i] >= 0) {
          const T* rpt = src + (int64_t)idx[i] * R.ld_src;
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const int c = (v * 32 + lane) * E;
            if (c < F) VT::load_nc(rpt + c, x[i][v]);
          }
        }
#pragma unroll
      for (int i = 0; i < kRows; ++i)
        if (idx[i] >= 0) {
#pragma unroll
          for (int v = 0; v < V; ++v)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[i][v][e] += x[i][v][e];
        }
    }
#pragma unroll
    for (int i = 0; i < kRows; ++i) {
      const int64_t row = base + i;
      if (row >= p.n_rows) break;
      if (skip[i] && concat) continue;             // the hub-row kernels write this slice (self term included)
      const float s = p.scale == AGNN_SCALE_MEAN ? 1.f / (float)max(deg[i], 1) : 1.f;
      const int64_t off = row * p.ld_out + R.out_col;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = (v * 32 + lane) * E;
        if (c < F) {
          float o[E], sv[E];
#pragma unroll
          for (int e = 0; e < E; ++e) sv[e] = 0.f;
          if (p.self_add) VT::load_nc(static_cast<const T*>(p.self_add) + row * p.ld_self + c, sv);
#pragma unroll
          for (int e = 0; e < E; ++e)               // CONCAT: s (self + acc); SUM: self + s acc  (include/agnn.h)
            o[e] = concat ? (acc[i][v][e] + sv[e]) * s : __fadd_rn(__fmul_rn(acc[i][v][e], s), sv[e]);
          mx = absmax_bits<E>(mx, o);
          store_split<T>(out, out_lo, off + c, o, f16s);
        }
      }
    }
  }
  if (p.amax_out) publish_amax(p.amax_out, mx);
}

// ---- COMBINE_SUM fast path ------------------------------------------------------------------------
// The backward gathers (d x_src = self + sum over all outgoing relations) see 1-2 edges in each of ~9
// relations per row: walking the relations one after another is a chain of ~4 dependent global loads per
// relation (rowptr, col, neighbour degree, row).  Here the index work is lane-parallel: lane r reads
// relation r's row extent, a warp scan flattens all (relation, edge) items of the row, lane t resolves item
// t (column id, neighbour-degree weight, source row address) -- three dependent round trips for up to 32
// items -- and then the whole warp streams the source rows, four in flight.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads, V <= 2 ? 4 : 1) gather_sum_kernel(const __grid_constant__ GatherParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int F = p.n_feat;
  T* const out = static_cast<T*>(p.out);
  T* const out_lo = static_cast<T*>(p.out_lo);
  uint32_t mx = 0;
  for (int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < p.n_rows; row += gridDim.x * (kThreads / 32)) {
    float tot[V][E];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = (v * 32 + lane) * E;
      if (p.self_add && c < F) {
        VT::load_nc(static_cast<const T*>(p.self_add) + (int64_t)row * p.ld_self + c, tot[v]);
      } else {
#pragma unroll
        for (int e = 0; e < E; ++e) tot[v][e] = 0.f;
      }
    }
    // lane r < n_rel: extent of relation r's row
    int beg = 0, deg = 0;
    if (lane < p.n_rel) {
      const agnn_rel_t& R = p.rel[lane];
      beg = __ldg(R.rowptr + row);
      deg = __ldg(R.rowptr + row + 1) - beg;
      if (is_heavy(p, R, deg)) deg = 0;              // handled by the heavy-row kernels
    }
    const float rel_scale = (p.scale == AGNN_SCALE_MEAN) ? 1.f / (float)max(deg, 1) : 1.f;
    int incl = deg;                                  // inclusive scan of the degrees over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const int excl = incl - deg;
    const int total = __shfl_sync(kFull, incl, 31);
    for (int t0 = 0; t0 < total; t0 += 32) {
      // item t = t0 + lane: which relation, which edge
      const int t = t0 + lane;
      int r = 0;
      for (int q = 0; q < p.n_rel; ++q) r += (__shfl_sync(kFull, incl, q) <= t) ? 1 : 0;
      r = min(r, p.n_rel - 1);
      const int k = __shfl_sync(kFull, beg, r) + (t - __shfl_sync(kFull, excl, r));
      const float sc = __shfl_sync(kFull, rel_scale, r);
      const T* rowp = nullptr;
      float w = 0.f;
      if (t < total) {
        const agnn_rel_t& R = p.rel[r];
        const int idx = __ldg(R.col + k);
        w = sc;
        if (R.nbr_deg_rowptr) {
          const int d = __ldg(R.nbr_deg_rowptr + idx + 1) - __ldg(R.nbr_deg_rowptr + idx);
          w = sc / (float)max(d, 1);
        }
        rowp = static_cast<const T*>(R.src) + (int64_t)idx * R.ld_src;
      }
      const int n_items = min(32, total - t0);
      for (int i0 = 0; i0 < n_items; i0 += kUnroll) {
        const T* ptr[kUnroll];
        float wi[kUnroll];
        float x[kUnroll][V][E];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int src_lane = min(i0 + u, 31);
          const unsigned long long a = __shfl_sync(kFull, (unsigned long long)rowp, src_lane);
          ptr[u] = (i0 + u < n_items) ? reinterpret_cast<const T*>(a) : nullptr;
          wi[u] = __shfl_sync(kFull, w, src_lane);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (ptr[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const int c = (v * 32 + lane) * E;
              if (c < F) VT::load_nc(ptr[u] + c, x[u][v]);
            }
          }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (ptr[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
              for (int e = 0; e < E; ++e) tot[v][e] = fmaf(wi[u], x[u][v][e], tot[v][e]);
          }
      }
    }
    const int64_t off = (int64_t)row * p.ld_out + p.rel[0].out_col;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = (v * 32 + lane) * E;
      if (c < F) {
        mx = absmax_bits<E>(mx, tot[v]);
        store_split<T>(out, out_lo, off + c, tot[v]);
      }
    }
    if (p.copy) {
      const T* cp = static_cast<const T*>(p.copy) + (int64_t)row * p.ld_copy;
      const int64_t coff = (int64_t)row * p.ld_out + p.copy_col;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int c = (v * 32 + lane) * E;
        if (c < F) {
          float tmp[E];
          VT::load_nc(cp + c, tmp);
          mx = absmax_bits<E>(mx, tmp);
          store_split<T>(out, out_lo, coff + c, tmp);
        }
      }
    }
  }
  publish_amax(p.amax_out, mx);
}

// ---- heavy rows: split across warps -------------------------------------------------------------
// Chunk numbering: relations in order, their heavy rows in list (= row) order, kHeavyChunk edges per chunk; the
// list carries the exclusive prefix of the chunk counts (agnn_csr_build), so warp g finds its chunk by a binary
// search and never walks the lists.  Partial sums go to heavy_ws and are combined per row in chunk order by the second
// kernel (deterministic, no atomics).
struct HeavyIndex {
  int nh;        // lane r: heavy rows of relation r
  int64_t base;  // lane r: first chunk of relation r in the global numbering
  int64_t incl;  // lane r: base + chunks of relation r
  int item_incl; // lane r: heavy rows of relations 0..r
};

// lane r < n_rel resolves relation r's list; every lane gets the totals through shuffles
__device__ __forceinline__ HeavyIndex heavy_index(const GatherParams& p, int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  HeavyIndex ix;
  ix.nh = 0;
  long long chunks = 0;
  if (lane < p.n_rel) {
    const agnn_rel_t& R = p.rel[lane];
    if (R.heavy_rows && R.heavy_cap > 0) {
      ix.nh = (int)min((int64_t)__ldg(R.n_heavy), R.heavy_cap);
      if (ix.nh > 0) {
        const int row = __ldg(R.heavy_rows + ix.nh - 1);
        const int deg = __ldg(R.rowptr + row + 1) - __ldg(R.rowptr + row);
        chunks = (long long)__ldg(R.heavy_rows + R.heavy_cap + ix.nh - 1) + (deg + kHeavyChunk - 1) / kHeavyChunk;
      }
    }
  }
  long long incl = chunks;
  int items = ix.nh;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const long long t = __shfl_up_sync(kFull, incl, d);
    const int ti = __shfl_up_sync(kFull, items, d);
    if (lane >= d) { incl += t; items += ti; }
  }
  ix.incl = incl;
  ix.base = incl - chunks;
  ix.item_incl = items;
  return ix;
}

// the last slot h in [0, n) with a[h] <= key (a ascending, a[0] <= key)
__device__ __forceinline__ int upper_slot(const int32_t* __restrict__ a, int n, int key) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(a + mid) <= key) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename T, int V>
__global__ void __launch_bounds__(kThreads) gather_heavy_partial_kernel(const __grid_constant__ GatherParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  constexpr unsigned kFull = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int64_t warp_id = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
  const int F = p.n_feat;
  const HeavyIndex ix = heavy_index(p, lane);
  const int64_t total = min((int64_t)__shfl_sync(kFull, ix.incl, 31), p.max_chunks);
  for (int64_t g = warp_id; g < total; g += n_warps) {
    int r = 0;
    for (int q = 0; q < p.n_rel; ++q) r += (__shfl_sync(kFull, ix.incl, q) <= g) ? 1 : 0;
    r = min(r, p.n_rel - 1);
    const agnn_rel_t& R = p.rel[r];
    const int local = (int)(g - __shfl_sync(kFull, ix.base, r));
    const int nh = __shfl_sync(kFull, ix.nh, r);
    const int h = upper_slot(R.heavy_rows + R.heavy_cap, nh, local);
    const int c = local - __ldg(R.heavy_rows + R.heavy_cap + h);
    const int row = __ldg(R.heavy_rows + h);
    const int beg = __ldg(R.rowptr + row), end = __ldg(R.rowptr + row + 1);
    const T* src = static_cast<const T*>(R.src);
    float acc[V][E];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int e = 0; e < E; ++e) acc[v][e] = 0.f;
    const int k0 = beg + c * kHeavyChunk, k1 = min(k0 + kHeavyChunk, end);
    for (int kb = k0; kb < k1; kb += 32) {
      // lane t resolves edge kb + t (column id, neighbour-degree weight), then the warp streams the rows
      int my_idx = -1;
      float my_w = 1.f;
      if (kb + lane < k1) {
        my_idx = __ldg(R.col + kb + lane);
        if (R.nbr_deg_rowptr) {
          const int d = __ldg(R.nbr_deg_rowptr + my_idx + 1) - __ldg(R.nbr_deg_rowptr + my_idx);
          my_w = 1.f / (float)max(d, 1);
        }
      }
      const int n_items = min(32, k1 - kb);
      for (int i0 = 0; i0 < n_items; i0 += kUnroll) {
        int idx[kUnroll];
        float w[kUnroll];
        float x[kUnroll][V][E];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int sl = min(i0 + u, 31);
          const int t = __shfl_sync(kFull, my_idx, sl);
          idx[u] = (i0 + u < n_items) ? t : -1;
          w[u] = __shfl_sync(kFull, my_w, sl);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (idx[u] >= 0) {
            const T* rp = src + (int64_t)idx[u] * R.ld_src;
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const int cc = (v * 32 + lane) * E;
              if (cc < F) VT::load_nc(rp + cc, x[u][v]);
            }
          }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
          if (idx[u] >= 0) {
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
              for (int e = 0; e < E; ++e) acc[v][e] = fmaf(w[u], x[u][v][e], acc[v][e]);
          }
      }
    }
    float* wp = p.heavy_ws + g * F;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int cc = (v * 32 + lane) * E;
      if (cc < F) {
#pragma unroll
        for (int e = 0; e < E; ++e) wp[cc + e] = acc[v][e];
      }
    }
  }
}

// The chunk partials of one (relation, heavy row), added up by the whole block: warp w takes chunks w, w + 8, ..
// (four loads in flight), warp 0 then adds the eight warp sums in warp order -- a fixed order, so the result is
// bit-identical from run to run, and a hub of a million edges is not one warp's serial chain.  Warp 0 holds the result.
template <int V, int E>
__device__ __forceinline__ void block_partials(const float* __restrict__ ws, int64_t g0, int chunks, int F,
                                               float (*part)[32 * V * E], float (&acc)[V][E]) {
  constexpr int kWarps = kThreads / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int e = 0; e < E; ++e) acc[v][e] = 0.f;
  for (int c = warp; c < chunks; c += 4 * kWarps) {
    float x[4][V][E];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int cu = c + u * kWarps;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int cc = (v * 32 + lane) * E;
#pragma unroll
        for (int e = 0; e < E; e += 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cu < chunks && cc < F) t = *reinterpret_cast<const float4*>(ws + (g0 + cu) * F + cc + e);
          x[u][v][e] = t.x; x[u][v][e + 1] = t.y; x[u][v][e + 2] = t.z; x[u][v][e + 3] = t.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[v][e] += x[u][v][e];
  }
  if (chunks <= 1) return;                              // (block-uniform) warp 0 read the only chunk
  __syncthreads();                                      // the previous call's readers are done with `part`
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int e = 0; e < E; ++e) part[warp][(v * 32 + lane) * E + e] = acc[v][e];
  __syncthreads();
  if (warp == 0) {
    const int n = min(chunks, kWarps);
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int e = 0; e < E; ++e) {
        float t = part[0][(v * 32 + lane) * E + e];
        for (int w = 1; w < n; ++w) t += part[w][(v * 32 + lane) * E + e];
        acc[v][e] = t;
      }
  }
}

// One block per heavy (relation, row).  CONCAT: every (relation, row) owns its output slice.  SUM: the block of the
// row's FIRST heavy relation adds the chunk partials of every heavy relation of that row in relation order to what the
// main kernel wrote (self + light relations) and stores once.  One writer per output element, fixed summation order:
// no atomics, bit-identical from run to run.
template <typename T, int V>
__global__ void __launch_bounds__(kThreads) gather_heavy_combine_kernel(const __grid_constant__ GatherParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  constexpr unsigned kFull = 0xffffffffu;
  __shared__ __align__(16) float part[kThreads / 32][32 * V * E];
  const int lane = threadIdx.x & 31;
  const bool writer = threadIdx.x < 32;
  const int F = p.n_feat;
  T* const out = static_cast<T*>(p.out);
  T* const out_lo = static_cast<T*>(p.out_lo);
  const float f16s = p.pair_amax ? f16_scale_of(__ldg(p.pair_amax)) : 0.f;
  const bool sum_mode = p.combine == AGNN_COMBINE_SUM;
  uint32_t hmx = 0;
  const HeavyIndex ix = heavy_index(p, lane);
  const int n_items = __shfl_sync(kFull, ix.item_incl, 31);
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    int r = 0;
    for (int q = 0; q < p.n_rel; ++q) r += (__shfl_sync(kFull, ix.item_incl, q) <= item) ? 1 : 0;
    r = min(r, p.n_rel - 1);
    const agnn_rel_t& R = p.rel[r];
    const int h = item - (__shfl_sync(kFull, ix.item_incl, r) - __shfl_sync(kFull, ix.nh, r));
    const int row = __ldg(R.heavy_rows + h);
    const int beg = __ldg(R.rowptr + row), end = __ldg(R.rowptr + row + 1);
    const int chunks = (end - beg + kHeavyChunk - 1) / kHeavyChunk;
    const int64_t g0 = __shfl_sync(kFull, ix.base, r) + __ldg(R.heavy_rows + R.heavy_cap + h);
    if (g0 + chunks > p.max_chunks) continue;
    if (sum_mode) {
      // the row belongs to the block of its first heavy relation
      bool first = true;
      for (int q = 0; q < r && first; ++q) {
        const agnn_rel_t& Q = p.rel[q];
        if (is_heavy(p, Q, __ldg(Q.rowptr + row + 1) - __ldg(Q.rowptr + row))) first = false;
      }
      if (!first) continue;
    }
    float acc[V][E];
    block_partials<V, E>(p.heavy_ws, g0, chunks, F, part, acc);
    const float s = p.scale == AGNN_SCALE_MEAN ? 1.f / (float)max(end - beg, 1) : 1.f;
    if (!sum_mode) {
      if (!writer) continue;
      const int64_t off = (int64_t)row * p.ld_out + R.out_col;
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int cc = (v * 32 + lane) * E;
        if (cc < F) {
          float o[E];
          if (p.self_add) {
            float sv[E];
            VT::load_nc(static_cast<const T*>(p.self_add) + (int64_t)row * p.ld_self + cc, sv);
#pragma unroll
            for (int e = 0; e < E; ++e) acc[v][e] += sv[e];
          }
#pragma unroll
          for (int e = 0; e < E; ++e) o[e] = acc[v][e] * s;
          hmx = absmax_bits<E>(hmx, o);
          store_split<T>(out, out_lo, off + cc, o, f16s);
        }
      }
      continue;
    }
    // COMBINE_SUM: what the main kernel wrote (self + light relations), then relation r, then the later heavy
    // relations of the same row in relation order
    float tot[V][E];
    const int64_t off = (int64_t)row * p.ld_out + p.rel[0].out_col;
    if (writer) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int cc = (v * 32 + lane) * E;
        if (cc < F) {
          VT::load(out + off + cc, tot[v]);
          if (out_lo) {
            float lo[E];
            VT::load(out_lo + off + cc, lo);
#pragma unroll
            for (int e = 0; e < E; ++e) tot[v][e] += lo[e];
          }
#pragma unroll
          for (int e = 0; e < E; ++e) tot[v][e] = fmaf(acc[v][e], s, tot[v][e]);
        }
      }
    }
    for (int q = r + 1; q < p.n_rel; ++q) {
      const agnn_rel_t& Q = p.rel[q];
      const int b2 = __ldg(Q.rowptr + row), e2 = __ldg(Q.rowptr + row + 1);
      if (!is_heavy(p, Q, e2 - b2)) continue;
      const int nq = __shfl_sync(kFull, ix.nh, q);
      if (nq <= 0) continue;
      const int h2 = upper_slot(Q.heavy_rows, nq, row);
      if (__ldg(Q.heavy_rows + h2) != row) continue;      // (list overflow: the main kernel cannot have skipped it)
      const int ch2 = (e2 - b2 + kHeavyChunk - 1) / kHeavyChunk;
      const int64_t g2 = __shfl_sync(kFull, ix.base, q) + __ldg(Q.heavy_rows + Q.heavy_cap + h2);
      if (g2 + ch2 > p.max_chunks) continue;
      const float s2 = p.scale == AGNN_SCALE_MEAN ? 1.f / (float)max(e2 - b2, 1) : 1.f;
      float a2[V][E];
      block_partials<V, E>(p.heavy_ws, g2, ch2, F, part, a2);
      if (writer) {
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
          for (int e = 0; e < E; ++e) tot[v][e] = fmaf(a2[v][e], s2, tot[v][e]);
      }
    }
    if (writer) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int cc = (v * 32 + lane) * E;
        if (cc < F) {
          hmx = absmax_bits<E>(hmx, tot[v]);
          store_split<T>(out, out_lo, off + cc, tot[v]);
        }
      }
    }
  }
  publish_amax(p.amax_out, hmx);
}

template <typename T, int V>
int launch_heavy(const GatherParams& p, cudaStream_t stream) {
  gather_heavy_partial_kernel<T, V><<<kNumSM * 4, kThreads, 0, stream>>>(p);
  gather_heavy_combine_kernel<T, V><<<kNumSM * 4, kThreads, 0, stream>>>(p);
  return check_launch("gather_reduce (heavy rows)");
}

template <typename T, int LANES, int V>
int launch(const GatherParams& p, cudaStream_t stream) {
  constexpr int kRowsPerBlock = kThreads / LANES;
  int64_t blocks = ceil_div(p.n_rows, kRowsPerBlock);
  const int64_t cap = (int64_t)kNumSM * 8;  // 8 resident CTAs/SM; grid-stride beyond that
  if (blocks > cap) blocks = cap;
  bool flags_set = false;
  for (int r = 0; r < p.n_rel; ++r) flags_set = flags_set || (p.rel[r].flags & AGNN_REL_IDENTITY_IF_EMPTY);
  bool low_degree = false;
  if constexpr (LANES == 32 && V * Vec16<T>::E <= 8)       // kRows x V x E accumulators + as many in flight: 128 registers
    low_degree = p.n_rel == 1 && p.rel[0].flags == AGNN_REL_LOW_DEGREE && !p.copy && !p.rel[0].nbr_deg_rowptr;
  if (low_degree) {
    if constexpr (LANES == 32 && V * Vec16<T>::E <= 8) {
      int64_t lb = ceil_div(p.n_rows, (int64_t)kRows * (kThreads / 32));
      if (lb > (int64_t)kNumSM * 2) lb = (int64_t)kNumSM * 2;                   // 2 resident CTAs per SM, grid-stride beyond
      gather_lowdeg_kernel<T, V><<<(unsigned)lb, kThreads, 0, stream>>>(p);
    }
  } else if (LANES == 32 && p.combine == AGNN_COMBINE_SUM && !flags_set && p.n_rel > 1) {
    gather_sum_kernel<T, V><<<(unsigned)blocks, kThreads, 0, stream>>>(p);      // lane-parallel index phase
  } else {
    gather_reduce_kernel<T, LANES, V><<<(unsigned)blocks, kThreads, 0, stream>>>(p);
  }
  if (p.heavy_ws) return launch_heavy<T, V>(p, stream);
  return check_launch("gather_reduce");
}

template <typename T>
int dispatch(const GatherParams& p, cudaStream_t stream) {
  const int vecs = p.n_feat / Vec16<T>::E;  // 16-byte vectors per row
  if (vecs <= 8) return launch<T, 8, 1>(p, stream);
  if (vecs <= 16) return launch<T, 16, 1>(p, stream);
  if (vecs <= 32) return launch<T, 32, 1>(p, stream);
  if (vecs <= 64) return launch<T, 32, 2>(p, stream);
  if (vecs <= 128) return launch<T, 32, 4>(p, stream);
  return fail(AGNN_ERR_UNSUPPORTED, "gather_reduce: n_feat %d too wide (max %d)", p.n_feat, 128 * Vec16<T>::E);
}

int check_matrix(const char* what, const void* ptr, int64_t ld, int elem_bytes) {
  if (!aligned16(ptr) || (ld * elem_bytes) % 16 != 0)
    return fail(AGNN_ERR_ARG, "%s must be 16-byte aligned with a 16-byte multiple row stride", what);
  return AGNN_OK;
}

// ---- self-term gradient: out = base + sum_r in[:, slice_r] / max(deg_r, 1) ----
struct RowScaleParams {
  int n_rows, n_feat, n_rel;
  agnn_rel_t rel[AGNN_MAX_REL];
  const void* in;
  int64_t ld_in;
  const void* base;
  int64_t ld_base;
  void* out;
  int64_t ld_out;
};

template <typename T>
__global__ void __launch_bounds__(kThreads) rowscale_sum_kernel(const __grid_constant__ RowScaleParams p) {
  using VT = Vec16<T>;
  constexpr int E = VT::E;
  const int vecs = p.n_feat / E;
  const int64_t total = (int64_t)p.n_rows * vecs;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int row = (int)(i / vecs), c = (int)(i % vecs) * E;
    float acc[E];
    if (p.base) {
      VT::load_nc(static_cast<const T*>(p.base) + (int64_t)row * p.ld_base + c, acc);
    } else {
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = 0.f;
    }
    for (int r = 0; r < p.n_rel; ++r) {
      const agnn_rel_t& R = p.rel[r];
      if ((R.flags & AGNN_REL_IDENTITY_IF_EMPTY) && __ldg(R.rowptr + p.n_rows) == __ldg(R.rowptr)) continue;
      const int d = __ldg(R.rowptr + row + 1) - __ldg(R.rowptr + row);
      const float s = 1.f / (float)max(d, 1);
      float t[E];
      VT::load_nc(static_cast<const T*>(p.in) + (int64_t)row * p.ld_in + R.out_col + c, t);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e] = fmaf(t[e], s, acc[e]);
    }
    VT::store(static_cast<T*>(p.out) + (int64_t)row * p.ld_out + c, acc);
  }
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_gather_reduce(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                                  const agnn_rel_t* rels, const void* self_add, int64_t ld_self, const void* copy,
                                  int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out, void* out_lo,
                                  void* heavy_workspace, size_t heavy_workspace_bytes, agnn_stream_t stream) {
  return agnn_gather_reduce_f16(n_rows, n_feat, dtype, scale, combine, n_rel, rels, self_add, ld_self, copy, ld_copy,
                                copy_col, out, ld_out, out_lo, nullptr, heavy_workspace, heavy_workspace_bytes, stream);
}

extern "C" int agnn_gather_reduce_f16(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                                      const agnn_rel_t* rels, const void* self_add, int64_t ld_self, const void* copy,
                                      int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out, void* out_lo,
                                      const float* pair_amax, void* heavy_workspace, size_t heavy_workspace_bytes,
                                      agnn_stream_t stream) {
  return agnn_gather_reduce_amax(n_rows, n_feat, dtype, scale, combine, n_rel, rels, self_add, ld_self, copy, ld_copy,
                                 copy_col, out, ld_out, out_lo, pair_amax, nullptr, heavy_workspace, heavy_workspace_bytes,
                                 stream);
}

extern "C" int agnn_gather_reduce_amax(int32_t n_rows, int32_t n_feat, int dtype, int scale, int combine, int n_rel,
                                       const agnn_rel_t* rels, const void* self_add, int64_t ld_self, const void* copy,
                                       int64_t ld_copy, int32_t copy_col, void* out, int64_t ld_out, void* out_lo,
                                       const float* pair_amax, float* amax_out, void* heavy_workspace,
                                       size_t heavy_workspace_bytes, agnn_stream_t stream) {
  if (n_rows < 0 || n_feat <= 0 || n_rel < 1 || n_rel > AGNN_MAX_REL || !rels || !out)
    return fail(AGNN_ERR_ARG, "gather_reduce: bad sizes (n_rows=%d n_feat=%d n_rel=%d)", n_rows, n_feat, n_rel);
  if (dtype != AGNN_F32 && dtype != AGNN_BF16) return fail(AGNN_ERR_ARG, "gather_reduce: dtype %d", dtype);
  if ((scale != AGNN_SCALE_NONE && scale != AGNN_SCALE_MEAN) ||
      (combine != AGNN_COMBINE_CONCAT && combine != AGNN_COMBINE_SUM))
    return fail(AGNN_ERR_ARG, "gather_reduce: bad scale/combine");
  const int eb = dtype == AGNN_F32 ? 4 : 2, ev = 16 / eb;
  if (n_feat % ev) return fail(AGNN_ERR_UNSUPPORTED, "gather_reduce: n_feat %d is not a multiple of %d", n_feat, ev);
  if (n_rows == 0) return AGNN_OK;
  GatherParams p;
  p.n_rows = n_rows; p.n_feat = n_feat; p.n_rel = n_rel; p.scale = scale; p.combine = combine;
  p.copy_col = copy_col;
  p.self_add = self_add; p.ld_self = ld_self; p.copy = copy; p.ld_copy = ld_copy; p.out = out; p.ld_out = ld_out;
  p.out_lo = out_lo;
  p.pair_amax = pair_amax;
  p.amax_out = amax_out;
  if (amax_out && (pair_amax || out_lo))
    return fail(AGNN_ERR_ARG, "gather_reduce: amax_out reports the plain output (not available with an operand pair)");
  if (pair_amax && (dtype != AGNN_F32 || combine != AGNN_COMBINE_CONCAT || !out_lo || (ld_out * 2) % 16))
    return fail(AGNN_ERR_ARG, "gather_reduce: the fp16 hi/lo output needs fp32 inputs, the concatenated layout, out_lo "
                              "and a row stride that is a multiple of 8 fp16 elements");
  p.heavy_ws = nullptr;
  p.max_chunks = 0;
  bool any_heavy = false;
  for (int r = 0; r < n_rel; ++r) any_heavy = any_heavy || (rels[r].heavy_rows && rels[r].n_heavy && rels[r].heavy_cap > 0);
  if (any_heavy && heavy_workspace && heavy_workspace_bytes >= (size_t)n_feat * 4) {
    if (!aligned16(heavy_workspace)) return fail(AGNN_ERR_ARG, "gather_reduce: heavy workspace must be 16-byte aligned");
    p.heavy_ws = static_cast<float*>(heavy_workspace);
    p.max_chunks = (int64_t)(heavy_workspace_bytes / ((size_t)n_feat * 4));
  }
  if (out_lo && (dtype != AGNN_F32 || !aligned16(out_lo)))
    return fail(AGNN_ERR_ARG, "gather_reduce: the TF32 hi/lo output needs fp32 and a 16-byte aligned out_lo");
  int rc;
  if ((rc = check_matrix("gather_reduce: out", out, ld_out, pair_amax ? 2 : eb))) return rc;
  if (self_add && (rc = check_matrix("gather_reduce: self_add", self_add, ld_self, eb))) return rc;
  if (copy && ((rc = check_matrix("gather_reduce: copy", copy, ld_copy, eb)) || copy_col % ev))
    return rc ? rc : fail(AGNN_ERR_ARG, "gather_reduce: copy_col must be a multiple of %d", ev);
  for (int r = 0; r < n_rel; ++r) {
    p.rel[r] = rels[r];
    if (!rels[r].src || !rels[r].rowptr) return fail(AGNN_ERR_ARG, "gather_reduce: relation %d has null pointers", r);
    if ((rc = check_matrix("gather_reduce: src", rels[r].src, rels[r].ld_src, eb))) return rc;
    if (rels[r].out_col % ev) return fail(AGNN_ERR_ARG, "gather_reduce: out_col must be a multiple of %d", ev);
  }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == AGNN_F32 ? dispatch<float>(p, st) : dispatch<__nv_bfloat16>(p, st);
}

extern "C" void agnn_heavy_params(int32_t* heavy_row, int32_t* heavy_chunk) {
  if (heavy_row) *heavy_row = kHeavyRow;
  if (heavy_chunk) *heavy_chunk = kHeavyChunk;
}

extern "C" size_t agnn_gather_heavy_workspace(int64_t total_edges, int64_t total_heavy_cap, int32_t n_feat) {
  if (total_edges < 0 || total_heavy_cap < 0 || n_feat <= 0) return 0;
  return (size_t)(total_edges / kHeavyChunk + total_heavy_cap + 1) * (size_t)n_feat * sizeof(float);
}

extern "C" int agnn_rowscale_sum(int32_t n_rows, int32_t n_feat, int dtype, int n_rel, const agnn_rel_t* rels,
                                 const void* in, int64_t ld_in, const void* base, int64_t ld_base, void* out,
                                 int64_t ld_out, agnn_stream_t stream) {
  if (n_rows < 0 || n_feat <= 0 || n_rel < 1 || n_rel > AGNN_MAX_REL || !rels || !in || !out)
    return fail(AGNN_ERR_ARG, "rowscale_sum: bad arguments");
  if (dtype != AGNN_F32 && dtype != AGNN_BF16) return fail(AGNN_ERR_ARG, "rowscale_sum: dtype %d", dtype);
  const int eb = dtype == AGNN_F32 ? 4 : 2, ev = 16 / eb;
  if (n_feat % ev) return fail(AGNN_ERR_UNSUPPORTED, "rowscale_sum: n_feat %d is not a multiple of %d", n_feat, ev);
  if (n_rows == 0) return AGNN_OK;
  int rc;
  if ((rc = check_matrix("rowscale_sum: in", in, ld_in, eb)) || (rc = check_matrix("rowscale_sum: out", out, ld_out, eb)))
    return rc;
  if (base && (rc = check_matrix("rowscale_sum: base", base, ld_base, eb))) return rc;
  RowScaleParams p;
  p.n_rows = n_rows; p.n_feat = n_feat; p.n_rel = n_rel;
  p.in = in; p.ld_in = ld_in; p.base = base; p.ld_base = ld_base; p.out = out; p.ld_out = ld_out;
  for (int r = 0; r < n_rel; ++r) {
    p.rel[r] = rels[r];
    if (!rels[r].rowptr)
      return fail(AGNN_ERR_ARG, "rowscale_sum: relation %d has a null rowptr", r);
    if (rels[r].out_col % ev) return fail(AGNN_ERR_ARG, "rowscale_sum: column offsets must be multiples of %d", ev);
  }
  const int64_t total = (int64_t)n_rows * (n_feat / ev);
  int64_t blocks = ceil_div(total, kThreads);
  if (blocks > (int64_t)kNumSM * 8) blocks = (int64_t)kNumSM * 8;
  if (dtype == AGNN_F32)
    rowscale_sum_kernel<float><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(p);
  else
    rowscale_sum_kernel<__nv_bfloat16><<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(p);
  return check_launch("rowscale_sum");
}
