// Fused HGT edge-softmax attention (forward + the two backward passes).
//
// Replaces, for one destination node type, PyG HGTConv's per-edge pipeline (third
// party; semantics fixed in SURVEY.md section 8c and oracle/pyg.py::HGTConv):
//     score_e = (q_i . k_e) * p_rel / sqrt(D)      gather q[dst], k[src]      [E,H]
//     alpha   = softmax over ALL in-edges of i     4 scatter kernels          [E,H]
//     out_i   = sum_e alpha_e v_e                  gather v[src] + scatter    [N,H,D]
// as ONE pass over the dst-sorted CSR of every relation: a warp owns a destination
// row, keeps q_i, the running max / sum and the weighted value sum in registers
// (online softmax), and never materialises a per-edge tensor.  No atomics.
//
// Backward is two more passes of the same shape:
//   bwd_dst (dst-sorted CSR):  delta_i = dO_i . out_i ; dq_i ; d p_rel partials
//   bwd_src (src-sorted CSR):  dk_j, dv_j  (recomputes alpha_e from the saved max / sum)
//
// Roofline: HBM.  Algorithmic bytes, b = bytes / element:
//   fwd      E * (2*H*D*b + 4) + N_dst * (2*H*D*b + 8*H) + rowptr
//   bwd_dst  E * (2*H*D*b + 4) + N_dst * (4*H*D*b + 12*H)
//   bwd_src  E * (2*H*D*b + 4 + 12*H) + N_src * 4*H*D*b
#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
// Edges in flight per warp (their k / v -- and q / dO -- row slices stay packed in registers until used).  Measured on
// the config-3 batch (tools/hgt_probe.py, bf16): 8 per batch instead of 4 costs the forward 30 % and the backward 15 %
// (a batch's compute is serial, and most rows have fewer than ten edges); 2 instead of 4 in bwd_src costs 9 %.
#ifndef AGNN_HGT_DST_U16
#define AGNN_HGT_DST_U16 4
#endif
#ifndef AGNN_HGT_SRC_U16
#define AGNN_HGT_SRC_U16 4
#endif
template <typename T, int V>
struct Batch {
  static constexpr bool kHalf = sizeof(T) == 2;
  static constexpr int fwd = V <= 2 ? 4 : 2;
  static constexpr int dst = V <= 2 ? (kHalf ? AGNN_HGT_DST_U16 : 4) : 2;
  static constexpr int src = V <= 2 ? (kHalf ? AGNN_HGT_SRC_U16 : 2) : 1;
};

struct HgtParams {
  int n_rows, heads, head_dim, n_rel;
  agnn_hgt_rel_t rel[AGNN_MAX_REL];
  const void* q;  // fwd / bwd_dst: this dst type's q.  bwd_src: unused (per relation in rel[].q)
  int64_t ld_q;
  const float* pscale;  // [n_rel * heads]
  void* out;            // fwd: output; bwd_dst: forward output (read)
  int64_t ld_out;
  const void* dout;
  int64_t ld_dout;
  float* row_max;  // [n_rows * heads]
  float* row_den;
  float* delta;
  void* dq;
  int64_t ld_dq;
  float* dpscale_partial;  // [gridDim.x, n_rel * heads]
};

// Softmax weights go through ex2.approx (relative error 2^-22) on scores pre-multiplied by log2(e): one MUFU per
// edge and head instead of two full-precision expf calls; the subtraction s - m is exact in either base.
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int LPH>
__device__ __forceinline__ float head_sum(float x) {
#pragma unroll
  for (int d = LPH / 2; d >= 1; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
  return x;
}

template <int E>
__device__ __forceinline__ float dot(const float (&a)[E], const float (&b)[E]) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) s = fmaf(a[e], b[e], s);
  return s;
}

// Column mapping shared by the three kernels: vector slot v of lane l covers
// columns [(v*32 + l) * E, +E); LPH = head_dim / E consecutive lanes share a head.  FULL: heads * head_dim fills
// all 32 * V slots (the usual 256- / 512-wide layers), so no slot needs a guard.
#define AGNN_HGT_COLS()                                    \
  using VT = Vec16<T>;                                     \
  constexpr int E = VT::E;                                 \
  const int lane = threadIdx.x & 31;                       \
  const int HD = p.heads * p.head_dim;                     \
  int colv[V], headv[V];                                   \
  bool on[V];                                              \
  _Pragma("unroll") for (int v = 0; v < V; ++v) {          \
    colv[v] = (v * 32 + lane) * E;                         \
    on[v] = FULL || colv[v] < HD;                          \
    headv[v] = on[v] ? colv[v] / p.head_dim : 0;           \
  }

// Index work of one destination row, lane-parallel (the relations of a typed score graph hold 1-2 edges per row each:
// walking them one after another is a chain of three dependent loads per relation).  Lane r reads relation r's row
// extent, a warp scan flattens the row's (relation, edge) items in relation order, and lane t resolves item t0 + t:
// three dependent round trips for up to 32 edges, after which the warp streams the k / v rows several at a time.
struct RowItems {
  int beg, deg, incl, total;
};

// lane r < n_rel: relation r's extent of `row`
__device__ __forceinline__ void load_extent(const HgtParams& p, int row, int lane, int& beg, int& end) {
  beg = end = 0;
  if (lane < p.n_rel) {
    const agnn_hgt_rel_t& R = p.rel[lane];
    beg = __ldg(R.rowptr + row);
    end = __ldg(R.rowptr + row + 1);
  }
}

__device__ __forceinline__ RowItems row_items(int beg, int end, int lane) {
  constexpr unsigned kFull = 0xffffffffu;
  RowItems it;
  it.beg = beg;
  it.deg = end - beg;
  it.incl = it.deg;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(kFull, it.incl, d);
    if (lane >= d) it.incl += t;
  }
  it.total = __shfl_sync(kFull, it.incl, 31);
  return it;
}

// item t of the row: its relation, and the k / v rows of its source node (null past the end)
template <typename T>
__device__ __forceinline__ void resolve_item(const HgtParams& p, const RowItems& it, int t, int& rel, const T*& kp,
                                             const T*& vp) {
  constexpr unsigned kFull = 0xffffffffu;
  // the first relation whose inclusive count exceeds t (lanes >= n_rel carry the total): 4 steps over 16 lanes
  static_assert(AGNN_MAX_REL == 16, "the search below covers 16 relations");
  int r = 0;
#pragma unroll
  for (int step = 8; step >= 1; step >>= 1)
    if (__shfl_sync(kFull, it.incl, r + step - 1) <= t) r += step;
  r = min(r, p.n_rel - 1);
  const int k = __shfl_sync(kFull, it.beg, r) + (t - (__shfl_sync(kFull, it.incl, r) - __shfl_sync(kFull, it.deg, r)));
  rel = r;
  kp = vp = nullptr;
  if (t < it.total) {
    const agnn_hgt_rel_t& R = p.rel[r];
    const int64_t off = (int64_t)__ldg(R.col + k) * R.ld_kv;
    kp = static_cast<const T*>(R.k) + off;
    vp = static_cast<const T*>(R.v) + off;
  }
}

template <typename P>
__device__ __forceinline__ const P* shfl_ptr(const P* ptr, int src_lane) {
  return reinterpret_cast<const P*>(__shfl_sync(0xffffffffu, (unsigned long long)ptr, src_lane));
}

template <typename T, int V, int LPH, bool FULL>
__global__ void __launch_bounds__(kThreads, 2) hgt_fwd_kernel(const __grid_constant__ HgtParams p) {
  AGNN_HGT_COLS();
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int U = Batch<T, V>::fwd;
  const T* const q = static_cast<const T*>(p.q);
  T* const out = static_cast<T*>(p.out);
  for (int row = blockIdx.x * kWarps + (threadIdx.x >> 5); row < p.n_rows; row += gridDim.x * kWarps) {
    float qv[V][E], acc[V][E], m[V], l[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      m[v] = -INFINITY;
      l[v] = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) { acc[v][e] = 0.f; qv[v][e] = 0.f; }
      if (on[v]) VT::load_nc(q + (int64_t)row * p.ld_q + colv[v], qv[v]);
    }
    int xbeg, xend;
    load_extent(p, row, lane, xbeg, xend);
    const RowItems it = row_items(xbeg, xend, lane);
    for (int t0 = 0; t0 < it.total; t0 += 32) {
      int my_rel;
      const T *my_k, *my_v;
      resolve_item<T>(p, it, t0 + lane, my_rel, my_k, my_v);
      const int n_items = min(32, it.total - t0);
      for (int i0 = 0; i0 < n_items; i0 += U) {
        const T *kp[U], *vp[U];
        uint4 kraw[U][V], vraw[U][V];
        float ps[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int sl = min(i0 + u, 31);
          kp[u] = shfl_ptr(my_k, sl);
          vp[u] = shfl_ptr(my_v, sl);
          const int r = __shfl_sync(kFull, my_rel, sl);
          if (i0 + u >= n_items) kp[u] = nullptr;
          if (kp[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              ps[u][v] = __ldg(p.pscale + r * p.heads + headv[v]);     // (no arithmetic on loaded values here:
                                                                         //  it would stall the loads queued behind it)
              if (on[v]) {
                kraw[u][v] = VT::load_raw_nc(kp[u] + colv[v]);
                vraw[u][v] = VT::load_raw_nc(vp[u] + colv[v]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (kp[u]) {  // warp-uniform
#pragma unroll
            for (int v = 0; v < V; ++v) {
              // online softmax in base 2: one of the two factors exp2(m - max), exp2(s - max) is exactly 1
              float kx[E], vx[E];
              VT::unpack(kraw[u][v], kx);
              VT::unpack(vraw[u][v], vx);
              const float s = head_sum<LPH>(on[v] ? dot<E>(qv[v], kx) : 0.f) * (ps[u][v] * kLog2e);
              const float d = s - m[v];
              const float ex = ex2(-fabsf(d));
              const bool up = d > 0.f;
              const float c = up ? ex : 1.f, w = up ? 1.f : ex;
              l[v] = fmaf(l[v], c, w);
#pragma unroll
              for (int e = 0; e < E; ++e) acc[v][e] = fmaf(acc[v][e], c, w * (on[v] ? vx[e] : 0.f));
              m[v] = up ? s : m[v];
            }
          }
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      if (!on[v]) continue;
      const float den = l[v] + 1e-16f;  // torch_geometric.utils.softmax: out / (sum + 1e-16)
      const float inv = 1.f / den;
      float o[E];
#pragma unroll
      for (int e = 0; e < E; ++e) o[e] = acc[v][e] * inv;
      VT::store(out + (int64_t)row * p.ld_out + colv[v], o);
      if ((lane % LPH) == 0) {
        p.row_max[(int64_t)row * p.heads + headv[v]] = m[v];           // in log2 units, as the backward wants it
        p.row_den[(int64_t)row * p.heads + headv[v]] = den;
      }
    }
  }
}

template <typename T, int V, int LPH, bool FULL>
__global__ void __launch_bounds__(kThreads, 2) hgt_bwd_dst_kernel(const __grid_constant__ HgtParams p) {
  AGNN_HGT_COLS();
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int U = Batch<T, V>::dst;
  __shared__ float part[kWarps][AGNN_MAX_REL * AGNN_HGT_MAX_HEADS];
  const int warp = threadIdx.x >> 5;
  const int slots = p.n_rel * p.heads;
  for (int i = lane; i < slots; i += 32) part[warp][i] = 0.f;
  __syncwarp();
  const T* const q = static_cast<const T*>(p.q);
  const T* const out = static_cast<const T*>(p.out);
  const T* const dout = static_cast<const T*>(p.dout);
  T* const dq = static_cast<T*>(p.dq);
  for (int row = blockIdx.x * kWarps + warp; row < p.n_rows; row += gridDim.x * kWarps) {
    float qv[V][E], gv[V][E], acc[V][E], m[V], den[V], dl[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float ov[E];
#pragma unroll
      for (int e = 0; e < E; ++e) { acc[v][e] = 0.f; qv[v][e] = 0.f; gv[v][e] = 0.f; ov[e] = 0.f; }
      if (on[v]) {
        VT::load_nc(q + (int64_t)row * p.ld_q + colv[v], qv[v]);
        VT::load_nc(dout + (int64_t)row * p.ld_dout + colv[v], gv[v]);
        VT::load_nc(out + (int64_t)row * p.ld_out + colv[v], ov);
      }
      dl[v] = head_sum<LPH>(dot<E>(gv[v], ov));
      m[v] = p.row_max[(int64_t)row * p.heads + headv[v]];
      den[v] = 1.f / p.row_den[(int64_t)row * p.heads + headv[v]];
      if (on[v] && (lane % LPH) == 0) p.delta[(int64_t)row * p.heads + headv[v]] = dl[v];
    }
    int xbeg, xend;
    load_extent(p, row, lane, xbeg, xend);
    const RowItems it = row_items(xbeg, xend, lane);
    for (int t0 = 0; t0 < it.total; t0 += 32) {
      int my_rel;
      const T *my_k, *my_v;
      resolve_item<T>(p, it, t0 + lane, my_rel, my_k, my_v);
      const int n_items = min(32, it.total - t0);
      for (int i0 = 0; i0 < n_items; i0 += U) {
        const T *kp[U], *vp[U];
        int rel[U];
        uint4 kraw[U][V], vraw[U][V];
        float ps[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int sl = min(i0 + u, 31);
          kp[u] = shfl_ptr(my_k, sl);
          vp[u] = shfl_ptr(my_v, sl);
          rel[u] = __shfl_sync(kFull, my_rel, sl);
          if (i0 + u >= n_items) kp[u] = nullptr;
          if (kp[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              ps[u][v] = __ldg(p.pscale + rel[u] * p.heads + headv[v]);
              if (on[v]) {
                kraw[u][v] = VT::load_raw_nc(kp[u] + colv[v]);
                vraw[u][v] = VT::load_raw_nc(vp[u] + colv[v]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (kp[u]) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              float kx[E], vx[E];
              VT::unpack(kraw[u][v], kx);
              VT::unpack(vraw[u][v], vx);
              const float raw = head_sum<LPH>(on[v] ? dot<E>(qv[v], kx) : 0.f);
              const float da = head_sum<LPH>(on[v] ? dot<E>(gv[v], vx) : 0.f);
              const float alpha = ex2(raw * (ps[u][v] * kLog2e) - m[v]) * den[v];
              const float ds = alpha * (da - dl[v]);
              if (on[v] && (lane % LPH) == 0) part[warp][rel[u] * p.heads + headv[v]] += ds * raw;
              const float t = ds * ps[u][v];
#pragma unroll
              for (int e = 0; e < E; ++e) acc[v][e] = fmaf(t, on[v] ? kx[e] : 0.f, acc[v][e]);
            }
          }
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
      if (on[v]) VT::store(dq + (int64_t)row * p.ld_dq + colv[v], acc[v]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < slots; i += kThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += part[w][i];
    p.dpscale_partial[(int64_t)blockIdx.x * slots + i] = s;
  }
}

// blockIdx.y = relation; rows = source nodes of that relation (transposed CSR).  A warp takes 32 consecutive source
// rows at a time: lane i reads row i's extent (coalesced), the rows' edges are one contiguous range of the transposed
// CSR, so lane t resolves edge t (destination id -> q / dO rows) and the warp streams the edges a few at a time across
// row boundaries; the per-row sums are flushed when the row changes.  Rows without edges get zeros.
template <typename T, int V, int LPH, bool FULL>
__global__ void __launch_bounds__(kThreads, 2) hgt_bwd_src_kernel(const __grid_constant__ HgtParams p) {
  AGNN_HGT_COLS();
  constexpr unsigned kFull = 0xffffffffu;
  constexpr int U = Batch<T, V>::src;             // q, dO, k, v rows per edge
  const int r = blockIdx.y;
  const agnn_hgt_rel_t& R = p.rel[r];
  const T* const q = static_cast<const T*>(p.q);
  const T* const dout = static_cast<const T*>(p.dout);
  const T* const kk = static_cast<const T*>(R.k);
  const T* const vv = static_cast<const T*>(R.v);
  T* const dk = static_cast<T*>(R.dk);
  T* const dv = static_cast<T*>(R.dv);
  float ps[V];
#pragma unroll
  for (int v = 0; v < V; ++v) ps[v] = __ldg(p.pscale + r * p.heads + headv[v]);
  for (int row0 = (blockIdx.x * kWarps + (threadIdx.x >> 5)) * 32; row0 < R.n_src; row0 += gridDim.x * kWarps * 32) {
    const int my_row = row0 + lane;
    int beg = 0, deg = 0;
    if (my_row < R.n_src) {
      beg = __ldg(R.t_rowptr + my_row);
      deg = __ldg(R.t_rowptr + my_row + 1) - beg;
    }
    int incl = deg;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(kFull, incl, 31);
    const int base = __shfl_sync(kFull, beg, 0);
    // rows without edges: zeros
    unsigned empty = __ballot_sync(kFull, deg == 0 && my_row < R.n_src);
    while (empty) {
      const int rl = __ffs(empty) - 1;
      empty &= empty - 1;
      float z[E];
#pragma unroll
      for (int e = 0; e < E; ++e) z[e] = 0.f;
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (on[v]) {
          VT::store(dk + (int64_t)(row0 + rl) * R.ld_dkv + colv[v], z);
          VT::store(dv + (int64_t)(row0 + rl) * R.ld_dkv + colv[v], z);
        }
    }
    float ak[V][E], av[V][E];
    int cur = -1;                                   // local row whose sums are in ak / av
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      int my_rl = 0, my_idx = -1;
      for (int j = 0; j < 32; ++j) my_rl += (__shfl_sync(kFull, incl, j) <= t) ? 1 : 0;
      my_rl = min(my_rl, 31);
      if (t < total) my_idx = __ldg(R.t_col + base + t);
      const int n_items = min(32, total - t0);
      for (int i0 = 0; i0 < n_items; i0 += U) {
        int idx[U], rl[U];
        uint4 qraw[U][V], graw[U][V], kraw[U][V], vraw[U][V];
        float mx[U][V], dn[U][V], dl[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int sl = min(i0 + u, 31);
          idx[u] = __shfl_sync(kFull, my_idx, sl);
          rl[u] = __shfl_sync(kFull, my_rl, sl);
          if (i0 + u >= n_items) idx[u] = -1;
          if (idx[u] >= 0) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (on[v]) {
                qraw[u][v] = VT::load_raw_nc(q + (int64_t)idx[u] * p.ld_q + colv[v]);
                graw[u][v] = VT::load_raw_nc(dout + (int64_t)idx[u] * p.ld_dout + colv[v]);
                kraw[u][v] = VT::load_raw_nc(kk + (int64_t)(row0 + rl[u]) * R.ld_kv + colv[v]);
                vraw[u][v] = VT::load_raw_nc(vv + (int64_t)(row0 + rl[u]) * R.ld_kv + colv[v]);
              }
              const int64_t o = (int64_t)idx[u] * p.heads + headv[v];
              mx[u][v] = __ldg(p.row_max + o);
              dn[u][v] = __ldg(p.row_den + o);
              dl[u][v] = __ldg(p.delta + o);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (idx[u] >= 0) {
            if (rl[u] != cur) {                      // warp-uniform: the previous row is complete
              if (cur >= 0) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                  if (on[v]) {
                    VT::store(dk + (int64_t)(row0 + cur) * R.ld_dkv + colv[v], ak[v]);
                    VT::store(dv + (int64_t)(row0 + cur) * R.ld_dkv + colv[v], av[v]);
                  }
              }
              cur = rl[u];
#pragma unroll
              for (int v = 0; v < V; ++v)
#pragma unroll
                for (int e = 0; e < E; ++e) { ak[v][e] = 0.f; av[v][e] = 0.f; }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
              float qx[E], gx[E], kx[E], vx[E];
              VT::unpack(qraw[u][v], qx);
              VT::unpack(graw[u][v], gx);
              VT::unpack(kraw[u][v], kx);
              VT::unpack(vraw[u][v], vx);
              const float raw = head_sum<LPH>(on[v] ? dot<E>(qx, kx) : 0.f);
              const float da = head_sum<LPH>(on[v] ? dot<E>(gx, vx) : 0.f);
              const float alpha = __fdividef(ex2(raw * (ps[v] * kLog2e) - mx[u][v]), dn[u][v]);
              const float tt = alpha * (da - dl[u][v]) * ps[v];
#pragma unroll
              for (int e = 0; e < E; ++e) {
                ak[v][e] = fmaf(tt, on[v] ? qx[e] : 0.f, ak[v][e]);
                av[v][e] = fmaf(alpha, on[v] ? gx[e] : 0.f, av[v][e]);
              }
            }
          }
      }
    }
    if (cur >= 0) {
#pragma unroll
      for (int v = 0; v < V; ++v)
        if (on[v]) {
          VT::store(dk + (int64_t)(row0 + cur) * R.ld_dkv + colv[v], ak[v]);
          VT::store(dv + (int64_t)(row0 + cur) * R.ld_dkv + colv[v], av[v]);
        }
    }
  }
}

enum Pass { kFwd, kBwdDst, kBwdSrc };

template <typename T, int V, int LPH, bool FULL>
int launch(Pass pass, const HgtParams& p, int max_rows, cudaStream_t st) {
  int64_t blocks = ceil_div(pass == kBwdSrc ? ceil_div(max_rows, 32) : max_rows, kWarps);
  const int64_t cap = (int64_t)kNumSM * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (pass == kFwd) {
    hgt_fwd_kernel<T, V, LPH, FULL><<<(unsigned)blocks, kThreads, 0, st>>>(p);
  } else if (pass == kBwdDst) {
    hgt_bwd_dst_kernel<T, V, LPH, FULL><<<(unsigned)blocks, kThreads, 0, st>>>(p);
  } else {
    hgt_bwd_src_kernel<T, V, LPH, FULL><<<dim3((unsigned)blocks, (unsigned)p.n_rel), kThreads, 0, st>>>(p);
  }
  return check_launch("hgt_attn");
}

template <typename T>
int dispatch(Pass pass, const HgtParams& p, int max_rows, cudaStream_t st) {
  constexpr int E = Vec16<T>::E;
  const int lph = p.head_dim / E;
  const int vecs = p.heads * p.head_dim / E;
  const int V = (vecs + 31) / 32;
  const bool full = vecs == 32 * V;
#define AGNN_HGT_CASE(VV, LL)                                                            \
  if (V == VV && lph == LL)                                                              \
    return full ? launch<T, VV, LL, true>(pass, p, max_rows, st) : launch<T, VV, LL, false>(pass, p, max_rows, st);
  AGNN_HGT_CASE(1, 1) AGNN_HGT_CASE(1, 2) AGNN_HGT_CASE(1, 4) AGNN_HGT_CASE(1, 8) AGNN_HGT_CASE(1, 16)
  AGNN_HGT_CASE(1, 32) AGNN_HGT_CASE(2, 4) AGNN_HGT_CASE(2, 8) AGNN_HGT_CASE(2, 16) AGNN_HGT_CASE(2, 32)
  AGNN_HGT_CASE(4, 8) AGNN_HGT_CASE(4, 16) AGNN_HGT_CASE(4, 32)
#undef AGNN_HGT_CASE
  return fail(AGNN_ERR_UNSUPPORTED, "hgt_attn: heads=%d head_dim=%d is not supported (head_dim*elem must be 16..512 bytes, "
              "a power of two, heads*head_dim <= %d)", p.heads, p.head_dim, 128 * E);
}

int common_checks(const char* what, int32_t n_rows, int heads, int head_dim, int dtype, int n_rel,
                  const agnn_hgt_rel_t* rels) {
  if (n_rows < 0 || heads < 1 || heads > AGNN_HGT_MAX_HEADS || head_dim < 1 || n_rel < 1 || n_rel > AGNN_MAX_REL || !rels)
    return fail(AGNN_ERR_ARG, "%s: bad sizes (n_rows=%d heads=%d head_dim=%d n_rel=%d)", what, n_rows, heads, head_dim, n_rel);
  if (dtype != AGNN_F32 && dtype != AGNN_BF16) return fail(AGNN_ERR_ARG, "%s: dtype %d", what, dtype);
  const int ev = dtype == AGNN_F32 ? 4 : 8;
  if (head_dim % ev || ((head_dim / ev) & (head_dim / ev - 1)) || head_dim / ev > 32)
    return fail(AGNN_ERR_UNSUPPORTED, "%s: head_dim %d must be %d * 2^k, k <= 5", what, head_dim, ev);
  return AGNN_OK;
}

int check_rows(const char* what, const void* ptr, int64_t ld, int eb) {
  if (!ptr || !aligned16(ptr) || (ld * eb) % 16 != 0)
    return fail(AGNN_ERR_ARG, "%s must be non-null, 16-byte aligned, with a 16-byte multiple row stride", what);
  return AGNN_OK;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_hgt_attn_fwd(int32_t n_dst, int heads, int head_dim, int dtype, int n_rel,
                                 const agnn_hgt_rel_t* rels, const void* q, int64_t ld_q, const float* pscale,
                                 void* out, int64_t ld_out, float* row_max, float* row_den, agnn_stream_t stream) {
  int rc = common_checks("hgt_attn_fwd", n_dst, heads, head_dim, dtype, n_rel, rels);
  if (rc) return rc;
  if (n_dst == 0) return AGNN_OK;
  const int eb = dtype == AGNN_F32 ? 4 : 2;
  if ((rc = check_rows("hgt_attn_fwd: q", q, ld_q, eb)) || (rc = check_rows("hgt_attn_fwd: out", out, ld_out, eb))) return rc;
  if (!pscale || !row_max || !row_den) return fail(AGNN_ERR_ARG, "hgt_attn_fwd: null pscale / row_max / row_den");
  HgtParams p{};
  p.n_rows = n_dst; p.heads = heads; p.head_dim = head_dim; p.n_rel = n_rel;
  p.q = q; p.ld_q = ld_q; p.pscale = pscale; p.out = out; p.ld_out = ld_out; p.row_max = row_max; p.row_den = row_den;
  for (int r = 0; r < n_rel; ++r) {
    p.rel[r] = rels[r];
    if (!rels[r].rowptr) return fail(AGNN_ERR_ARG, "hgt_attn_fwd: relation %d has a null rowptr", r);
    if ((rc = check_rows("hgt_attn_fwd: k", rels[r].k, rels[r].ld_kv, eb)) ||
        (rc = check_rows("hgt_attn_fwd: v", rels[r].v, rels[r].ld_kv, eb))) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == AGNN_F32 ? dispatch<float>(kFwd, p, n_dst, st) : dispatch<__nv_bfloat16>(kFwd, p, n_dst, st);
}

extern "C" int agnn_hgt_attn_bwd_dst_blocks(int32_t n_dst) {
  int64_t blocks = ceil_div(n_dst, kWarps);
  const int64_t cap = (int64_t)kNumSM * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

extern "C" int agnn_hgt_attn_bwd_dst(int32_t n_dst, int heads, int head_dim, int dtype, int n_rel,
                                     const agnn_hgt_rel_t* rels, const void* q, int64_t ld_q, const float* pscale,
                                     const void* out, int64_t ld_out, const void* dout, int64_t ld_dout,
                                     const float* row_max, const float* row_den, float* delta, void* dq,
                                     int64_t ld_dq, float* dpscale_partial, agnn_stream_t stream) {
  int rc = common_checks("hgt_attn_bwd_dst", n_dst, heads, head_dim, dtype, n_rel, rels);
  if (rc) return rc;
  if (n_dst == 0) return AGNN_OK;
  const int eb = dtype == AGNN_F32 ? 4 : 2;
  if ((rc = check_rows("hgt_attn_bwd_dst: q", q, ld_q, eb)) || (rc = check_rows("hgt_attn_bwd_dst: out", out, ld_out, eb)) ||
      (rc = check_rows("hgt_attn_bwd_dst: dout", dout, ld_dout, eb)) || (rc = check_rows("hgt_attn_bwd_dst: dq", dq, ld_dq, eb)))
    return rc;
  if (!pscale || !row_max || !row_den || !delta || !dpscale_partial)
    return fail(AGNN_ERR_ARG, "hgt_attn_bwd_dst: null statistics / partial pointer");
  HgtParams p{};
  p.n_rows = n_dst; p.heads = heads; p.head_dim = head_dim; p.n_rel = n_rel;
  p.q = q; p.ld_q = ld_q; p.pscale = pscale; p.out = const_cast<void*>(out); p.ld_out = ld_out;
  p.dout = dout; p.ld_dout = ld_dout; p.row_max = const_cast<float*>(row_max); p.row_den = const_cast<float*>(row_den);
  p.delta = delta; p.dq = dq; p.ld_dq = ld_dq; p.dpscale_partial = dpscale_partial;
  for (int r = 0; r < n_rel; ++r) {
    p.rel[r] = rels[r];
    if (!rels[r].rowptr) return fail(AGNN_ERR_ARG, "hgt_attn_bwd_dst: relation %d has a null rowptr", r);
    if ((rc = check_rows("hgt_attn_bwd_dst: k", rels[r].k, rels[r].ld_kv, eb)) ||
        (rc = check_rows("hgt_attn_bwd_dst: v", rels[r].v, rels[r].ld_kv, eb))) return rc;
  }
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == AGNN_F32 ? dispatch<float>(kBwdDst, p, n_dst, st) : dispatch<__nv_bfloat16>(kBwdDst, p, n_dst, st);
}

extern "C" int agnn_hgt_attn_bwd_src(int heads, int head_dim, int dtype, int n_rel, const agnn_hgt_rel_t* rels,
                                     const void* q, int64_t ld_q, const float* pscale, const void* dout,
                                     int64_t ld_dout, const float* row_max, const float* row_den,
                                     const float* delta, agnn_stream_t stream) {
  int rc = common_checks("hgt_attn_bwd_src", 0, heads, head_dim, dtype, n_rel, rels);
  if (rc) return rc;
  const int eb = dtype == AGNN_F32 ? 4 : 2;
  if ((rc = check_rows("hgt_attn_bwd_src: q", q, ld_q, eb)) || (rc = check_rows("hgt_attn_bwd_src: dout", dout, ld_dout, eb)))
    return rc;
  if (!pscale || !row_max || !row_den || !delta) return fail(AGNN_ERR_ARG, "hgt_attn_bwd_src: null statistics pointer");
  HgtParams p{};
  p.heads = heads; p.head_dim = head_dim; p.n_rel = n_rel;
  p.q = q; p.ld_q = ld_q; p.pscale = pscale; p.dout = dout; p.ld_dout = ld_dout;
  p.row_max = const_cast<float*>(row_max); p.row_den = const_cast<float*>(row_den); p.delta = const_cast<float*>(delta);
  int max_rows = 0;
  for (int r = 0; r < n_rel; ++r) {
    p.rel[r] = rels[r];
    if (rels[r].n_src < 0) return fail(AGNN_ERR_ARG, "hgt_attn_bwd_src: relation %d has n_src < 0", r);
    if (rels[r].n_src == 0) continue;
    if (!rels[r].t_rowptr) return fail(AGNN_ERR_ARG, "hgt_attn_bwd_src: relation %d has a null transposed rowptr", r);
    if ((rc = check_rows("hgt_attn_bwd_src: k", rels[r].k, rels[r].ld_kv, eb)) ||
        (rc = check_rows("hgt_attn_bwd_src: v", rels[r].v, rels[r].ld_kv, eb)) ||
        (rc = check_rows("hgt_attn_bwd_src: dk", rels[r].dk, rels[r].ld_dkv, eb)) ||
        (rc = check_rows("hgt_attn_bwd_src: dv", rels[r].dv, rels[r].ld_dkv, eb))) return rc;
    if (rels[r].n_src > max_rows) max_rows = rels[r].n_src;
  }
  if (max_rows == 0) return AGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == AGNN_F32 ? dispatch<float>(kBwdSrc, p, max_rows, st) : dispatch<__nv_bfloat16>(kBwdSrc, p, max_rows, st);
}
