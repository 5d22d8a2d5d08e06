// Subgraph sampling on the GPU: contiguous-window node-induced subgraphs with collation, and k-hop
// uniform neighbour sampling without replacement.
//
// Replaces the CPU data path behind the reference's loaders (analysisgnn/data/datamodules/analysis.py:
// 270-323 -> graphmuse MuseNeighborLoader -> PyG NeighborSampler -> pyg-lib hetero_neighbor_sample, all
// third party) and its in-tree statement of the window step (analysisgnn/data/datasets/chord.py:217-229;
// analysisgnn/utils/hgraph.py:404-452).  Semantics are those of oracle/graph.py (window_subgraph,
// neighbor_sample): results are bit-identical for the same seed, because every random draw is a pure
// function rng_u64(seed, hop, relation, destination, draw) (splitmix64 counter RNG) and every order-
// dependent step of the sequential algorithm (edge order, node discovery order) is reproduced with
// stable scans and an atomicMin over candidate positions.
#include <climits>
#include <cstring>

#include "scan.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxFanout = 32;

__device__ __forceinline__ int upper_bound64(const int64_t* a, int lo, int hi, int64_t v) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------ window subgraphs
struct WindowParams {
  int n_slots;
  int64_t n_cand;
  const int64_t* cand_ptr;   // [B+1] candidate edges before slot b
  const int64_t* edge_lo;    // [B] first corpus edge of the slot's score
  const int64_t* node_lo;    // [B] first global node id of the window (score offset + start)
  const int32_t* win_size;   // [B] nodes in the window
  const int64_t* out_off;    // [B] batch-local id of the window's first node
  const int64_t* src; const int64_t* dst; const int64_t* type;   // corpus COO, global node ids
  int32_t* flag;             // [n_cand + 1] scanned in place
  int64_t* edges;            // [3][capacity]
  int64_t* edge_id;          // [capacity] corpus edge position
  int64_t capacity;
  int32_t* n_out;
};

__device__ __forceinline__ bool window_edge(const WindowParams& p, int64_t t, int& b, int64_t& e) {
  // n_cand may be an UPPER BOUND of the candidate count (a static-shape loader sizes its launch without reading the
  // device): positions behind the real total, cand_ptr[n_slots], flag nothing
  if (t >= __ldg(p.cand_ptr + p.n_slots)) return false;
  b = upper_bound64(p.cand_ptr, 0, p.n_slots + 1, t) - 1;
  e = __ldg(p.edge_lo + b) + (t - __ldg(p.cand_ptr + b));
  const int64_t lo = __ldg(p.node_lo + b), hi = lo + __ldg(p.win_size + b);
  const int64_t s = __ldg(p.src + e), d = __ldg(p.dst + e);
  return s >= lo && s < hi && d >= lo && d < hi;
}

__global__ void __launch_bounds__(kThreads) window_flag_kernel(const __grid_constant__ WindowParams p) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t > p.n_cand) return;
  int b; int64_t e;
  p.flag[t] = (t < p.n_cand && window_edge(p, t, b, e)) ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads) window_emit_kernel(const __grid_constant__ WindowParams p) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) *p.n_out = p.flag[p.n_cand];
  if (t >= p.n_cand) return;
  int b; int64_t e;
  if (!window_edge(p, t, b, e)) return;
  const int64_t at = p.flag[t];
  if (at >= p.capacity) return;
  const int64_t shift = __ldg(p.out_off + b) - __ldg(p.node_lo + b);
  p.edges[at] = __ldg(p.src + e) + shift;
  p.edges[p.capacity + at] = __ldg(p.dst + e) + shift;
  p.edges[2 * p.capacity + at] = p.type ? __ldg(p.type + e) : 0;
  if (p.edge_id) p.edge_id[at] = e;
}

// ------------------------------------------------------------------ counter RNG (oracle/graph.py)
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}
__device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  constexpr uint64_t G = 0x9E3779B97F4A7C15ull;
  uint64_t h = mix64(seed + G);
  h = mix64(h ^ (a + G));
  h = mix64(h ^ (b + G));
  h = mix64(h ^ (c + G));
  h = mix64(h ^ (d + G));
  return h;
}

// ------------------------------------------------------------------ neighbour sampling, one hop
struct HopParams {
  int n_rel, n_nodes, hop, fanout;
  uint64_t seed;
  const int32_t* rowptr;     // [R][n_nodes + 1], absolute positions in col
  const int32_t* col;
  const int32_t* nodes;      // discovered nodes so far (global ids)
  int frontier_lo, frontier_n, n_known;
  int32_t* counts;           // [R * F + 1] scanned in place
  int32_t* cand_slot; int32_t* cand_src; int32_t* cand_dst;   // [T]
  int32_t* flag;             // [T + 1]
  int32_t* local;            // [n_nodes] batch-local id or -1
  int32_t* first_pos;        // [n_nodes]
  int32_t* nodes_out;        // == nodes (writable)
  int32_t* src_local;        // [T]
  int n_cand;
};

__global__ void __launch_bounds__(kThreads) hop_count_kernel(const __grid_constant__ HopParams p) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  const int total = p.n_rel * p.frontier_n;
  if (t > total) return;
  int c = 0;
  if (t < total) {
    const int r = t / p.frontier_n, f = t - r * p.frontier_n;
    const int g = __ldg(p.nodes + p.frontier_lo + f);
    const int32_t* rp = p.rowptr + (int64_t)r * (p.n_nodes + 1);
    const int deg = __ldg(rp + g + 1) - __ldg(rp + g);
    c = (p.fanout < 0 || deg <= p.fanout) ? deg : p.fanout;
  }
  p.counts[t] = c;
}

__global__ void __launch_bounds__(kThreads) hop_sample_kernel(const __grid_constant__ HopParams p) {
  const int t = blockIdx.x * kThreads + threadIdx.x;
  if (t >= p.n_rel * p.frontier_n) return;
  const int r = t / p.frontier_n, f = t - r * p.frontier_n;
  const int li = p.frontier_lo + f;
  const int g = __ldg(p.nodes + li);
  const int32_t* rp = p.rowptr + (int64_t)r * (p.n_nodes + 1);
  const int beg = __ldg(rp + g), deg = __ldg(rp + g + 1) - beg;
  int at = p.counts[t];
  auto emit = [&](int slot) {
    const int s = __ldg(p.col + slot);
    p.cand_slot[at] = slot;
    p.cand_src[at] = s;
    p.cand_dst[at] = li;
    if (p.local[s] < 0) atomicMin(p.first_pos + s, at);
    ++at;
  };
  if (p.fanout < 0 || deg <= p.fanout) {
    for (int k = 0; k < deg; ++k) emit(beg + k);
    return;
  }
  // partial Fisher-Yates over a virtual identity array; swaps kept in a tiny list (oracle sample_row)
  int sw_idx[kMaxFanout * 2], sw_val[kMaxFanout * 2], n_sw = 0;
  auto get = [&](int i) {
    for (int k = 0; k < n_sw; ++k)
      if (sw_idx[k] == i) return sw_val[k];
    return i;
  };
  auto put = [&](int i, int v) {
    for (int k = 0; k < n_sw; ++k)
      if (sw_idx[k] == i) { sw_val[k] = v; return; }
    sw_idx[n_sw] = i; sw_val[n_sw] = v; ++n_sw;
  };
  for (int k = 0; k < p.fanout; ++k) {
    const int j = k + (int)(rng_u64(p.seed, (uint64_t)p.hop, (uint64_t)r, (uint64_t)g, (uint64_t)k) % (uint64_t)(deg - k));
    const int vk = get(k), vj = get(j);
    put(j, vk);
    put(k, vj);
    emit(beg + vj);
  }
}

__global__ void __launch_bounds__(kThreads) hop_flag_kernel(const __grid_constant__ HopParams p) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c > p.n_cand) return;
  int fl = 0;
  if (c < p.n_cand) {
    const int s = p.cand_src[c];
    fl = (p.local[s] < 0 && p.first_pos[s] == c) ? 1 : 0;
  }
  p.flag[c] = fl;
}

__global__ void __launch_bounds__(kThreads) hop_assign_kernel(const __grid_constant__ HopParams p) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= p.n_cand) return;
  const int s = p.cand_src[c];
  if (p.first_pos[s] == c && p.flag[c + 1] == p.flag[c] + 1) {      // this candidate discovered s
    const int id = p.n_known + p.flag[c];
    p.local[s] = id;
    p.nodes_out[id] = s;
  }
}

__global__ void __launch_bounds__(kThreads) hop_relabel_kernel(const __grid_constant__ HopParams p) {
  const int c = blockIdx.x * kThreads + threadIdx.x;
  if (c >= p.n_cand) return;
  p.src_local[c] = p.local[p.cand_src[c]];
}

__global__ void __launch_bounds__(kThreads) fill_i32(int32_t* a, int64_t n, int32_t v) {
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) a[i] = v;
}

__global__ void __launch_bounds__(kThreads) seed_kernel(const int32_t* seeds, int n, int32_t* local, int32_t* nodes) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  const int s = seeds[i];
  local[s] = i;      // seeds must be distinct (the loaders pass distinct target notes)
  nodes[i] = s;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" size_t agnn_window_workspace(int64_t n_cand) {
  if (n_cand < 0 || n_cand >= (1ll << 31) - 2) return 0;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  return up((size_t)(n_cand + 1) * 4) + up(scan_workspace_bytes(n_cand + 1)) + 256;
}

extern "C" int agnn_window_subgraph(int32_t n_slots, int64_t n_cand, const int64_t* cand_ptr, const int64_t* edge_lo,
                                    const int64_t* node_lo, const int32_t* win_size, const int64_t* out_off,
                                    const int64_t* src, const int64_t* dst, const int64_t* type, int64_t* edges,
                                    int64_t* edge_id, int64_t capacity, int32_t* n_out, void* workspace,
                                    size_t workspace_bytes, agnn_stream_t stream_) {
  if (n_slots < 0 || n_cand < 0 || n_cand >= (1ll << 31) - 2 || capacity < 0 || !n_out ||
      (n_slots > 0 && (!cand_ptr || !edge_lo || !node_lo || !win_size || !out_off)) || (n_cand > 0 && (!src || !dst)) ||
      (capacity > 0 && !edges))
    return fail(AGNN_ERR_ARG, "window_subgraph: bad arguments");
  const size_t need = agnn_window_workspace(n_cand);
  if (!workspace || workspace_bytes < need || !aligned16(workspace))
    return fail(AGNN_ERR_WORKSPACE, "window_subgraph: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaStream_t stream = (cudaStream_t)stream_;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  WindowParams p;
  p.n_slots = n_slots; p.n_cand = n_cand; p.cand_ptr = cand_ptr; p.edge_lo = edge_lo; p.node_lo = node_lo;
  p.win_size = win_size; p.out_off = out_off; p.src = src; p.dst = dst; p.type = type;
  p.flag = (int32_t*)workspace;
  p.edges = edges; p.edge_id = edge_id; p.capacity = capacity; p.n_out = n_out;
  int32_t* scan_ws = (int32_t*)((char*)workspace + up((size_t)(n_cand + 1) * 4));
  const int blocks = (int)ceil_div(n_cand + 1, kThreads);
  window_flag_kernel<<<blocks, kThreads, 0, stream>>>(p);
  int rc = exclusive_scan_i32(p.flag, n_cand + 1, scan_ws, stream);
  if (rc) return rc;
  window_emit_kernel<<<blocks, kThreads, 0, stream>>>(p);
  return check_launch("window_subgraph");
}

extern "C" int agnn_sample_init(int32_t n_nodes, const int32_t* seeds, int32_t n_seeds, int32_t* local, int32_t* nodes,
                                agnn_stream_t stream_) {
  if (n_nodes < 0 || n_seeds < 0 || !local || (n_seeds > 0 && (!seeds || !nodes)))
    return fail(AGNN_ERR_ARG, "sample_init: bad arguments");
  cudaStream_t stream = (cudaStream_t)stream_;
  if (n_nodes > 0) fill_i32<<<(int)min((int64_t)kNumSM * 8, ceil_div(n_nodes, kThreads)), kThreads, 0, stream>>>(local, n_nodes, -1);
  if (n_seeds > 0) seed_kernel<<<(int)ceil_div(n_seeds, kThreads), kThreads, 0, stream>>>(seeds, n_seeds, local, nodes);
  return check_launch("sample_init");
}

extern "C" size_t agnn_sample_hop_workspace(int32_t n_rel, int32_t frontier_n) {
  return scan_workspace_bytes((int64_t)n_rel * frontier_n + 1) + 256;
}

// phase 1: counts[r * F + f] = edges relation r samples for frontier node f, exclusive-scanned in place
// (counts[R * F] = total candidates of the hop).
extern "C" int agnn_sample_hop_count(int32_t n_rel, int32_t n_nodes, const int32_t* rowptr, const int32_t* nodes,
                                     int32_t frontier_lo, int32_t frontier_n, int32_t fanout, int32_t* counts,
                                     void* workspace, size_t workspace_bytes, agnn_stream_t stream_) {
  if (n_rel < 1 || n_nodes < 0 || frontier_n < 0 || fanout > kMaxFanout || !rowptr || !counts || (frontier_n > 0 && !nodes))
    return fail(AGNN_ERR_ARG, "sample_hop_count: bad arguments (fanout <= %d)", kMaxFanout);
  if (!workspace || workspace_bytes < agnn_sample_hop_workspace(n_rel, frontier_n))
    return fail(AGNN_ERR_WORKSPACE, "sample_hop_count: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  HopParams p;
  memset(&p, 0, sizeof(p));
  p.n_rel = n_rel; p.n_nodes = n_nodes; p.fanout = fanout; p.rowptr = rowptr; p.nodes = nodes;
  p.frontier_lo = frontier_lo; p.frontier_n = frontier_n; p.counts = counts;
  const int64_t n = (int64_t)n_rel * frontier_n + 1;
  hop_count_kernel<<<(int)ceil_div(n, kThreads), kThreads, 0, stream>>>(p);
  return exclusive_scan_i32(counts, n, (int32_t*)workspace, stream);
}

// phase 2: draw the samples, discover new nodes in the sequential algorithm's order, relabel.
// n_cand = counts[R * F] read back by the caller.  flag: [n_cand + 1] scratch; first_pos: [n_nodes] scratch.
// On return *n_new_out (device) = number of nodes appended to `nodes` (at n_known ...).
extern "C" int agnn_sample_hop_draw(int32_t n_rel, int32_t n_nodes, const int32_t* rowptr, const int32_t* col,
                                    int32_t* nodes, int32_t frontier_lo, int32_t frontier_n, int32_t n_known,
                                    int32_t hop, int32_t fanout, uint64_t seed, const int32_t* counts, int32_t n_cand,
                                    int32_t* cand_slot, int32_t* cand_src, int32_t* cand_dst, int32_t* src_local,
                                    int32_t* local, int32_t* first_pos, int32_t* flag, void* workspace,
                                    size_t workspace_bytes, agnn_stream_t stream_) {
  if (n_rel < 1 || n_nodes < 0 || frontier_n < 0 || n_cand < 0 || fanout > kMaxFanout || !rowptr || !counts || !local ||
      !first_pos || !flag)
    return fail(AGNN_ERR_ARG, "sample_hop_draw: bad arguments");
  if (!workspace || workspace_bytes < scan_workspace_bytes((int64_t)n_cand + 1))
    return fail(AGNN_ERR_WORKSPACE, "sample_hop_draw: workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  HopParams p;
  memset(&p, 0, sizeof(p));
  p.n_rel = n_rel; p.n_nodes = n_nodes; p.hop = hop; p.fanout = fanout; p.seed = seed; p.rowptr = rowptr; p.col = col;
  p.nodes = nodes; p.nodes_out = nodes; p.frontier_lo = frontier_lo; p.frontier_n = frontier_n; p.n_known = n_known;
  p.counts = const_cast<int32_t*>(counts); p.cand_slot = cand_slot; p.cand_src = cand_src; p.cand_dst = cand_dst;
  p.flag = flag; p.local = local; p.first_pos = first_pos; p.src_local = src_local; p.n_cand = n_cand;
  if (n_nodes > 0) fill_i32<<<(int)min((int64_t)kNumSM * 8, ceil_div(n_nodes, kThreads)), kThreads, 0, stream>>>(first_pos, n_nodes, INT_MAX);
  const int work = n_rel * frontier_n;
  if (work > 0) hop_sample_kernel<<<(int)ceil_div(work, kThreads), kThreads, 0, stream>>>(p);
  const int cblocks = (int)ceil_div((int64_t)n_cand + 1, kThreads);
  hop_flag_kernel<<<cblocks, kThreads, 0, stream>>>(p);
  int rc = exclusive_scan_i32(flag, (int64_t)n_cand + 1, (int32_t*)workspace, stream);
  if (rc) return rc;
  if (n_cand > 0) {
    hop_assign_kernel<<<cblocks, kThreads, 0, stream>>>(p);
    hop_relabel_kernel<<<cblocks, kThreads, 0, stream>>>(p);
  }
  return check_launch("sample_hop_draw");
}
