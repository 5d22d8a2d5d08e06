// Score-graph construction on the GPU.
//
// Replaces hetero_graph_from_note_array (analysisgnn/utils/hgraph.py:214-300, rest_array=None,
// pot_edge_dist=0): an O(N^2) Python loop with one np.where per note and edge type.  Here every note
// finds its onset / consecutive / during partners by binary search in the onset-sorted note list
// (O(N log N)), rest edges come from a counting sort of the note end times (agnn_csr_build), and the
// edges are written in EXACTLY the reference's emission order, so the result is bit-identical:
//
//   for each score: for each note i (ascending):  type 0  onset_j == onset_i, j != i      (hgraph.py:233-237)
//                                                  type 1  onset_j == onset_i + dur_i      (:244-247)
//                                                  type 2  onset_i < onset_j < end_i       (:255-259)
//                   then for each end time et (ascending, the last one excluded) at which no note starts:
//                       every note ending at et -> every note at the next onset            (:273-285)
//                       (no later onset: the reference's argmin over an all-inf array selects ALL notes)
//
// A batch of scores is built in one call; node ids are offset by the score's first note (collation).
// Integer / HBM-bound; bytes per note ~ 8 (onset, duration) + 24 per emitted edge.
#include <cstring>

#include "scan.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;

struct ScoreParams {
  int n_scores, n_notes;
  int64_t key_slots;
  const int32_t* score_ptr;  // [S+1]
  const int32_t* key_base;   // [S+1]: slot range of each score's end-time axis (span + 1 slots)
  const int32_t* onset;
  const int32_t* duration;
  int32_t* score_of;         // [N]
  int32_t* max_end;          // [S]
  int64_t* key;              // [N] csr row = end-time slot
  int64_t* idx;              // [N] csr col = note index
  const int32_t* rowptr;     // [K+1] notes per end-time slot
  const int32_t* by_end;     // [N] note indices sorted by (end, index)
  int32_t* counts;           // [N + K + 1] edges per emitter, in output order; scanned in place
  int64_t* edges;            // [3][capacity]
  int64_t capacity;
  int32_t* n_edges;
};

__device__ __forceinline__ int lower_bound(const int32_t* a, int lo, int hi, int v) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound(const int32_t* a, int lo, int hi, int v) {
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kThreads) note_keys_kernel(const __grid_constant__ ScoreParams p) {
  const int i = blockIdx.x * kThreads + threadIdx.x;
  if (i >= p.n_notes) return;
  const int s = upper_bound(p.score_ptr, 0, p.n_scores + 1, i) - 1;
  const int lo = __ldg(p.score_ptr + s);
  const int e = __ldg(p.onset + i) + __ldg(p.duration + i);
  p.score_of[i] = s;
  atomicMax(p.max_end + s, e);
  p.key[i] = (int64_t)__ldg(p.key_base + s) + (e - __ldg(p.onset + lo));
  p.idx[i] = i;
}

struct NoteRanges { int same_lo, same_hi, cons_lo, cons_hi; };

__device__ __forceinline__ NoteRanges note_ranges(const ScoreParams& p, int i, int lo, int hi) {
  const int o = __ldg(p.onset + i), e = o + __ldg(p.duration + i);
  NoteRanges r;
  r.same_lo = lower_bound(p.onset, lo, hi, o);
  r.same_hi = upper_bound(p.onset, lo, hi, o);
  r.cons_lo = lower_bound(p.onset, lo, hi, e);
  r.cons_hi = upper_bound(p.onset, lo, hi, e);
  return r;
}

// rest-edge destinations of end time `et` in score [lo, hi): empty unless no note starts at et
__device__ __forceinline__ bool rest_targets(const ScoreParams& p, int s, int lo, int hi, int et, int& d_lo, int& d_hi) {
  if (et >= __ldg(p.max_end + s)) return false;                          // np.sort(unique(end))[:-1]
  const int at = lower_bound(p.onset, lo, hi, et);
  if (at < hi && __ldg(p.onset + at) == et) return false;                // a note starts here: no rest
  const int nxt = upper_bound(p.onset, lo, hi, et);
  if (nxt >= hi) { d_lo = lo; d_hi = hi; return true; }                  // quirk: all notes (hgraph.py:277-279)
  d_lo = nxt;
  d_hi = upper_bound(p.onset, lo, hi, __ldg(p.onset + nxt));
  return true;
}

// output order inside a score: its notes' counts first, then its end-time slots
__device__ __forceinline__ int64_t pos_note(const ScoreParams& p, int i, int s) { return (int64_t)i + __ldg(p.key_base + s); }
__device__ __forceinline__ int64_t pos_slot(const ScoreParams& p, int64_t kv, int s) {
  return (int64_t)__ldg(p.score_ptr + s + 1) + kv;
}

__global__ void __launch_bounds__(kThreads) count_kernel(const __grid_constant__ ScoreParams p) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t < p.n_notes) {
    const int i = (int)t, s = p.score_of[i];
    const int lo = __ldg(p.score_ptr + s), hi = __ldg(p.score_ptr + s + 1);
    const NoteRanges r = note_ranges(p, i, lo, hi);
    const int c = (r.same_hi - r.same_lo - 1) + (r.cons_hi - r.cons_lo) + max(r.cons_lo - r.same_hi, 0);
    p.counts[pos_note(p, i, s)] = c;
  } else if (t < p.n_notes + p.key_slots) {
    const int64_t kv = t - p.n_notes;
    const int s = upper_bound(p.key_base, 0, p.n_scores + 1, (int)kv) - 1;
    const int deg = __ldg(p.rowptr + kv + 1) - __ldg(p.rowptr + kv);
    int c = 0;
    if (deg > 0) {
      const int lo = __ldg(p.score_ptr + s), hi = __ldg(p.score_ptr + s + 1);
      const int et = __ldg(p.onset + lo) + (int)(kv - __ldg(p.key_base + s));
      int d_lo, d_hi;
      if (rest_targets(p, s, lo, hi, et, d_lo, d_hi)) c = deg * (d_hi - d_lo);
    }
    p.counts[pos_slot(p, kv, s)] = c;
  } else if (t == p.n_notes + p.key_slots) {
    p.counts[t] = 0;                                                     // scan slot that receives the total
  }
}

__device__ __forceinline__ void put(const ScoreParams& p, int64_t at, int64_t src, int64_t dst, int64_t type) {
  if (at < p.capacity) {
    p.edges[at] = src;
    p.edges[p.capacity + at] = dst;
    p.edges[2 * p.capacity + at] = type;
  }
}

__global__ void __launch_bounds__(kThreads) emit_kernel(const __grid_constant__ ScoreParams p) {
  const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (t == 0) *p.n_edges = p.counts[p.n_notes + p.key_slots];
  if (t < p.n_notes) {
    const int i = (int)t, s = p.score_of[i];
    const int lo = __ldg(p.score_ptr + s), hi = __ldg(p.score_ptr + s + 1);
    const NoteRanges r = note_ranges(p, i, lo, hi);
    int64_t at = p.counts[pos_note(p, i, s)];
    for (int j = r.same_lo; j < r.same_hi; ++j)
      if (j != i) put(p, at++, i, j, 0);
    for (int j = r.cons_lo; j < r.cons_hi; ++j) put(p, at++, i, j, 1);
    for (int j = r.same_hi; j < r.cons_lo; ++j) put(p, at++, i, j, 2);
  } else if (t < p.n_notes + p.key_slots) {
    const int64_t kv = t - p.n_notes;
    const int beg = __ldg(p.rowptr + kv), end = __ldg(p.rowptr + kv + 1);
    if (end == beg) return;
    const int s = upper_bound(p.key_base, 0, p.n_scores + 1, (int)kv) - 1;
    const int lo = __ldg(p.score_ptr + s), hi = __ldg(p.score_ptr + s + 1);
    const int et = __ldg(p.onset + lo) + (int)(kv - __ldg(p.key_base + s));
    int d_lo, d_hi;
    if (!rest_targets(p, s, lo, hi, et, d_lo, d_hi)) return;
    int64_t at = p.counts[pos_slot(p, kv, s)];
    for (int k = beg; k < end; ++k) {
      const int src = __ldg(p.by_end + k);
      for (int j = d_lo; j < d_hi; ++j) put(p, at++, src, j, 3);
    }
  }
}

struct Layout {
  size_t off_score_of, off_max_end, off_key, off_idx, off_rowptr, off_col, off_perm, off_status, off_counts,
      off_scan, off_csr, csr_bytes, total;
};

Layout make_layout(int32_t n_notes, int32_t n_scores, int64_t key_slots) {
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  Layout l;
  size_t o = 0;
  l.off_score_of = o; o += up((size_t)n_notes * 4);
  l.off_max_end = o; o += up((size_t)n_scores * 4);
  l.off_key = o; o += up((size_t)n_notes * 8);
  l.off_idx = o; o += up((size_t)n_notes * 8);
  l.off_rowptr = o; o += up((size_t)(key_slots + 1) * 4);
  l.off_col = o; o += up((size_t)(n_notes + 1) * 4);
  l.off_perm = o; o += up((size_t)(n_notes + 1) * 4);
  l.off_status = o; o += 256;
  l.off_counts = o; o += up((size_t)(n_notes + key_slots + 1) * 4);
  l.off_scan = o; o += up(scan_workspace_bytes(n_notes + key_slots + 1));
  agnn_coo_t seg;
  memset(&seg, 0, sizeof(seg));
  seg.n_edges = n_notes; seg.n_rows = (int32_t)key_slots; seg.n_cols = n_notes; seg.n_rel = 1;
  l.csr_bytes = agnn_csr_build_workspace(1, &seg);
  l.off_csr = o; o += up(l.csr_bytes);
  l.total = o;
  return l;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" size_t agnn_score_graph_workspace(int32_t n_notes, int32_t n_scores, int64_t key_slots) {
  if (n_notes < 0 || n_scores < 1 || key_slots < 1 || key_slots >= (1ll << 31)) return 0;
  return make_layout(n_notes, n_scores, key_slots).total;
}

extern "C" int agnn_score_graph_build(int32_t n_scores, const int32_t* score_ptr, const int32_t* key_base,
                                      const int32_t* onset, const int32_t* duration, int32_t n_notes,
                                      int64_t key_slots, int64_t* edges, int64_t capacity, int32_t* n_edges,
                                      void* workspace, size_t workspace_bytes, agnn_stream_t stream_) {
  if (n_scores < 1 || n_notes < 0 || key_slots < 1 || key_slots >= (1ll << 31) || capacity < 0 || !score_ptr ||
      !key_base || !n_edges || (n_notes > 0 && (!onset || !duration)) || (capacity > 0 && !edges))
    return fail(AGNN_ERR_ARG, "score_graph_build: bad arguments");
  const Layout l = make_layout(n_notes, n_scores, key_slots);
  if (!workspace || workspace_bytes < l.total || !aligned16(workspace))
    return fail(AGNN_ERR_WORKSPACE, "score_graph_build: workspace %zu < %zu bytes", workspace_bytes, l.total);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  ScoreParams p;
  p.n_scores = n_scores; p.n_notes = n_notes; p.key_slots = key_slots;
  p.score_ptr = score_ptr; p.key_base = key_base; p.onset = onset; p.duration = duration;
  p.score_of = (int32_t*)(ws + l.off_score_of);
  p.max_end = (int32_t*)(ws + l.off_max_end);
  p.key = (int64_t*)(ws + l.off_key);
  p.idx = (int64_t*)(ws + l.off_idx);
  int32_t* rowptr = (int32_t*)(ws + l.off_rowptr);
  int32_t* col = (int32_t*)(ws + l.off_col);
  int32_t* perm = (int32_t*)(ws + l.off_perm);
  int32_t* status = (int32_t*)(ws + l.off_status);
  p.rowptr = rowptr; p.by_end = col;
  p.counts = (int32_t*)(ws + l.off_counts);
  p.edges = edges; p.capacity = capacity; p.n_edges = n_edges;
  if (n_notes == 0) {
    if (cudaMemsetAsync(n_edges, 0, 4, stream) != cudaSuccess) return check_launch("score_graph_build memset");
    return AGNN_OK;
  }
  // INT_MIN as the identity of atomicMax
  if (cudaMemsetAsync(p.max_end, 0x80, (size_t)n_scores * 4, stream) != cudaSuccess ||
      cudaMemsetAsync(status, 0, 4, stream) != cudaSuccess)
    return check_launch("score_graph_build memset");
  const int note_blocks = (int)ceil_div(n_notes, kThreads);
  note_keys_kernel<<<note_blocks, kThreads, 0, stream>>>(p);
  agnn_coo_t seg;
  memset(&seg, 0, sizeof(seg));
  seg.row = p.key; seg.col = p.idx; seg.etype = nullptr;
  seg.n_edges = n_notes; seg.n_rows = (int32_t)key_slots; seg.n_cols = n_notes; seg.n_rel = 1;
  int rc = agnn_csr_build(1, &seg, rowptr, col, perm, status, nullptr, nullptr, ws + l.off_csr, l.csr_bytes, stream_);
  if (rc) return rc;
  const int64_t emitters = (int64_t)n_notes + key_slots + 1;
  const int blocks = (int)ceil_div(emitters, kThreads);
  count_kernel<<<blocks, kThreads, 0, stream>>>(p);
  rc = exclusive_scan_i32(p.counts, emitters, (int32_t*)(ws + l.off_scan), stream);
  if (rc) return rc;
  emit_kernel<<<blocks, kThreads, 0, stream>>>(p);
  return check_launch("score_graph_build");
}
