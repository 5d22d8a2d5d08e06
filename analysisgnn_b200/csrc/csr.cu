// GPU COO -> relation-major CSR, bit-identical to a stable sort by (relation, row).
//
// Replaces the per-relation mask / edge_index[:, mask] of the reference's
// HeteroConv.forward (analysisgnn/models/core/hgnn.py:480-483) and the CSC
// conversion inside PyG's NeighborSampler (third-party).  See include/agnn.h.
//
// Pipeline (all segments of a call share every launch; blocks never straddle
// segments, the block -> segment map is a <=32-entry table in kernel params):
//   zero      rowptr blocks and the per-key cursors
//   count     histogram of (relation, row) keys           (int atomics: result is order-free)
//   scan      in-place exclusive scan per segment          (tile scan, tile-sum scan, add)
//   fill      perm[rowptr[key] + cursor++] = edge id       (slot order within a row is arbitrary ...)
//   finalize  ... so each row's slots are sorted by edge id => stable order; col = col_in[perm]
//             rows longer than 32 go to a list that `long_rows` sorts with a block-wide
//             all-ascending bitonic network (shared memory up to 8192 entries, else in place).
// HBM traffic per edge: 2 x 24 B COO reads + 4 B perm write/read + 8 B col/perm write.
#include <climits>

#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;
constexpr int kKeysPerTile = 4096;    // scan tile
constexpr int kEdgesPerBlock = 1024;  // count / fill
constexpr int kShortRow = 32;
constexpr int kRankSort = 8;       // rows up to this long are ordered in registers (finalize)
constexpr int kSmemSort = 8192;

struct SegTable {
  int n_seg;
  agnn_coo_t seg[AGNN_MAX_SEG];
  int key_tile_start[AGNN_MAX_SEG + 1];
  int edge_blk_start[AGNN_MAX_SEG + 1];
};

__device__ __forceinline__ int find_seg(const int* start, int n_seg, int blk) {
  int s = 0;
  while (s + 1 < n_seg && blk >= start[s + 1]) ++s;
  return s;
}

__device__ __forceinline__ int64_t seg_keys(const agnn_coo_t& g) {
  return (int64_t)g.n_rel * (g.n_rows + 1);
}

__global__ void __launch_bounds__(kThreads) zero_kernel(const __grid_constant__ SegTable tab, int32_t* rowptr,
                                                         int32_t* cursor) {
  const int s = find_seg(tab.key_tile_start, tab.n_seg, blockIdx.x);
  const agnn_coo_t& g = tab.seg[s];
  const int64_t base = (int64_t)(blockIdx.x - tab.key_tile_start[s]) * kKeysPerTile;
  const int64_t n = seg_keys(g);
  for (int i = threadIdx.x; i < kKeysPerTile; i += kThreads) {
    int64_t k = base + i;
    if (k < n) {
      rowptr[g.rowptr_off + k] = 0;
      cursor[g.rowptr_off + k] = 0;
    }
  }
}

// returns the key of edge e, or -1 when the edge is dropped
__device__ __forceinline__ int64_t edge_key(const agnn_coo_t& g, int64_t e, int32_t* status) {
  const int64_t r = g.etype ? g.etype[e] : 0;
  if (r < 0 || r >= g.n_rel) return -1;
  const int64_t row = g.row[e], col = g.col[e];
  if (row < 0 || row >= g.n_rows || col < 0 || col >= g.n_cols) {
    *status = 1;
    return -1;
  }
  return r * (g.n_rows + 1) + row;
}

template <bool FILL>
__global__ void __launch_bounds__(kThreads) edge_kernel(const __grid_constant__ SegTable tab, int32_t* rowptr,
                                                         int32_t* cursor, int32_t* perm, int32_t* status) {
  const int s = find_seg(tab.edge_blk_start, tab.n_seg, blockIdx.x);
  const agnn_coo_t& g = tab.seg[s];
  const int64_t base = (int64_t)(blockIdx.x - tab.edge_blk_start[s]) * kEdgesPerBlock;
#pragma unroll
  for (int i = 0; i < kEdgesPerBlock / kThreads; ++i) {
    const int64_t e = base + i * kThreads + threadIdx.x;
    if (e >= g.n_edges) continue;
    const int64_t key = edge_key(g, e, status);
    if (key < 0) continue;
    if (!FILL) {
      atomicAdd(&rowptr[g.rowptr_off + key], 1);
    } else {
      const int pos = atomicAdd(&cursor[g.rowptr_off + key], 1);
      perm[g.edge_off + rowptr[g.rowptr_off + key] + pos] = (int32_t)e;
    }
  }
}

// exclusive scan of 256 values across the block; returns the prefix of this thread, total in `total`
__device__ __forceinline__ int block_exclusive_scan(int v, int& total, int* warp_sums) {
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
  for (int w = 0; w < kThreads / 32; ++w) {
    int sw = warp_sums[w];
    if (w < warp) wprefix += sw;
    tot += sw;
  }
  __syncthreads();
  total = tot;
  return wprefix + inc - v;
}

__global__ void __launch_bounds__(kThreads) scan_tile_kernel(const __grid_constant__ SegTable tab, int32_t* rowptr,
                                                              int32_t* tile_sums) {
  __shared__ int warp_sums[kThreads / 32];
  const int s = find_seg(tab.key_tile_start, tab.n_seg, blockIdx.x);
  const agnn_coo_t& g = tab.seg[s];
  const int64_t base = (int64_t)(blockIdx.x - tab.key_tile_start[s]) * kKeysPerTile;
  const int64_t n = seg_keys(g);
  int32_t* p = rowptr + g.rowptr_off;
  int carry = 0;
  for (int pass = 0; pass < kKeysPerTile / kThreads; ++pass) {
    const int64_t k = base + pass * kThreads + threadIdx.x;
    const int v = k < n ? p[k] : 0;
    int total;
    const int pre = block_exclusive_scan(v, total, warp_sums);
    if (k < n) p[k] = carry + pre;
    carry += total;
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = carry;
}

// one block per segment: exclusive scan of that segment's tile sums, in place
__global__ void __launch_bounds__(kThreads) scan_sums_kernel(const __grid_constant__ SegTable tab,
                                                              int32_t* tile_sums) {
  __shared__ int warp_sums[kThreads / 32];
  const int s = blockIdx.x;
  const int lo = tab.key_tile_start[s], hi = tab.key_tile_start[s + 1];
  int carry = 0;
  for (int t0 = lo; t0 < hi; t0 += kThreads) {
    const int t = t0 + threadIdx.x;
    const int v = t < hi ? tile_sums[t] : 0;
    int total;
    const int pre = block_exclusive_scan(v, total, warp_sums);
    if (t < hi) tile_sums[t] = carry + pre;
    carry += total;
  }
}

__global__ void __launch_bounds__(kThreads) scan_add_kernel(const __grid_constant__ SegTable tab, int32_t* rowptr,
                                                             const int32_t* tile_sums) {
  const int s = find_seg(tab.key_tile_start, tab.n_seg, blockIdx.x);
  if (blockIdx.x == tab.key_tile_start[s]) return;  // first tile of a segment: offset 0
  const agnn_coo_t& g = tab.seg[s];
  const int64_t base = (int64_t)(blockIdx.x - tab.key_tile_start[s]) * kKeysPerTile;
  const int64_t n = seg_keys(g);
  const int add = tile_sums[blockIdx.x];
  for (int i = threadIdx.x; i < kKeysPerTile; i += kThreads) {
    const int64_t k = base + i;
    if (k < n) rowptr[g.rowptr_off + k] += add;
  }
}

// One block per (segment, relation): the rows with >= AGNN_HEAVY_ROW entries in ascending row order, then (second half
// of the relation's block in `heavy`) the exclusive prefix of their chunk counts, ceil(deg / AGNN_HEAVY_CHUNK) each:
// agnn_gather_reduce finds chunk g's row, and a row's slot, by binary search instead of walking the list.
constexpr int kHeavyThreads = 1024;
constexpr int kHeavyPerThread = 8;

__global__ void __launch_bounds__(kHeavyThreads) heavy_list_kernel(const __grid_constant__ SegTable tab,
                                                                    const int32_t* __restrict__ rowptr,
                                                                    int32_t* __restrict__ heavy,
                                                                    int32_t* __restrict__ n_heavy) {
  int r = blockIdx.x, s = 0;
  while (r >= tab.seg[s].n_rel) r -= tab.seg[s++].n_rel;
  const agnn_coo_t& g = tab.seg[s];
  if (g.heavy_cap <= 0) return;
  const int32_t* rp = rowptr + g.rowptr_off + (int64_t)r * (g.n_rows + 1);
  int32_t* rows_out = heavy + g.heavy_off + (int64_t)r * 2 * g.heavy_cap;
  int32_t* chunk_out = rows_out + g.heavy_cap;
  __shared__ unsigned long long warp_tot[kHeavyThreads / 32];
  __shared__ unsigned long long carry_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long carry = 0;                       // (heavy rows so far) << 32 | chunks so far
  for (int base = 0; base < g.n_rows; base += kHeavyThreads * kHeavyPerThread) {
    // thread t: rows base + t * kHeavyPerThread .. (consecutive, so slots come out in row order)
    const int row0 = base + threadIdx.x * kHeavyPerThread;
    int bound[kHeavyPerThread + 1];
#pragma unroll
    for (int i = 0; i <= kHeavyPerThread; ++i) bound[i] = row0 + i <= g.n_rows ? rp[row0 + i] : 0;
    unsigned long long item[kHeavyPerThread], mine = 0;
#pragma unroll
    for (int i = 0; i < kHeavyPerThread; ++i) {
      const int deg = row0 + i < g.n_rows ? bound[i + 1] - bound[i] : 0;
      item[i] = deg >= AGNN_HEAVY_ROW
                    ? (1ull << 32) | (unsigned)((deg + AGNN_HEAVY_CHUNK - 1) / AGNN_HEAVY_CHUNK) : 0ull;
      mine += item[i];
    }
    unsigned long long incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned long long w = warp_tot[lane], wi = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, d);
        if (lane >= d) wi += t;
      }
      warp_tot[lane] = wi - w;                         // exclusive over warps
      if (lane == 31) carry_s = wi;                    // block total
    }
    __syncthreads();
    unsigned long long excl = carry + warp_tot[warp] + incl - mine;
    if (mine) {
#pragma unroll
      for (int i = 0; i < kHeavyPerThread; ++i) {
        if (item[i]) {
          const long long slot = (long long)(excl >> 32);
          if (slot < g.heavy_cap) {
            rows_out[slot] = row0 + i;
            chunk_out[slot] = (int32_t)(excl & 0xffffffffu);
          }
          excl += item[i];
        }
      }
    }
    carry += carry_s;
    __syncthreads();
  }
  if (threadIdx.x == 0) n_heavy[g.count_off + r] = (int32_t)(carry >> 32);
}

// thread per key: restore input order inside the row, emit col
__global__ void __launch_bounds__(kThreads) finalize_kernel(const __grid_constant__ SegTable tab,
                                                             const int32_t* rowptr, int32_t* col, int32_t* perm,
                                                             int32_t* long_count, int64_t* long_list) {
  // one key per thread (kKeysPerTile / kThreads blocks per scan tile): the two dependent round trips of a key --
  // extent, then the row's slots -- are the whole cost, so the more keys in flight the better
  constexpr int kSub = kKeysPerTile / kThreads;
  const int tile = blockIdx.x / kSub;
  const int s = find_seg(tab.key_tile_start, tab.n_seg, tile);
  const agnn_coo_t& g = tab.seg[s];
  const int64_t k = (int64_t)(tile - tab.key_tile_start[s]) * kKeysPerTile + (blockIdx.x % kSub) * kThreads + threadIdx.x;
  if (k >= seg_keys(g) || (k % (g.n_rows + 1)) == g.n_rows) return;
  const int beg = rowptr[g.rowptr_off + k], end = rowptr[g.rowptr_off + k + 1];
  const int deg = end - beg;
  if (deg <= 1) return;
  if (deg > kShortRow) {
    const int slot = atomicAdd(long_count, 1);
    long_list[slot] = ((int64_t)s << 40) | k;
    return;
  }
  int32_t* pp = perm + g.edge_off + beg;
  if (deg <= kRankSort) {
    // the usual row of a score graph: all slots loaded at once, every slot's rank = how many are smaller (edge ids are
    // distinct), one store each -- no dependent global accesses
    int v[kRankSort];
#pragma unroll
    for (int a = 0; a < kRankSort; ++a) v[a] = a < deg ? pp[a] : INT_MAX;
#pragma unroll
    for (int a = 0; a < kRankSort; ++a) {
      int rank = 0;
#pragma unroll
      for (int b = 0; b < kRankSort; ++b) rank += (v[b] < v[a]) ? 1 : 0;
      if (a < deg) pp[rank] = v[a];
    }
    return;
  }
  for (int a = 1; a < deg; ++a) {  // insertion sort: rows arrive almost ordered
    const int v = pp[a];
    int b = a - 1;
    while (b >= 0 && pp[b] > v) {
      pp[b + 1] = pp[b];
      --b;
    }
    pp[b + 1] = v;
  }
}

// col[k] = col_in[perm[k]] for the short rows, edge-parallel (the long rows write theirs after their block sort)
__global__ void __launch_bounds__(kThreads) gather_col_kernel(const __grid_constant__ SegTable tab, int32_t* col,
                                                               const int32_t* perm) {
  const int s = find_seg(tab.edge_blk_start, tab.n_seg, blockIdx.x);
  const agnn_coo_t& g = tab.seg[s];
  const int64_t base = (int64_t)(blockIdx.x - tab.edge_blk_start[s]) * kEdgesPerBlock;
#pragma unroll
  for (int i = 0; i < kEdgesPerBlock / kThreads; ++i) {
    const int64_t e = base + i * kThreads + threadIdx.x;
    if (e >= g.n_edges) continue;
    const int32_t src = perm[g.edge_off + e];
    // slots past the kept edges (dropped relation codes) hold stale values: guard the read
    if (src >= 0 && src < g.n_edges) col[g.edge_off + e] = (int32_t)g.col[src];
  }
}

// all comparators ascending, so virtual +inf padding beyond n never moves
__device__ void block_bitonic(int32_t* a, int n, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    const int half = k >> 1;
    for (int t = threadIdx.x; t < (n2 >> 1); t += kThreads) {
      const int blk = t / half, off = t - blk * half;
      const int lo = blk * k + off, hi = blk * k + k - 1 - off;
      if (hi < n) {
        const int x = a[lo], y = a[hi];
        if (x > y) { a[lo] = y; a[hi] = x; }
      }
    }
    __syncthreads();
    for (int j = k >> 2; j >= 1; j >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += kThreads) {
        const int lo = (t / j) * 2 * j + (t % j), hi = lo + j;
        if (hi < n) {
          const int x = a[lo], y = a[hi];
          if (x > y) { a[lo] = y; a[hi] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kThreads) long_rows_kernel(const __grid_constant__ SegTable tab,
                                                              const int32_t* rowptr, int32_t* col, int32_t* perm,
                                                              const int32_t* long_count, const int64_t* long_list) {
  __shared__ int32_t buf[kSmemSort];
  const int n_long = *long_count;
  for (int item = blockIdx.x; item < n_long; item += gridDim.x) {
    const int64_t packed = long_list[item];
    const int s = (int)(packed >> 40);
    const int64_t k = packed & ((1ll << 40) - 1);
    const agnn_coo_t& g = tab.seg[s];
    const int beg = rowptr[g.rowptr_off + k], end = rowptr[g.rowptr_off + k + 1];
    const int deg = end - beg;
    int n2 = 1;
    while (n2 < deg) n2 <<= 1;
    int32_t* pp = perm + g.edge_off + beg;
    if (n2 <= kSmemSort) {
      for (int i = threadIdx.x; i < deg; i += kThreads) buf[i] = pp[i];
      __syncthreads();
      block_bitonic(buf, deg, n2);
      for (int i = threadIdx.x; i < deg; i += kThreads) pp[i] = buf[i];
    } else {
      __syncthreads();
      block_bitonic(pp, deg, n2);
    }
    __syncthreads();
    int32_t* cc = col + g.edge_off + beg;
    for (int i = threadIdx.x; i < deg; i += kThreads) cc[i] = (int32_t)g.col[pp[i]];
    __syncthreads();
  }
}

struct Layout {
  int64_t cursor_ints, tile_ints, long_cap;
  size_t off_tiles, off_count, off_list, total;
};

int make_tables(int n_seg, const agnn_coo_t* segs, SegTable& tab, Layout& lay, bool need_ptrs) {
  if (n_seg < 1 || n_seg > AGNN_MAX_SEG || !segs) return fail(AGNN_ERR_ARG, "csr_build: n_seg must be 1..%d", AGNN_MAX_SEG);
  tab.n_seg = n_seg;
  int64_t key_tiles = 0, edge_blks = 0, max_key_end = 0, edges = 0;
  for (int s = 0; s < n_seg; ++s) {
    const agnn_coo_t& g = segs[s];
    if (g.n_edges < 0 || g.n_rows < 0 || g.n_cols < 0 || g.n_rel < 1 || g.rowptr_off < 0 || g.edge_off < 0)
      return fail(AGNN_ERR_ARG, "csr_build: segment %d has a negative size or n_rel < 1", s);
    if (need_ptrs && g.n_edges > 0 && (!g.row || !g.col)) return fail(AGNN_ERR_ARG, "csr_build: segment %d has null COO pointers", s);
    const int64_t keys = (int64_t)g.n_rel * (g.n_rows + 1);
    if (g.n_edges >= (1ll << 31) || keys >= (1ll << 31) || g.edge_off + g.n_edges >= (1ll << 31))
      return fail(AGNN_ERR_UNSUPPORTED, "csr_build: segment %d exceeds int32 indexing", s);
    tab.seg[s] = g;
    tab.key_tile_start[s] = (int)key_tiles;
    tab.edge_blk_start[s] = (int)edge_blks;
    key_tiles += ceil_div(keys, kKeysPerTile);
    edge_blks += ceil_div(g.n_edges, kEdgesPerBlock);
    if (g.rowptr_off + keys > max_key_end) max_key_end = g.rowptr_off + keys;
    edges += g.n_edges;
  }
  // (finalize launches kKeysPerTile / kThreads = 16 blocks per key tile)
  if (key_tiles >= (1ll << 27) || edge_blks >= (1ll << 31)) return fail(AGNN_ERR_UNSUPPORTED, "csr_build: too large");
  tab.key_tile_start[n_seg] = (int)key_tiles;
  tab.edge_blk_start[n_seg] = (int)edge_blks;
  lay.cursor_ints = max_key_end;
  lay.tile_ints = key_tiles;
  lay.long_cap = edges / (kShortRow + 1) + 1;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  lay.off_tiles = up((size_t)lay.cursor_ints * 4);
  lay.off_count = lay.off_tiles + up((size_t)lay.tile_ints * 4);
  lay.off_list = lay.off_count + 256;
  lay.total = lay.off_list + up((size_t)lay.long_cap * 8);
  return AGNN_OK;
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" size_t agnn_csr_build_workspace(int n_seg, const agnn_coo_t* segs) {
  SegTable tab;
  Layout lay;
  if (make_tables(n_seg, segs, tab, lay, false) != AGNN_OK) return 0;
  return lay.total;
}

extern "C" int agnn_csr_build(int n_seg, const agnn_coo_t* segs, int32_t* rowptr, int32_t* col, int32_t* perm,
                              int32_t* status, int32_t* heavy, int32_t* n_heavy, void* workspace,
                              size_t workspace_bytes, agnn_stream_t stream_) {
  SegTable tab;
  Layout lay;
  int rc = make_tables(n_seg, segs, tab, lay, true);
  if (rc != AGNN_OK) return rc;
  if (!rowptr || !status || !workspace) return fail(AGNN_ERR_ARG, "csr_build: null output pointer");
  if (workspace_bytes < lay.total)
    return fail(AGNN_ERR_WORKSPACE, "csr_build: workspace %zu < %zu bytes", workspace_bytes, lay.total);
  if (!aligned16(workspace)) return fail(AGNN_ERR_ARG, "csr_build: workspace must be 16-byte aligned");
  cudaStream_t stream = (cudaStream_t)stream_;
  char* ws = (char*)workspace;
  int32_t* cursor = (int32_t*)ws;
  int32_t* tile_sums = (int32_t*)(ws + lay.off_tiles);
  int32_t* long_count = (int32_t*)(ws + lay.off_count);
  int64_t* long_list = (int64_t*)(ws + lay.off_list);
  const int key_tiles = tab.key_tile_start[n_seg], edge_blks = tab.edge_blk_start[n_seg];

  if (cudaMemsetAsync(long_count, 0, 4, stream) != cudaSuccess) return check_launch("csr_build memset");
#define AGNN_STEP(name) if ((rc = sync_check(stream, "csr_build: " name)) != AGNN_OK) return rc
  zero_kernel<<<key_tiles, kThreads, 0, stream>>>(tab, rowptr, cursor);
  AGNN_STEP("zero");
  if (edge_blks > 0) edge_kernel<false><<<edge_blks, kThreads, 0, stream>>>(tab, rowptr, cursor, perm, status);
  AGNN_STEP("count");
  scan_tile_kernel<<<key_tiles, kThreads, 0, stream>>>(tab, rowptr, tile_sums);
  scan_sums_kernel<<<n_seg, kThreads, 0, stream>>>(tab, tile_sums);
  scan_add_kernel<<<key_tiles, kThreads, 0, stream>>>(tab, rowptr, tile_sums);
  AGNN_STEP("scan");
  if (heavy && n_heavy) {
    int n_blocks = 0;
    for (int s = 0; s < n_seg; ++s) n_blocks += segs[s].n_rel;
    heavy_list_kernel<<<n_blocks, kHeavyThreads, 0, stream>>>(tab, rowptr, heavy, n_heavy);
    AGNN_STEP("heavy lists");
  }
  if (edge_blks > 0) {
    if (!col || !perm) return fail(AGNN_ERR_ARG, "csr_build: null col/perm with edges present");
    edge_kernel<true><<<edge_blks, kThreads, 0, stream>>>(tab, rowptr, cursor, perm, status);
    AGNN_STEP("fill");
    finalize_kernel<<<key_tiles * (kKeysPerTile / kThreads), kThreads, 0, stream>>>(tab, rowptr, col, perm, long_count,
                                                                                    long_list);
    AGNN_STEP("finalize");
    gather_col_kernel<<<edge_blks, kThreads, 0, stream>>>(tab, col, perm);
    AGNN_STEP("gather col");
    long_rows_kernel<<<kNumSM * 2, kThreads, 0, stream>>>(tab, rowptr, col, perm, long_count, long_list);
    AGNN_STEP("long rows");
  }
#undef AGNN_STEP
  return check_launch("csr_build");
}
