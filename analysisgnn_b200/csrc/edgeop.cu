// Fused per-edge message + segmented reduction for the alternative conv blocks of the reference
// (analysisgnn/models/core/gnn.py): ResGatedGraphConv (:243-258, sum_j sigmoid(a_i + b_j) * c_j), OnsetEmbedding
// (:300-311) and RelEdgeConv (:99-106) (sum_j |a_i - b_j|).  The reference materialises [E, F] tensors per stage
// (index_select x 2, the gate, the product); here one warp per destination row walks the row's CSR segment, forms every
// message in registers and adds it up in CSR (= input edge) order -- nothing per-edge touches memory, no atomics.
// The backward is the same walk: the row-side gradient on the forward CSR, the neighbour-side gradients on the
// transposed CSR (every edge's term is recomputed from the node matrices).
//
// Roofline: HBM.  Algorithmic bytes per launch: E * (n_nbr_operands * F * 4 + 4) + n_rows * (n_row_operands + n_out) *
// F * 4 + 4 * (n_rows + 1).
#include "common.cuh"

namespace agnn {
namespace {

constexpr int kThreads = 256;

struct EdgeOpParams {
  int n_rows, n_feat, op;
  const int32_t* rowptr;
  const int32_t* col;
  const float* row0; int64_t ld_row0;   // row-side operands (indexed by the row)
  const float* row1; int64_t ld_row1;
  const float* nbr0; int64_t ld_nbr0;   // neighbour-side operands (indexed by col[k])
  const float* nbr1; int64_t ld_nbr1;
  const float* self_add; int64_t ld_self;   // optional: added to the sum before scaling
  float* out0; int64_t ld_out0;
  float* out1; int64_t ld_out1;          // second output of AGNN_EDGE_GATE_DNBR
  int mean;                              // divide by max(deg, 1)
};

__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + __expf(-x)); }
__device__ __forceinline__ float sign_(float x) { return (x > 0.f) - (x < 0.f); }   // torch.sign / |.|' (0 at 0)

template <int OP, int V>
__global__ void __launch_bounds__(kThreads) edge_op_kernel(const __grid_constant__ EdgeOpParams p) {
  const int lane = threadIdx.x & 31;
  const int F = p.n_feat;
  constexpr bool kTwoRow = OP == AGNN_EDGE_ABSDIFF_DROW || OP == AGNN_EDGE_GATE_DROW || OP == AGNN_EDGE_GATE_DNBR;
  constexpr bool kTwoNbr = OP != AGNN_EDGE_ABSDIFF && OP != AGNN_EDGE_ABSDIFF_DROW;
  constexpr bool kTwoOut = OP == AGNN_EDGE_GATE_DNBR;
  for (int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < p.n_rows; row += gridDim.x * (kThreads / 32)) {
    float r0[V][4], r1[V][4], acc0[V][4], acc1[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = (v * 32 + lane) * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) acc0[v][e] = acc1[v][e] = r0[v][e] = r1[v][e] = 0.f;
      if (c < F) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.row0 + (int64_t)row * p.ld_row0 + c));
        r0[v][0] = t.x; r0[v][1] = t.y; r0[v][2] = t.z; r0[v][3] = t.w;
        if (kTwoRow) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(p.row1 + (int64_t)row * p.ld_row1 + c));
          r1[v][0] = u.x; r1[v][1] = u.y; r1[v][2] = u.z; r1[v][3] = u.w;
        }
      }
    }
    const int beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
    for (int k = beg; k < end; k += 2) {
      int idx[2];
      float n0[2][V][4], n1[2][V][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) idx[u] = (k + u < end) ? __ldg(p.col + k + u) : -1;
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (idx[u] >= 0) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            const int c = (v * 32 + lane) * 4;
            if (c < F) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.nbr0 + (int64_t)idx[u] * p.ld_nbr0 + c));
              n0[u][v][0] = t.x; n0[u][v][1] = t.y; n0[u][v][2] = t.z; n0[u][v][3] = t.w;
              if (kTwoNbr) {
                const float4 w = __ldg(reinterpret_cast<const float4*>(p.nbr1 + (int64_t)idx[u] * p.ld_nbr1 + c));
                n1[u][v][0] = w.x; n1[u][v][1] = w.y; n1[u][v][2] = w.z; n1[u][v][3] = w.w;
              }
            }
          }
        }
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (idx[u] >= 0) {
#pragma unroll
          for (int v = 0; v < V; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = r0[v][e], b = r1[v][e], x = n0[u][v][e], y = n1[u][v][e];
              if (OP == AGNN_EDGE_ABSDIFF) {                      // row0 = a_i, nbr0 = b_j
                acc0[v][e] += fabsf(a - x);
              } else if (OP == AGNN_EDGE_ABSDIFF_DROW) {          // row0 = g_i, row1 = a_i, nbr0 = b_j
                acc0[v][e] += a * sign_(b - x);
              } else if (OP == AGNN_EDGE_ABSDIFF_DNBR) {          // row0 = b_j, nbr0 = g_i, nbr1 = a_i (transposed CSR)
                acc0[v][e] -= x * sign_(y - a);
              } else if (OP == AGNN_EDGE_GATE) {                  // row0 = a_i, nbr0 = b_j, nbr1 = c_j
                acc0[v][e] = fmaf(sigmoid_(a + x), y, acc0[v][e]);
              } else if (OP == AGNN_EDGE_GATE_DROW) {             // row0 = g_i, row1 = a_i, nbr0 = b_j, nbr1 = c_j
                const float s = sigmoid_(b + x);
                acc0[v][e] = fmaf(a * y, s * (1.f - s), acc0[v][e]);
              } else {                                            // GATE_DNBR: row0 = b_j, row1 = c_j, nbr0 = g_i, nbr1 = a_i
                const float s = sigmoid_(y + a);
                acc0[v][e] = fmaf(x * b, s * (1.f - s), acc0[v][e]);      // d b_j
                acc1[v][e] = fmaf(x, s, acc1[v][e]);                      // d c_j
              }
            }
        }
    }
    const float scale = p.mean ? 1.f / (float)max(end - beg, 1) : 1.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int c = (v * 32 + lane) * 4;
      if (c < F) {
        float o[4] = {acc0[v][0], acc0[v][1], acc0[v][2], acc0[v][3]};
        if (p.self_add) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.self_add + (int64_t)row * p.ld_self + c));
          o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
        }
        *reinterpret_cast<float4*>(p.out0 + (int64_t)row * p.ld_out0 + c) =
            make_float4(o[0] * scale, o[1] * scale, o[2] * scale, o[3] * scale);
        if (kTwoOut)
          *reinterpret_cast<float4*>(p.out1 + (int64_t)row * p.ld_out1 + c) =
              make_float4(acc1[v][0] * scale, acc1[v][1] * scale, acc1[v][2] * scale, acc1[v][3] * scale);
      }
    }
  }
}

template <int OP>
int launch_op(const EdgeOpParams& p, cudaStream_t st) {
  int64_t blocks = ceil_div(p.n_rows, kThreads / 32);
  if (blocks > (int64_t)kNumSM * 8) blocks = (int64_t)kNumSM * 8;
  const int vecs = p.n_feat / 4;
  if (vecs <= 32) edge_op_kernel<OP, 1><<<(unsigned)blocks, kThreads, 0, st>>>(p);
  else if (vecs <= 64) edge_op_kernel<OP, 2><<<(unsigned)blocks, kThreads, 0, st>>>(p);
  else edge_op_kernel<OP, 4><<<(unsigned)blocks, kThreads, 0, st>>>(p);
  return check_launch("edge_op");
}

}  // namespace
}  // namespace agnn

using namespace agnn;

extern "C" int agnn_edge_op(int op, int32_t n_rows, int32_t n_feat, const int32_t* rowptr, const int32_t* col,
                            const float* row0, int64_t ld_row0, const float* row1, int64_t ld_row1, const float* nbr0,
                            int64_t ld_nbr0, const float* nbr1, int64_t ld_nbr1, const float* self_add, int64_t ld_self,
                            int mean, float* out0, int64_t ld_out0, float* out1, int64_t ld_out1, agnn_stream_t stream) {
  if (op < AGNN_EDGE_ABSDIFF || op > AGNN_EDGE_GATE_DNBR) return fail(AGNN_ERR_ARG, "edge_op: unknown op %d", op);
  if (n_rows < 0 || n_feat <= 0 || n_feat % 4 || n_feat > 512 || !rowptr || !row0 || !nbr0 || !out0)
    return fail(AGNN_ERR_UNSUPPORTED, "edge_op: n_feat must be a multiple of 4 up to 512 and the operands non-null");
  const bool two_row = op == AGNN_EDGE_ABSDIFF_DROW || op == AGNN_EDGE_GATE_DROW || op == AGNN_EDGE_GATE_DNBR;
  const bool two_nbr = op != AGNN_EDGE_ABSDIFF && op != AGNN_EDGE_ABSDIFF_DROW;
  if ((two_row && !row1) || (two_nbr && !nbr1) || (op == AGNN_EDGE_GATE_DNBR && !out1))
    return fail(AGNN_ERR_ARG, "edge_op: op %d misses an operand", op);
  const void* ptrs[] = {row0, row1, nbr0, nbr1, self_add, out0, out1};
  const int64_t lds[] = {ld_row0, ld_row1, ld_nbr0, ld_nbr1, ld_self, ld_out0, ld_out1};
  for (int i = 0; i < 7; ++i)
    if (ptrs[i] && (!aligned16(ptrs[i]) || lds[i] % 4))
      return fail(AGNN_ERR_ARG, "edge_op: matrices must be 16-byte aligned with row strides that are multiples of 4");
  if (n_rows == 0) return AGNN_OK;
  EdgeOpParams p;
  p.n_rows = n_rows; p.n_feat = n_feat; p.op = op; p.rowptr = rowptr; p.col = col;
  p.row0 = row0; p.ld_row0 = ld_row0; p.row1 = row1; p.ld_row1 = ld_row1;
  p.nbr0 = nbr0; p.ld_nbr0 = ld_nbr0; p.nbr1 = nbr1; p.ld_nbr1 = ld_nbr1;
  p.self_add = self_add; p.ld_self = ld_self; p.out0 = out0; p.ld_out0 = ld_out0; p.out1 = out1; p.ld_out1 = ld_out1;
  p.mean = mean;
  cudaStream_t st = (cudaStream_t)stream;
  switch (op) {
    case AGNN_EDGE_ABSDIFF: return launch_op<AGNN_EDGE_ABSDIFF>(p, st);
    case AGNN_EDGE_ABSDIFF_DROW: return launch_op<AGNN_EDGE_ABSDIFF_DROW>(p, st);
    case AGNN_EDGE_ABSDIFF_DNBR: return launch_op<AGNN_EDGE_ABSDIFF_DNBR>(p, st);
    case AGNN_EDGE_GATE: return launch_op<AGNN_EDGE_GATE>(p, st);
    case AGNN_EDGE_GATE_DROW: return launch_op<AGNN_EDGE_GATE_DROW>(p, st);
    default: return launch_op<AGNN_EDGE_GATE_DNBR>(p, st);
  }
}
